"""CPU: the steps on either side of the hot path: on-disk layout of extract_indices.py,
waveform preparation, and the oracle restatements of the codebook statistics."""
import collections
import os
import struct
import wave

import numpy as np
import pytest
import torch

from audiotokenization_b200 import extract
from oracle import bigcodec_oracle as oracle


@pytest.mark.parametrize("fileid,want", [
    ("103_1240_000001", ("103", "1240")),          # '_' separated (extract_indices.py:535-538)
    ("103-1240-0001", ("103", "1240")),            # '-' separated (:539-541)
    ("a_b-c", ("a", "b-c")),                       # '_' wins when both occur
    ("single", ("unknown", "unknown")),            # no separator (:542-545)
])
def test_index_file_path_matches_reference_layout(fileid, want, tmp_path):
    got = extract.index_file_path(str(tmp_path), "train-clean-100", fileid)
    assert got == os.path.join(str(tmp_path), "train-clean-100", want[0], want[1], fileid + ".npy")
    assert got == oracle.index_file_path(str(tmp_path), "train-clean-100", fileid)


def test_save_indices_shapes_and_dtype(tmp_path):
    one = np.arange(7, dtype=np.int64).reshape(7, 1) * 1000
    p = extract.save_indices(str(tmp_path), "dev", "1-2-3", one)
    a = np.load(p)
    assert a.dtype == np.int16 and a.shape == (7, 1) and a[:, 0].tolist() == (np.arange(7) * 1000).tolist()   # (T', 1) for n_q = 1
    two = np.stack([np.arange(5), np.arange(5) + 10], axis=1)
    b = np.load(extract.save_indices(str(tmp_path), "dev", "1-2-4", two))
    assert b.dtype == np.int16 and b.shape == (5, 2)                                                     # (T', n_q)
    idx = torch.arange(12).view(2, 1, 6)                        # [n_q, 1, T'] as the model returns it
    assert np.array_equal(oracle.indices_to_int16(idx), idx.squeeze(1).permute(1, 0).numpy().astype(np.int16))
    assert oracle.indices_to_int16(torch.arange(6).view(1, 1, 6)).shape == (6, 1)       # one quantizer: (T', 1), not (T',)


def test_load_wav_pcm16_and_float32(tmp_path):
    x = (np.sin(np.arange(800) * 0.05) * 0.5).astype(np.float32)
    p16 = str(tmp_path / "a.wav")
    with wave.open(p16, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes((x * 32767).astype("<i2").tobytes())
    y, sr = extract.load_wav(p16)
    assert sr == 16000 and tuple(y.shape) == (1, 800) and float((y[0] - torch.from_numpy(x)).abs().max()) < 1e-4
    pf = str(tmp_path / "b.wav")
    stereo = np.stack([x, -x], axis=1)
    data = stereo.astype("<f4").tobytes()
    with open(pf, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 2, 24000, 24000 * 8, 8, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
    y, sr = extract.load_wav(pf)
    assert sr == 24000 and tuple(y.shape) == (2, 800) and torch.equal(y[1], -y[0])


def test_prepare_waveform_pads_like_the_reference_and_resamples():
    w = torch.randn(1, 1001)
    got = extract.prepare_waveform(w, 16000, 16000, pad_to_stride=200)
    assert torch.equal(got, oracle.pad_to_stride(w, 200)) and got.shape[1] == 1200
    assert torch.equal(extract.prepare_waveform(torch.randn(1000), 16000, 16000, pad_to_stride=200)[0, :0], torch.empty(0))
    aligned = torch.randn(1, 800)
    assert extract.prepare_waveform(aligned, 16000, 16000, pad_to_stride=200).shape[1] == 800   # no extra hop when aligned
    r = extract.prepare_waveform(torch.randn(2, 24000), 24000, 16000)
    assert tuple(r.shape) == (1, 16000)


def test_oracle_codebook_statistics_closed_forms():
    K = 64
    idx = np.repeat(np.arange(16), 5)                  # 16 codes, uniform
    assert abs(oracle.codebook_perplexity(idx, K) - 16.0) < 1e-9
    assert oracle.codebook_utilization(idx, K) == 16 / 64
    assert oracle.codebook_perplexity(np.zeros(0, dtype=np.int64), K) == 0.0
    skew = np.array([0] * 3 + [1])
    p = np.array([0.75, 0.25])
    assert abs(oracle.codebook_perplexity(skew, K) - np.exp(-(p * np.log(p)).sum())) < 1e-12
    cnt = collections.Counter({0: 3, 1: 1, 99: 4})     # key 99 is out of range: counts towards the total only
    npx, px = oracle.calculate_perplexity(cnt, K)
    q = np.array([3 / 8, 1 / 8])
    ent = -(q * np.log(q)).sum()
    assert abs(px - np.exp(ent)) < 1e-12 and abs(npx - np.exp(ent / np.log(K))) < 1e-12
    assert oracle.calculate_perplexity(collections.Counter(), K) == 0.0


class _FakeModel:
    """Stands in for BigCodecModel on CPU: 'indices' = sample count and first sample, so the files identify their input."""

    def __init__(self, fail_length=None):
        self.calls = []
        self.fail_length = fail_length

    def extract_indices(self, host, micro_batch=8, rnn_batch=256):
        n, _, t = host.shape
        self.calls.append((n, t))
        if t == self.fail_length:
            raise RuntimeError("boom")
        out = np.zeros((n, 3, 1), dtype=np.int16)
        out[:, 0, 0] = t % 30000
        out[:, 1, 0] = (host[:, 0, 0].numpy() * 100).round().astype(np.int16)
        return out


def test_extract_to_directory_groups_equal_lengths_writes_layout_and_counts_errors(tmp_path):
    items = [(torch.full((400,), 0.01 * i), "dev", f"7-{i}-x") for i in range(5)]
    items += [(torch.full((600,), 0.5), "dev", "8_1_y"), (torch.zeros(0), "dev", "bad-0-0"), (torch.ones(777), "dev", "9-9-z")]
    model = _FakeModel(fail_length=777)
    saved, errors = extract.extract_to_directory(model, items, str(tmp_path), group_size=4, writers=2, verbose=False)
    assert (saved, errors) == (6, 2)                                   # empty waveform + the failing group
    assert sorted(model.calls) == [(1, 400), (1, 600), (1, 777), (4, 400)]      # equal lengths only, groups of <= 4
    for i in range(5):
        a = np.load(os.path.join(str(tmp_path), "dev", "7", str(i), f"7-{i}-x.npy"))
        assert a.dtype == np.int16 and a[:, 0].tolist() == [400, i, 0]
    assert np.load(os.path.join(str(tmp_path), "dev", "8", "1", "8_1_y.npy"))[:, 0].tolist() == [600, 50, 0]
    assert not os.path.exists(os.path.join(str(tmp_path), "dev", "9", "9", "9-9-z.npy"))


def test_extract_to_directory_bounds_its_buffer_and_retries_failed_groups_one_by_one(tmp_path):
    """Full-length files: every length unique.  The buffer must not grow with the corpus, and a group that fails is
    retried per utterance so that one bad utterance costs one error (extract_indices.py:565-574)."""
    items = ((torch.full((300 + i,), 0.01), "dev", f"1-{i}-u") for i in range(40))       # a generator: nothing is pre-loaded
    model = _FakeModel()
    seen_max = [0]
    orig = model.extract_indices

    def spy(host, **kw):
        seen_max[0] = max(seen_max[0], len(model.calls))
        return orig(host, **kw)

    model.extract_indices = spy
    saved, errors = extract.extract_to_directory(model, items, str(tmp_path), group_size=8, writers=2, verbose=False,
                                                 max_buffered_items=5)
    assert (saved, errors) == (40, 0)
    assert len(model.calls) == 40 and all(n == 1 for n, _ in model.calls)
    # with a cap of 5 waiting utterances the first encode happens after 6 items, not after the whole corpus
    assert model.calls[0][1] in range(300, 306)

    class _FailsInGroups(_FakeModel):
        def extract_indices(self, host, micro_batch=8, rnn_batch=256):
            if host.shape[0] > 1 or float(host[0, 0, 0]) > 0.9:          # any group fails; alone only the bad one does
                self.calls.append(tuple(host.shape[::2]))
                raise RuntimeError("boom")
            return super().extract_indices(host, micro_batch, rnn_batch)

    items = [(torch.full((500,), 1.0 if i == 2 else 0.1), "dev", f"2-{i}-v") for i in range(4)]
    model = _FailsInGroups()
    saved, errors = extract.extract_to_directory(model, items, str(tmp_path), group_size=4, writers=1, verbose=False)
    assert (saved, errors) == (3, 1)
    assert not os.path.exists(os.path.join(str(tmp_path), "dev", "2", "2", "2-2-v.npy"))
    assert os.path.exists(os.path.join(str(tmp_path), "dev", "2", "3", "2-3-v.npy"))
