"""GPU: codebook statistics kernels against the oracle restatement of the reference metrics, and the extraction
shell end to end (files on disk == the model's indices)."""
import collections
import os

import numpy as np
import pytest
import torch

from audiotokenization_b200 import configs, extract, metrics, synth
from audiotokenization_b200.model import BigCodecModel
from oracle import bigcodec_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(8192, 100000), (32768, 250000), (1024, 1), (16, 4097)])
def test_code_histogram_matches_bincount_and_accumulates(K, N):
    g = torch.Generator().manual_seed(K + N)
    a = torch.randint(0, K, (N,), generator=g)
    b = (torch.randn(N // 2 + 1, generator=g).abs() * K / 8).long().clamp(0, K - 1)     # skewed second update
    counts = metrics.code_histogram(a.cuda(), K)
    assert torch.equal(counts.cpu(), torch.bincount(a, minlength=K))
    metrics.code_histogram(b.cuda().view(1, -1, 1), K, counts)
    assert torch.equal(counts.cpu(), torch.bincount(torch.cat([a, b]), minlength=K))
    with pytest.raises(IndexError):
        metrics.code_histogram(torch.tensor([0, K], device="cuda"), K)


def test_metric_classes_match_reference_formulas():
    K = 8192
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(5000, generator=g).abs() * 700).long().clamp(0, K - 1).view(1, 10, 500) for _ in range(3)]
    ppl, util = metrics.CodebookPerplexity(K), metrics.CodebookUtilization(K)
    assert float(ppl.compute()) == 0.0 and float(util.compute()) == 0.0
    for b in batches:
        ppl.update(b.cuda())
        util.update(b.cuda())
    allidx = torch.cat([b.reshape(-1) for b in batches]).numpy()
    assert abs(float(ppl.compute()) - oracle.codebook_perplexity(allidx, K)) <= 1e-4 * oracle.codebook_perplexity(allidx, K)
    assert abs(float(util.compute()) - oracle.codebook_utilization(allidx, K)) < 1e-7
    assert int(ppl.total_counts) == allidx.size and int(util.used_codes.sum()) == len(set(allidx.tolist()))
    want = oracle.calculate_perplexity(collections.Counter(allidx.tolist()), K)
    got = metrics.calculate_perplexity(ppl.codebook_counts, K)
    assert abs(got[0] - want[0]) < 1e-9 and abs(got[1] - want[1]) < 1e-6 * want[1]
    cnt = collections.Counter({0: 3, 1: 1, K + 5: 4})          # out-of-range key: total only (inference_full.py:587-589)
    got, want = metrics.calculate_perplexity(cnt, K), oracle.calculate_perplexity(cnt, K)
    assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    assert metrics.calculate_perplexity(collections.Counter(), K) == 0.0
    ppl.reset()
    assert float(ppl.compute()) == 0.0


def test_extract_to_directory_end_to_end(tmp_path):
    cfg = configs.get_config("tiny")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="fp32")
    hop = int(model.encoder.hop_length)
    lens = [40 * hop, 40 * hop, 25 * hop, 40 * hop, 25 * hop + 7]
    waves = [synth.fast_synth_batch(i, 1, n)[0, 0] for i, n in enumerate(lens)]
    waves[4] = extract.prepare_waveform(waves[4], 16000, 16000, pad_to_stride=hop)[0]        # -> 26 hops
    items = [(w, "test-clean", f"61-70968-{i:04d}") for i, w in enumerate(waves)]
    saved, errors = extract.extract_to_directory(model, items, str(tmp_path), micro_batch=2, group_size=2, verbose=False)
    assert (saved, errors) == (5, 0)
    K = cfg["codec_decoder"]["codebook_size"]
    ppl = metrics.CodebookPerplexity(K)
    for i, w in enumerate(waves):
        a = np.load(os.path.join(str(tmp_path), "test-clean", "61", "70968", f"61-70968-{i:04d}.npy"))
        alone = model(w.view(1, 1, -1).cuda())["indices"]                        # the reference's one-utterance call
        assert a.dtype == np.int16 and np.array_equal(a, oracle.indices_to_int16(alone))
        ppl.update(alone)
    allidx = np.concatenate([np.load(os.path.join(str(tmp_path), "test-clean", "61", "70968", f"61-70968-{i:04d}.npy")) for i in range(5)])
    assert abs(float(ppl.compute()) - oracle.codebook_perplexity(allidx, K)) <= 1e-4 * oracle.codebook_perplexity(allidx, K)
