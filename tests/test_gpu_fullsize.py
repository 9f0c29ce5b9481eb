"""GPU: BASELINE.json-sized cases.  Config 1 (one 10 s clip, base model) is small enough for the CPU
oracle, so it is compared directly; the larger configurations are checked through size-independent
properties: batch invariance, shard-union == whole, determinism, index -> embedding round trip."""
import numpy as np
import pytest
import torch

from audiotokenization_b200 import configs, sharding, synth
from audiotokenization_b200.model import BigCodecModel
from oracle import bigcodec_oracle as oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def base_model():
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    return cfg, enc_sd, dec_sd, BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="fp32")


def test_config1_ten_second_clip_against_oracle(base_model):
    cfg, enc_sd, dec_sd, model = base_model
    x = synth.synth_batch(0, 1, 160000)
    want = oracle.round_trip(enc_sd, dec_sd, cfg, x)
    out = model(x.cuda(), round_trip=True)
    z = model.encoder(x.cuda())
    assert z.shape == (1, 512, 800) and out["indices"].shape == (1, 1, 800)
    assert rel(z, want["z"]) <= 1e-4
    idx = out["indices"].cpu()
    decided = want["margin"] > 1e-5
    assert torch.equal(idx[decided], want["indices"][decided])
    assert float((idx == want["indices"]).float().mean()) >= 0.995
    if torch.equal(idx, want["indices"]):
        assert rel(out["x_rec"], want["x_rec"]) <= 1e-4
    i16 = model.extract_indices(x.pin_memory(), micro_batch=1)
    assert i16.dtype == np.int16 and i16.shape == (1, 800, 1)
    assert np.array_equal(i16[0], oracle.indices_to_int16(idx))


def test_config2_thirty_second_clips_batch_invariance_and_sharding(base_model):
    cfg, _, _, model = base_model
    n, T = 6, 480000
    x = synth.fast_synth_batch(0, n, T).pin_memory()
    whole = model.extract_indices(x, micro_batch=3)
    assert whole.shape == (n, 2400, 1) and whole.min() >= 0 and whole.max() < 8192
    # a clip encoded alone gives bit-identical indices (no cross-item coupling in any kernel)
    alone = model.extract_indices(x[4:5], micro_batch=1)
    assert np.array_equal(alone[0], whole[4])
    # union of rank shards == single-process result, for every world size the bench uses
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            s, e = sharding.shard_range(n, r, world)
            if e > s:
                parts.append(model.extract_indices(x[s:e], micro_batch=2))
        assert np.array_equal(np.concatenate(parts), whole)
    # determinism
    assert np.array_equal(model.extract_indices(x, micro_batch=6), whole)
    assert len(np.unique(whole)) > 100
    # the two-stage front-end schedule (deep stages over several micro-batches at once) changes nothing: no
    # gathering, a hand-off buffer that fills unevenly (2+2 | 2), one that is never full, and the device entry point
    for mb, deep in ((2, 1), (2, 5), (1, 4), (3, 64)):
        assert np.array_equal(model.extract_indices(x, micro_batch=mb, deep_batch=deep), whole), (mb, deep)
    dev = model.indices_device(x.cuda(), micro_batch=2, rnn_batch=4, deep_batch=3).cpu().numpy()
    assert np.array_equal(dev, whole)
    f_direct = model.encoder.front_cl(x[:2].cuda().view(2, T, 1))
    f_staged = model.encoder.front_deep_cl(model.encoder.front_shallow_cl(x[:2].cuda().view(2, T, 1)))
    assert torch.equal(f_direct, f_staged)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_two_stage_front_schedule_is_bit_identical_on_the_tensor_core_path(base_model, precision):
    """Tensor-core modes: the last shallow ResidualUnit writes straight into a slice of the hand-off buffer
    (`out=` of the streamed-weight kernel); gathering must not change a single index."""
    cfg, enc_sd, dec_sd, _ = base_model
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=precision)
    x = synth.fast_synth_batch(0, 5, 160000).pin_memory()
    plain = model.extract_indices(x, micro_batch=2, deep_batch=1)
    for mb, deep in ((2, 4), (1, 64), (2, 3)):
        assert np.array_equal(model.extract_indices(x, micro_batch=mb, deep_batch=deep), plain), (mb, deep)
    assert np.array_equal(model.indices_device(x.cuda(), micro_batch=2, deep_batch=5).cpu().numpy(), plain)


def test_config3_round_trip_batch(base_model):
    cfg, enc_sd, dec_sd, model = base_model
    x = synth.synth_batch(10, 8, 160000).cuda()
    out = model(x, round_trip=True)
    assert out["x_rec"].shape == (8, 1, 160000) and out["indices"].shape == (1, 8, 800)
    assert float(out["x_rec"].abs().max()) <= 1.0
    # indices -> embedding -> decoder reproduces the round trip (the "index -> waveform" entry)
    emb = model.decoder.vq2emb(out["indices"].permute(1, 2, 0))              # [B, T', C] channel-last
    y2 = model.decoder(emb.transpose(1, 2), vq=False)
    assert torch.equal(y2, out["x_rec"])
    # spot-check two items against the oracle
    want = oracle.round_trip(enc_sd, dec_sd, cfg, x[:2].cpu())
    same = (out["indices"][:, :2].cpu() == want["indices"])
    assert float(same.float().mean()) >= 0.995
    if bool(same.all()):
        assert rel(out["x_rec"][:2], want["x_rec"]) <= 1e-4


def test_inference_full_padding_convention(base_model):
    """inference_full.py:712 pads T to the next multiple of 200 (a full extra hop when aligned)."""
    cfg, enc_sd, dec_sd, model = base_model
    x = synth.synth_batch(3, 1, 16000 - 37)
    pad = 200 - (x.shape[2] % 200)
    xp = torch.nn.functional.pad(x, (0, pad))
    y = model.inference(xp.squeeze(1).cuda())
    assert y.shape == (1, xp.shape[2])
    want = oracle.round_trip(enc_sd, dec_sd, cfg, xp)
    if torch.equal(model(xp.cuda())["indices"].cpu(), want["indices"]):
        assert rel(y, want["x_rec"].squeeze(1)) <= 1e-4


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_longform_chunked_equals_whole_recording(precision):
    """BASELINE configs[3] (scaled to 40 s): hop-aligned chunks with the receptive-field halo through the front end,
    one LSTM + VQ pass over the stitched features == the recording encoded in one piece, bit for bit."""
    from audiotokenization_b200 import configs, longform, synth
    from audiotokenization_b200.model import BigCodecModel
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=precision)
    T = 40 * 16000
    x = synth.fast_synth_batch(7, 1, T).cuda()
    whole = model.indices_device(x, micro_batch=1, rnn_batch=1)                    # [1, T', 1] int16
    chunked = model.indices_longform(x[0, 0], chunk_seconds=9.0, micro_batch=2)     # 5 chunks, the last one short
    assert chunked.shape == whole.shape and torch.equal(chunked, whole)
    hop = int(model.encoder.hop_length)
    halo = longform.halo_frames(model.encoder)
    from audiotokenization_b200.vq import precision_scope
    with precision_scope(precision):
        f_whole = model.encoder.front_cl(x.reshape(1, T, 1))
        f_chunk = longform.stitch(longform.chunked_front(model.encoder.front_cl, x[0, 0], hop, 1800, halo, 2))
    assert torch.equal(f_whole, f_chunk)


@pytest.mark.parametrize("clips,seconds", [(1, 10), (4, 30)])
def test_benched_mode_bf16x3_at_baseline_sizes_against_oracle(clips, seconds):
    """The arithmetic mode bench.py times (bf16x3) at BASELINE configs[0] (1 x 10 s) and at the clip length of
    configs[1] (4 x 30 s, through the same `extract_indices` entry point and schedule as the bench): indices
    bit-exact wherever the oracle's top-1/top-2 cosine margin exceeds 1e-5, latents within 1e-3 relative."""
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="bf16x3")
    T = seconds * 16000
    x = synth.fast_synth_batch(100, clips, T)
    want = oracle.encode_to_indices(enc_sd, dec_sd, cfg, x)
    i16 = model.extract_indices(x.pin_memory(), micro_batch=2, rnn_batch=256, deep_batch=64)       # [N, T', 1]
    got = torch.from_numpy(i16[:, :, 0].astype("int64"))
    ref = want["indices"][0]                                                                       # [N, T']
    margin = want["margin"][0] if want["margin"].dim() == 3 else want["margin"]
    decided = margin > 1e-5
    flips = (got != ref)
    assert not bool((flips & decided).any()), ("flipped at margins", margin[flips & decided])
    agree = 1.0 - float(flips.float().mean())
    assert agree >= 0.998, agree
    z = model.encoder(x[:1].cuda())
    e_z = rel(z, want["z"][:1])
    assert e_z <= 1e-3, e_z
    print(f"bf16x3 {clips} x {seconds} s: idx agree {agree:.5f} ({int(flips.sum())} flips, all at margin <= 1e-5), z rel {e_z:.2e}")
