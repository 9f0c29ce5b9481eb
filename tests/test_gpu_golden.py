"""GPU: end-to-end parity against fixtures generated from the LIVE reference
(scripts/make_golden.py): encoder latents, VQ indices, quantised latents, decoded waveform.

Float32 mode contract (BASELINE.json north_star): latents / waveforms within 1e-3 relative,
indices bit-exact wherever the reference's top-1/top-2 cosine margin exceeds 1e-5.  The
float32 CUDA-core path is asserted at a 10x tighter 1e-4."""
import numpy as np
import pytest
import torch

from audiotokenization_b200 import configs, synth
from audiotokenization_b200.model import BigCodecModel
from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_round_trip_matches_reference_fixture(case):
    g = load_golden(case)
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="fp32")
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"]).cuda()

    z = model.encoder(x)
    assert tuple(z.shape) == g["z_f32"].shape
    e_z32, e_z64 = rel(z.cpu().numpy(), g["z_f32"]), rel(z.cpu().numpy(), g["z_f64"])
    assert e_z32 <= 1e-4 and e_z64 <= 1e-4, (e_z32, e_z64)

    z_q, idx, loss = model.decoder(z, vq=True)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == g["idx_f32"].shape
    idx_np = idx.cpu().numpy()
    decided = g["margin_f64"][None] > 1e-5
    assert np.array_equal(idx_np[decided], g["idx_f64"][decided])
    agree = float((idx_np == g["idx_f32"]).mean())
    assert agree >= 0.99, agree
    same = (idx_np == g["idx_f32"])[0]                                   # [B, T']
    zq_np = z_q.cpu().numpy().transpose(0, 2, 1)
    assert rel(zq_np[same], g["zq_f32"].transpose(0, 2, 1)[same]) <= 1e-5

    # decode the REFERENCE's quantised latents so the waveform check is independent of index flips
    y = model.decoder(torch.from_numpy(g["zq_f32"]).cuda(), vq=False)
    assert tuple(y.shape) == g["y_f32"].shape
    e_y = rel(y.cpu().numpy(), g["y_f32"])
    assert e_y <= 1e-4, e_y

    out = model(x, round_trip=True)
    assert set(out) == {"x_rec", "indices", "loss"} and out["loss"] == {}
    assert np.array_equal(out["indices"].cpu().numpy(), idx_np)
    if agree == 1.0:
        assert rel(out["x_rec"].cpu().numpy(), g["y_f32"]) <= 1e-4
    emb = model.decoder.vq2emb(idx.permute(1, 2, 0))
    assert rel(emb.cpu().numpy()[same], g["emb_f32"][same]) <= 1e-5
    print(f"{case}: z rel {e_z32:.2e} (vs f64 {e_z64:.2e}), y rel {e_y:.2e}, idx agree {agree:.4f}")


@pytest.mark.parametrize("case", ["base_1s", "debug_1s", "config9_base_1s", "debug_causal_1s", "tiny", "base_aa_1s", "default_half_s"])
def test_tensor_core_split_mode_meets_the_fp32_contract(case):
    """bf16x3 (tcgen05, hi/lo split operands, fused ResidualUnits, tensor-core LSTM) -- the mode bench.py times --
    against the reference fixtures at the contract's own thresholds: indices bit-exact wherever the top-1/top-2
    cosine margin exceeds 1e-5, latents and waveforms within 1e-3 relative (asserted 5x tighter: 2e-4).  The
    sub-modules are called directly, like the reference's ``lm.model['CodecEnc'](x)``: they run in the model's mode."""
    g = load_golden(case)
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="bf16x3")
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"]).cuda()
    out = model(x, round_trip=True)
    from audiotokenization_b200 import ops
    n0 = ops.STATS["launches"]
    prof, ops.PROFILE = ops.PROFILE, []
    try:
        z = model.encoder(x)                              # direct sub-module call: must be the tensor-core path
        kernels = {rec[4] for rec in ops.PROFILE}
    finally:
        ops.PROFILE = prof
    assert ops.STATS["launches"] > n0
    if cfg["codec_encoder"]["ngf"] >= 16 and not g["antialias"]:
        assert kernels & {"conv_stream_kernel", "ru_persist_kernel", "ru_group_kernel", "ru_pair_kernel", "conv1d_tc_kernel"}, kernels
    e_z = rel(z.cpu().numpy(), g["z_f64"])
    assert e_z <= 2e-4, e_z
    idx = out["indices"].cpu().numpy()
    decided = g["margin_f64"][None] > 1e-5
    assert np.array_equal(idx[decided], g["idx_f64"][decided]), (
        "flipped frames at margins", g["margin_f64"][None][decided & (idx != g["idx_f64"])])
    agree = float((idx == g["idx_f32"]).mean())
    assert agree >= 0.99, agree
    y = model.decoder(torch.from_numpy(g["zq_f32"]).cuda(), vq=False)
    e_y = rel(y.cpu().numpy(), g["y_f64"])
    assert e_y <= 2e-4, e_y
    print(f"{case} bf16x3: z rel {e_z:.2e}, y rel {e_y:.2e}, idx agree {agree:.4f}")


@pytest.mark.parametrize("case", ["base_1s", "debug_1s"])
def test_single_pass_bf16_fast_mode_is_within_its_documented_envelope(case):
    """bf16 single pass: ~1e-2 latent error and >= 90 % raw index agreement (SURVEY.md section 7 measured
    9.3e-3 / 96.9 % for bf16-rounded conv operands on the reference itself)."""
    g = load_golden(case)
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision="bf16")
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"]).cuda()
    out = model(x)
    z = model.encoder(x)
    e_z = rel(z.cpu().numpy(), g["z_f64"])
    agree = float((out["indices"].cpu().numpy() == g["idx_f32"]).mean())
    assert e_z <= 5e-2, e_z
    assert agree >= 0.85, agree
    print(f"{case} bf16: z rel {e_z:.2e}, idx agree {agree:.4f}")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_fsq_decoder_branch_round_trip_matches_reference_fixture(precision):
    """BigCodecDecoder(fsq=True) (vq/codec_decoder.py:41-47,87-89) end to end against the live-reference fixture:
    int32 indices [B, T'] exact away from rounding boundaries, (x, q, zeros[B]) return convention, waveform, the
    driver-level index path (int16 (T', 1) arrays)."""
    g = load_golden("tiny_fsq")
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=precision)
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"]).cuda()
    z = model.encoder(x)
    assert rel(z.cpu().numpy(), g["z_f64"]) <= 2e-4
    z_q, idx, loss = model.decoder(z, vq=True)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == g["idx_f32"].shape and tuple(loss.shape) == (g["batch"],)
    assert float(loss.abs().sum()) == 0.0
    decided = g["margin_f64"] > (1e-5 if precision == "fp32" else 2e-3)     # boundary distance of the bounded latents
    assert np.array_equal(idx.cpu().numpy()[decided], g["idx_f64"][decided])
    same = idx.cpu().numpy() == g["idx_f32"]
    assert same.mean() >= 0.98
    assert rel(z_q.cpu().numpy().transpose(0, 2, 1)[same], g["zq_f32"].transpose(0, 2, 1)[same]) <= 1e-5
    y = model.decoder(torch.from_numpy(g["zq_f32"]).cuda(), vq=False)
    assert rel(y.cpu().numpy(), g["y_f32"]) <= 2e-4
    out = model(x, round_trip=True)
    assert np.array_equal(out["indices"].cpu().numpy(), idx.cpu().numpy())
    emb = model.decoder.quantizer.indices_to_codes(idx)
    assert rel(emb.cpu().numpy().transpose(0, 2, 1)[same], g["emb_f32"][same]) <= 1e-5
    i16 = model.extract_indices(x.cpu().pin_memory(), micro_batch=2)
    assert i16.shape == (g["batch"], g["idx_f32"].shape[1], 1) and np.array_equal(i16[:, :, 0], idx.cpu().numpy().astype(np.int16))
