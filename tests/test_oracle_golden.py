"""CPU: the oracle restatement vs. fixtures produced by the live reference
(scripts/make_golden.py).  Bit-exact indices; latents / waveforms to float32
rounding (the reference and the oracle call the same torch CPU kernels)."""
import numpy as np
import pytest
import torch

from audiotokenization_b200 import configs, synth
from oracle import bigcodec_oracle as oracle
from conftest import GOLDEN_CASES, load_golden


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_fixture(case):
    g = load_golden(case)
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"])
    out = oracle.round_trip(enc_sd, dec_sd, cfg, x)
    assert tuple(out["z"].shape) == g["z_f32"].shape
    assert out["z"].shape[2] == oracle.encoder_output_length(cfg["codec_encoder"], g["num_samples"])
    assert rel(out["z"].numpy(), g["z_f32"]) <= 2e-6
    assert np.array_equal(out["indices"].numpy().astype(np.int32), g["idx_f32"])
    assert rel(out["z_q"].numpy(), g["zq_f32"]) <= 2e-6
    assert rel(out["x_rec"].numpy(), g["y_f32"]) <= 2e-6
    assert np.abs(out["margin"][0].numpy() - g["margin_f32"]).max() <= 2e-6
    emb = oracle.vq2emb(dec_sd, cfg["codec_decoder"], out["indices"].permute(1, 2, 0))
    assert rel(emb.numpy(), g["emb_f32"]) <= 2e-6
    assert rel(emb.transpose(1, 2).numpy(), out["z_q"].numpy()) <= 2e-6


@pytest.mark.parametrize("case", ["tiny", "tiny_aa", "debug_causal_1s"])
def test_oracle_float64_matches_reference_float64(case):
    g = load_golden(case)
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    enc_sd, dec_sd = oracle.cast_sd(enc_sd, torch.float64), oracle.cast_sd(dec_sd, torch.float64)
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"]).double()
    out = oracle.round_trip(enc_sd, dec_sd, cfg, x)
    assert rel(out["z"].numpy(), g["z_f64"]) <= 1e-6      # fixture stored as f32
    assert np.array_equal(out["indices"].numpy().astype(np.int32), g["idx_f64"])
    assert rel(out["x_rec"].numpy(), g["y_f64"]) <= 1e-6


def test_lstm_loop_matches_library_lstm():
    cfg = configs.get_config("tiny")
    enc_sd, _ = synth.make_state_dicts(cfg, seed=3)
    x = torch.randn(2, 64, 37, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    sd = oracle.cast_sd(enc_sd, torch.float64)
    a = oracle.res_lstm(sd, "block.4.", x, 2)
    b = oracle.res_lstm_loop(sd, "block.4.", x, 2)
    assert rel(a.numpy(), b.numpy()) < 1e-12


def test_filter_closed_form():
    f = oracle.kaiser_sinc_filter12(torch.float64).flatten().numpy()
    want = [0.0020290, 0.0093895, -0.0255435, -0.0576574, 0.1285726, 0.4432098]
    assert np.allclose(f[:6], want, atol=5e-7) and np.allclose(f[6:], want[::-1], atol=5e-7)
    assert abs(f.sum() - 1) < 1e-6
    assert np.array_equal(synth.kaiser_sinc_filter12().flatten().numpy(),
                          oracle.kaiser_sinc_filter12().flatten().numpy())


def test_int16_disk_form():
    idx = torch.arange(7).view(1, 1, 7)
    arr = oracle.indices_to_int16(idx)
    assert arr.dtype == np.int16 and arr.shape == (7, 1)


def test_oracle_fsq_branch_matches_reference_fixture():
    """BigCodecDecoder(fsq=True): the oracle's FSQ restatement against the live reference (fixture tiny_fsq): int32
    indices of shape [B, T'], bit-exact; quantised latents / waveform to float32 rounding; boundary distances."""
    g = load_golden("tiny_fsq")
    cfg = configs.get_config(g["cfg_name"], antialias=g["antialias"])
    assert cfg["codec_decoder"]["fsq"] is True
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=g["seed"])
    x = synth.synth_batch(0, g["batch"], g["num_samples"], g["kind"])
    out = oracle.round_trip(enc_sd, dec_sd, cfg, x)
    assert out["indices"].dtype == torch.int32 and tuple(out["indices"].shape) == g["idx_f32"].shape == (2, 100)
    assert np.array_equal(out["indices"].numpy(), g["idx_f32"])
    assert rel(out["z_q"].numpy(), g["zq_f32"]) <= 2e-6 and rel(out["x_rec"].numpy(), g["y_f32"]) <= 2e-6
    assert np.abs(out["margin"].numpy() - g["margin_f32"]).max() <= 1e-5      # bounded latents reach |3.5|: float32 rounding of z
    # every level combination decodes back through the implicit codebook (random latents reach many more codes)
    z = torch.randn(3, 64, 500, generator=torch.Generator().manual_seed(4)) * 3
    q, idx, _, _ = oracle.quantize(dec_sd, cfg["codec_decoder"], z)
    assert len(torch.unique(idx)) > 150 and int(idx.min()) >= 0 and int(idx.max()) < 512
