"""Synthetic weights and inputs for parity tests and benchmarks.

There is no network, so no trained checkpoint: weights are random-init with the
reference's parameter names and shapes, but -- unlike
the reference's pristine init, where every Conv1d bias and every Snake
alpha/beta is exactly zero (vq/codec_encoder.py:9-12, vq/activations.py:95-97)
-- biases, alpha, beta and the weight-norm gains are randomised so that a
kernel which ignores one of them fails parity (SURVEY.md section 8c hazard 1).

The generator does not import the reference: the same seeded state dicts are
loaded (strict) into the reference modules by ``scripts/make_golden.py`` and
into this package's modules by the tests / bench, on any machine.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .configs import SAMPLE_RATE


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _randn(gen, shape, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def _wn_conv(sd, gen, prefix, out_ch, in_ch, k, *, transpose=False):
    """Old-style weight_norm triple ``weight_g / weight_v / bias`` (dim=0)."""
    if transpose:  # ConvTranspose1d weight is [in, out, k]; norm is per *input* channel
        shape, gshape, fan_in = (in_ch, out_ch, k), (in_ch, 1, 1), out_ch * k
    else:
        shape, gshape, fan_in = (out_ch, in_ch, k), (out_ch, 1, 1), in_ch * k
    bound = 1.0 / math.sqrt(fan_in)
    v = _uniform(gen, shape, bound)
    norm = v.flatten(1).norm(dim=1).view(gshape)
    g = norm * torch.exp(_randn(gen, gshape, 0.1))
    sd[prefix + "weight_g"] = g
    sd[prefix + "weight_v"] = v
    sd[prefix + "bias"] = _uniform(gen, (out_ch,), bound)


def _wn_linear(sd, gen, prefix, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    v = _uniform(gen, (out_f, in_f), bound)
    g = v.norm(dim=1, keepdim=True) * torch.exp(_randn(gen, (out_f, 1), 0.1))
    sd[prefix + "weight_g"] = g
    sd[prefix + "weight_v"] = v
    sd[prefix + "bias"] = _uniform(gen, (out_f,), bound)


_FIR = None


def kaiser_sinc_filter12() -> torch.Tensor:
    """The 12-tap Kaiser-windowed sinc low-pass used by both FIRs of the
    anti-aliased activation (cutoff 0.25, half-width 0.3), [1,1,12] float32.

    Restates vq/alias_free_torch/filter.py:28-57 for the only parameters the
    codec uses (UpSample1d/DownSample1d with ratio 2, kernel 12:
    resample.py:10-22,36-45).
    """
    global _FIR
    if _FIR is None:
        kernel_size, cutoff, half_width = 12, 0.25, 0.3
        half = kernel_size // 2
        delta_f = 4 * half_width
        att = 2.285 * (half - 1) * math.pi * delta_f + 7.95
        if att > 50.0:
            beta = 0.1102 * (att - 8.7)
        elif att >= 21.0:
            beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
        else:
            beta = 0.0
        window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
        time = torch.arange(-half, half) + 0.5
        f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
        f = f / f.sum()
        _FIR = f.view(1, 1, kernel_size).to(torch.float32)
    return _FIR.clone()


def _act(sd, gen, prefix, ch, antialias):
    sd[prefix + "act.alpha"] = _randn(gen, (ch,), 0.3)
    sd[prefix + "act.beta"] = _randn(gen, (ch,), 0.3)
    if antialias:
        sd[prefix + "upsample.filter"] = kaiser_sinc_filter12()
        sd[prefix + "downsample.lowpass.filter"] = kaiser_sinc_filter12()


def _res_unit(sd, gen, prefix, ch, causal, antialias):
    cpre = "conv." if causal else ""
    _act(sd, gen, prefix + "block.0.", ch, antialias)
    _wn_conv(sd, gen, prefix + "block.1." + cpre, ch, ch, 7)
    _act(sd, gen, prefix + "block.2.", ch, antialias)
    _wn_conv(sd, gen, prefix + "block.3.", ch, ch, 1)  # the 1x1 conv is never causal-wrapped


def _lstm(sd, gen, prefix, ch, layers, bidirectional):
    hid = ch // 2 if bidirectional else ch
    bound = 1.0 / math.sqrt(hid)
    for l in range(layers):
        in_f = ch if l == 0 else hid * (2 if bidirectional else 1)
        for suffix in ([""] + (["_reverse"] if bidirectional else [])):
            sd[f"{prefix}lstm.weight_ih_l{l}{suffix}"] = _uniform(gen, (4 * hid, in_f), bound)
            sd[f"{prefix}lstm.weight_hh_l{l}{suffix}"] = _uniform(gen, (4 * hid, hid), bound)
            sd[f"{prefix}lstm.bias_ih_l{l}{suffix}"] = _uniform(gen, (4 * hid,), bound)
            sd[f"{prefix}lstm.bias_hh_l{l}{suffix}"] = _uniform(gen, (4 * hid,), bound)


def make_encoder_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the key names of ``BigCodecEncoder`` (vq/codec_encoder.py:35-57)."""
    gen = torch.Generator().manual_seed(1000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    causal, aa = bool(cfg.get("causal", False)), bool(cfg.get("antialias", False))
    cpre = "conv." if causal else ""
    d = cfg["ngf"]
    _wn_conv(sd, gen, "block.0." + cpre, d, 1, 7)
    idx = 1
    for stride in cfg["up_ratios"]:
        d *= 2
        half = d // 2
        for r, _ in enumerate(cfg["dilations"]):
            _res_unit(sd, gen, f"block.{idx}.block.{r}.", half, causal, aa)
        nd = len(cfg["dilations"])
        _act(sd, gen, f"block.{idx}.block.{nd}.", half, aa)
        k = 2 * stride if stride != 1 else 1
        _wn_conv(sd, gen, f"block.{idx}.block.{nd + 1}." + cpre, d, half, k)
        idx += 1
    if cfg.get("use_rnn", True):
        _lstm(sd, gen, f"block.{idx}.", d, cfg.get("rnn_num_layers", 2), cfg.get("rnn_bidirectional", False))
        idx += 1
    _act(sd, gen, f"block.{idx}.", d, aa)
    _wn_conv(sd, gen, f"block.{idx + 1}." + cpre, cfg["out_channels"], d, 3)
    return sd


def make_decoder_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the key names of ``BigCodecDecoder`` (vq/codec_decoder.py:48-81)."""
    gen = torch.Generator().manual_seed(2000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    causal, aa = bool(cfg.get("causal", False)), bool(cfg.get("antialias", False))
    cpre = "conv." if causal else ""
    cin, cdim = cfg["in_channels"], cfg["codebook_dim"]
    if cfg.get("fsq", False):   # FSQ(levels, dim=in_channels): plain nn.Linear projections, no persistent buffers
        d = len(cfg["fsq_levels"])
        # project_in scaled so that the projected latents spread over several quantisation levels
        sd["quantizer.project_in.weight"] = _uniform(gen, (d, cin), 4.0 / math.sqrt(cin))
        sd["quantizer.project_in.bias"] = _uniform(gen, (d,), 0.3)
        sd["quantizer.project_out.weight"] = _uniform(gen, (cin, d), 1.0 / math.sqrt(d))
        sd["quantizer.project_out.bias"] = _uniform(gen, (cin,), 1.0 / math.sqrt(d))
    for q in range(0 if cfg.get("fsq", False) else cfg.get("vq_num_quantizers", 1)):
        p = f"quantizer.layers.{q}."
        if cin != cdim:
            _wn_linear(sd, gen, p + "in_proj.", cdim, cin)
            _wn_linear(sd, gen, p + "out_proj.", cin, cdim)
        sd[p + "_codebook.weight"] = _randn(gen, (cfg["codebook_size"], cdim), 1.0)
    ch = cfg["upsample_initial_channel"]
    _wn_conv(sd, gen, "model.0." + cpre, ch, cin, 7)
    idx = 1
    if cfg.get("use_rnn", True):
        _lstm(sd, gen, f"model.{idx}.", ch, cfg.get("rnn_num_layers", 2), cfg.get("rnn_bidirectional", False))
        idx += 1
    out_dim = ch
    for i, stride in enumerate(cfg["up_ratios"]):
        in_dim, out_dim = ch // 2 ** i, ch // 2 ** (i + 1)
        _act(sd, gen, f"model.{idx}.block.0.", in_dim, aa)
        k = 2 * stride if stride != 1 else 1
        _wn_conv(sd, gen, f"model.{idx}.block.1." + cpre, out_dim, in_dim, k, transpose=True)
        for r, _ in enumerate(cfg["dilations"]):
            _res_unit(sd, gen, f"model.{idx}.block.{2 + r}.", out_dim, causal, aa)
        idx += 1
    _act(sd, gen, f"model.{idx}.", out_dim, aa)
    _wn_conv(sd, gen, f"model.{idx + 1}." + cpre, 1, out_dim, 7)
    return sd


def make_state_dicts(cfg: dict, seed: int = 0) -> Tuple[dict, dict]:
    return (make_encoder_state_dict(cfg["codec_encoder"], seed),
            make_decoder_state_dict(cfg["codec_decoder"], seed))


def synth_clip(index: int, num_samples: int, kind: str = "tones") -> torch.Tensor:
    """One synthetic 16 kHz mono clip in [-1, 1].

    ``tones``: 0.5 * sum of 8 random sinusoids (80-4000 Hz) + 0.3 * white noise,
    clamped; ``noise``: unit white noise clamped.  Seeded per clip index, so any
    rank can regenerate exactly its own shard.
    """
    gen = torch.Generator().manual_seed(1234 + index)
    if kind == "noise":
        return torch.randn(num_samples, generator=gen).clamp_(-1.0, 1.0)
    if kind != "tones":
        raise ValueError(kind)
    t = torch.arange(num_samples, dtype=torch.float64) / SAMPLE_RATE
    f = 80.0 + (4000.0 - 80.0) * torch.rand(8, generator=gen, dtype=torch.float64)
    a = torch.rand(8, generator=gen, dtype=torch.float64)
    ph = 2 * math.pi * torch.rand(8, generator=gen, dtype=torch.float64)
    x = (a[:, None] * torch.sin(2 * math.pi * f[:, None] * t[None, :] + ph[:, None])).sum(0)
    x = 0.5 * x.to(torch.float32) + 0.3 * torch.randn(num_samples, generator=gen)
    return x.clamp_(-1.0, 1.0)


def synth_batch(first_index: int, count: int, num_samples: int, kind: str = "tones") -> torch.Tensor:
    """``[count, 1, num_samples]`` float32 batch of clips ``first_index ..``."""
    return torch.stack([synth_clip(first_index + i, num_samples, kind) for i in range(count)]).unsqueeze(1)


def fast_synth_batch(first_index: int, count: int, num_samples: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Cheap bulk generator for benchmark-sized workloads (thousands of 30 s
    clips): one base ``tones`` clip per 8 indices, deterministically scaled and
    circularly shifted per clip.  Values stay in [-1, 1]."""
    if out is None:
        out = torch.empty(count, 1, num_samples, dtype=torch.float32)
    base = {}
    for i in range(count):
        idx = first_index + i
        b = idx // 8
        if b not in base:
            base[b] = synth_clip(100000 + b, num_samples)
        shift = (idx % 8) * 977 + 1
        scale = 1.0 - 0.05 * (idx % 8)
        out[i, 0] = torch.roll(base[b], shift) * scale
    return out
