"""CPU oracle for the BigCodec hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product path
(``audiotokenization_b200``) never imports it and has no CPU fallback.

What it is: a functional restatement, on CPU tensors, of the arithmetic of the
reference's hot path (paths relative to ``/root/reference/BigCodec_SSL/``):

  weight-norm fold          torch.nn.utils.weight_norm(dim=0) as applied at
                            vq/module.py:59-72, vq/factorized_vector_quantize.py:18-19
  SnakeBeta                 vq/activations.py:107-119
  anti-aliased Activation1d vq/alias_free_torch/act.py:25-32, resample.py:25-33,
                            filter.py:86-95 (filter taps: filter.py:28-57)
  Conv / causal conv        vq/module.py:11-57
  ResidualUnit / blocks     vq/module.py:74-141
  ResLSTM                   vq/module.py:143-167
  encoder                   vq/codec_encoder.py:35-64
  factorized VQ             vq/factorized_vector_quantize.py:29-109, vq/residual_vq.py:21-53
  decoder                   vq/codec_decoder.py:59-94
  driver forward            extract_indices.py:347-371, inference_full.py:557-561

The reference's arithmetic *is* PyTorch CPU library calls (F.conv1d,
F.conv_transpose1d, nn.LSTM, F.normalize, matmul); PyTorch is the reference's
third-party numeric dependency (requirements.txt pins no version; here torch
2.11.0).  This file calls the same library entry points functionally on a
plain state dict, with no nn.Module tree, so it is usable in float32 (the
reference's precision) and float64 (error attribution).  The functions follow
the device of their inputs: ``bench.py``'s baseline leg also runs them on CUDA
tensors, which is exactly what the reference does on a GPU box
(extract_indices.py:399: ``model.to('cuda')`` -> cuDNN / cuBLAS eager) -- the
"library Blackwell path" of SURVEY.md section 2.2, a second reported baseline.

Pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against the *live reference modules* executed in the
build container: ``scripts/make_golden.py`` imports ``/root/reference``'s
``vq`` package, loads the seeded state dicts of ``audiotokenization_b200.synth``
with ``strict=True``, runs encoder -> VQ -> decoder and commits the outputs as
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
oracle against those fixtures on every run (bit-exact indices, <=2e-6 relative
on latents/waveforms).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------
def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v||, norm over all dims but 0 (weight_norm default dim=0).

    For Conv1d v is [out,in,k] (per-output-channel norm); for ConvTranspose1d v
    is [in,out,k] so the norm is per *input* channel; for Linear v is [out,in].
    """
    norm = v.flatten(1).norm(dim=1).view([-1] + [1] * (v.dim() - 1))
    return v * (g / norm)


def _wn(sd: SD, prefix: str, dtype) -> Tuple[Tensor, Tensor]:
    w = fold_weight_norm(sd[prefix + "weight_g"].to(dtype), sd[prefix + "weight_v"].to(dtype))
    return w, sd[prefix + "bias"].to(dtype)


def snake_beta(x: Tensor, alpha: Tensor, beta: Tensor) -> Tensor:
    """SnakeBeta with alpha_logscale=True: x + sin^2(x e^alpha) / (e^beta + 1e-9)."""
    a = torch.exp(alpha).view(1, -1, 1)
    b = torch.exp(beta).view(1, -1, 1)
    return x + (1.0 / (b + 0.000000001)) * torch.sin(x * a).pow(2)


def kaiser_sinc_filter12(dtype=torch.float32) -> Tensor:
    """12-tap Kaiser-sinc low-pass (cutoff .25, half-width .3); filter.py:28-57."""
    kernel_size, cutoff, half_width = 12, 0.25, 0.3
    half = kernel_size // 2
    att = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
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
    return f.view(1, 1, kernel_size).to(dtype)


def upsample2(x: Tensor, filt: Tensor) -> Tensor:
    """UpSample1d(ratio=2, kernel=12).forward; resample.py:25-33."""
    c = x.shape[1]
    ratio, k = 2, 12
    pad = k // ratio - 1                                  # 5
    pad_left = pad * ratio + (k - ratio) // 2             # 15
    pad_right = pad * ratio + (k - ratio + 1) // 2        # 15
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.expand(c, -1, -1), stride=ratio, groups=c)
    return x[..., pad_left:-pad_right]


def downsample2(x: Tensor, filt: Tensor) -> Tensor:
    """DownSample1d(ratio=2, kernel=12) = LowPassFilter1d(stride=2); filter.py:86-95."""
    c = x.shape[1]
    x = F.pad(x, (5, 6), mode="replicate")                # pad_left = 12//2-1, pad_right = 12//2
    return F.conv1d(x, filt.expand(c, -1, -1), stride=2, groups=c)


def activation1d(sd: SD, prefix: str, x: Tensor, antialias: bool) -> Tensor:
    """Activation1d(SnakeBeta(alpha_logscale=True), antialias); act.py:25-32."""
    alpha = sd[prefix + "act.alpha"].to(x.dtype)
    beta = sd[prefix + "act.beta"].to(x.dtype)
    if not antialias:
        return snake_beta(x, alpha, beta)
    fu = sd.get(prefix + "upsample.filter")
    fd = sd.get(prefix + "downsample.lowpass.filter")
    fu = (kaiser_sinc_filter12(x.dtype) if fu is None else fu).to(device=x.device, dtype=x.dtype)
    fd = (kaiser_sinc_filter12(x.dtype) if fd is None else fd).to(device=x.device, dtype=x.dtype)
    return downsample2(snake_beta(upsample2(x, fu), alpha, beta), fd)


def wn_conv1d(sd: SD, prefix: str, x: Tensor, *, stride=1, dilation=1, padding=0, causal=False) -> Tensor:
    """WNConv1d (vq/module.py:59-65); causal => CausalConv1d (vq/module.py:11-48)."""
    if causal:
        w, b = _wn(sd, prefix + "conv.", x.dtype)
        k = w.shape[-1]
        x = F.pad(x, ((k - stride) * dilation, 0))
        return F.conv1d(x, w, b, stride=stride, dilation=dilation)
    w, b = _wn(sd, prefix, x.dtype)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, padding=padding)


def wn_conv_transpose1d(sd: SD, prefix: str, x: Tensor, *, stride: int, causal=False) -> Tensor:
    """WNConvTranspose1d as configured by DecoderBlock (vq/module.py:67-72,119-136)."""
    if causal:
        w, b = _wn(sd, prefix + "conv.", x.dtype)
        return F.conv_transpose1d(x, w, b, stride=stride)[..., :-stride]
    w, b = _wn(sd, prefix, x.dtype)
    pad = stride // 2 + stride % 2 if stride != 1 else 0
    out_pad = stride % 2 if stride != 1 else 0
    return F.conv_transpose1d(x, w, b, stride=stride, padding=pad, output_padding=out_pad)


def residual_unit(sd: SD, prefix: str, x: Tensor, dilation: int, causal: bool, antialias: bool) -> Tensor:
    """ResidualUnit (vq/module.py:74-89)."""
    h = activation1d(sd, prefix + "block.0.", x, antialias)
    h = wn_conv1d(sd, prefix + "block.1.", h, dilation=dilation, padding=((7 - 1) * dilation) // 2, causal=causal)
    h = activation1d(sd, prefix + "block.2.", h, antialias)
    h = wn_conv1d(sd, prefix + "block.3.", h)
    return x + h


def res_lstm(sd: SD, prefix: str, x: Tensor, num_layers: int, bidirectional: bool = False) -> Tensor:
    """ResLSTM (vq/module.py:143-167): y = LSTM(x^T) + x^T, zero initial state."""
    xt = x.transpose(1, 2)
    dt = x.dtype
    suffixes = [""] + (["_reverse"] if bidirectional else [])
    flat: List[Tensor] = []
    for l in range(num_layers):
        for s in suffixes:
            for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                flat.append(sd[f"{prefix}lstm.{name}_l{l}{s}"].to(dt))
    hid = flat[1].shape[1]
    ndir = 2 if bidirectional else 1
    h0 = torch.zeros(num_layers * ndir, xt.shape[0], hid, dtype=dt, device=xt.device)
    y, _, _ = torch.lstm(xt.contiguous(), (h0, h0.clone()), flat, True, num_layers, 0.0, False, bidirectional, True)
    return (y + xt).transpose(1, 2)


def res_lstm_loop(sd: SD, prefix: str, x: Tensor, num_layers: int) -> Tensor:
    """Explicit-loop uni-directional LSTM (gate order i,f,g,o); cross-check of ``res_lstm``."""
    xt = x.transpose(1, 2)
    inp = xt
    for l in range(num_layers):
        w_ih = sd[f"{prefix}lstm.weight_ih_l{l}"].to(x.dtype)
        w_hh = sd[f"{prefix}lstm.weight_hh_l{l}"].to(x.dtype)
        b = (sd[f"{prefix}lstm.bias_ih_l{l}"] + sd[f"{prefix}lstm.bias_hh_l{l}"]).to(x.dtype)
        hid = w_hh.shape[1]
        h = torch.zeros(x.shape[0], hid, dtype=x.dtype, device=x.device)
        c = torch.zeros_like(h)
        outs = []
        pre = inp @ w_ih.t() + b
        for t in range(inp.shape[1]):
            g = pre[:, t] + h @ w_hh.t()
            i, f, gg, o = g.split(hid, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
    return (inp + xt).transpose(1, 2)


# ----------------------------------------------------------------------------
# encoder / VQ / decoder
# ----------------------------------------------------------------------------
def encoder_forward(sd: SD, cfg: dict, x: Tensor) -> Tensor:
    """BigCodecEncoder.forward (vq/codec_encoder.py:35-64): [B,1,T] -> [B,out,T']."""
    causal, aa = bool(cfg.get("causal", False)), bool(cfg.get("antialias", False))
    h = wn_conv1d(sd, "block.0.", x, padding=3, causal=causal)
    idx = 1
    nd = len(cfg["dilations"])
    for stride in cfg["up_ratios"]:
        p = f"block.{idx}."
        for r, d in enumerate(cfg["dilations"]):
            h = residual_unit(sd, f"{p}block.{r}.", h, d, causal, aa)
        h = activation1d(sd, f"{p}block.{nd}.", h, aa)
        pad = stride // 2 + stride % 2 if stride != 1 else 0
        h = wn_conv1d(sd, f"{p}block.{nd + 1}.", h, stride=stride, padding=pad, causal=causal)
        idx += 1
    if cfg.get("use_rnn", True):
        h = res_lstm(sd, f"block.{idx}.", h, cfg.get("rnn_num_layers", 2), cfg.get("rnn_bidirectional", False))
        idx += 1
    h = activation1d(sd, f"block.{idx}.", h, aa)
    return wn_conv1d(sd, f"block.{idx + 1}.", h, padding=1, causal=causal)


def encoder_output_length(cfg: dict, t: int) -> int:
    """Chain of floors of the strided convs (SURVEY.md section 7 'edge semantics')."""
    causal = bool(cfg.get("causal", False))
    for s in cfg["up_ratios"]:
        if s == 1:
            continue
        k = 2 * s
        if causal:
            t = (t + (k - s) - k) // s + 1
        else:
            p = s // 2 + s % 2
            t = (t + 2 * p - k) // s + 1
    return t


def vq_layer_forward(sd: SD, prefix: str, z: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """FactorizedVectorQuantize.forward in eval mode (factorized_vector_quantize.py:29-76,93-109).

    Returns (z_q [B,C,T], indices int64 [B,T], commit_loss [B] (zeros), margin [B,T])
    where margin is the top-1 minus top-2 value of the reference's own score
    ``-dist`` divided by 2 (i.e. the cosine margin, since dist = 2 - 2 cos).
    """
    dt = z.dtype
    zt = z.transpose(1, 2)                                              # b t d
    has_proj = (prefix + "in_proj.weight_v") in sd
    if has_proj:
        w_in, b_in = _wn(sd, prefix + "in_proj.", dt)
        z_e = F.linear(zt, w_in, b_in)
    else:
        z_e = zt
    cb = sd[prefix + "_codebook.weight"].to(dt)
    enc = F.normalize(z_e.reshape(-1, z_e.shape[-1]))
    cbn = F.normalize(cb)
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cbn.t() + cbn.pow(2).sum(1, keepdim=True).t()
    score = -dist
    top2 = score.topk(2, dim=1).values
    idx = score.max(1)[1]
    margin = (top2[:, 0] - top2[:, 1]) * 0.5
    idx = idx.view(z.shape[0], -1)
    z_q = F.embedding(idx, cb)                                          # b t d  (raw codebook rows)
    if has_proj:
        w_out, b_out = _wn(sd, prefix + "out_proj.", dt)
        z_q = F.linear(z_q, w_out, b_out)
    z_q = z_q.transpose(1, 2)
    return z_q, idx, torch.zeros(z.shape[0], dtype=dt, device=z.device), margin.view(z.shape[0], -1)


def fsq_forward(sd: SD, prefix: str, z: Tensor, levels) -> Tuple[Tensor, Tensor, Tensor]:
    """FSQ.forward for the decoder's configuration -- FSQ(levels, channel_first=True, dim=C), one codebook, eval
    (finite_scalar_quantization.py:111-116 bound, :142-148 quantize, :170-175 codes_to_indices, :203-259 forward).

    Returns (out [B,C,T], indices int32 [B,T], boundary [B,T]) where ``boundary`` is the distance of the closest bounded
    component to a rounding boundary (0.5 = exactly between two boundaries): the analogue of the VQ cosine margin."""
    dt = z.dtype
    lv = torch.tensor(list(levels), dtype=torch.int32, device=z.device)
    basis = torch.cumprod(torch.tensor([1] + list(levels)[:-1], device=z.device), dim=0, dtype=torch.int32)
    zt = z.transpose(1, 2)                                              # b n d
    has_proj = (prefix + "project_in.weight") in sd
    if has_proj:
        zt = F.linear(zt, sd[prefix + "project_in.weight"].to(dt), sd[prefix + "project_in.bias"].to(dt))
    eps = 1e-3
    half_l = (lv - 1) * (1 + eps) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    bounded = (zt + shift).tanh() * half_l - offset
    quantized = bounded.round()
    half_width = lv // 2
    codes = quantized / half_width
    zhat = (codes * half_width) + half_width
    indices = (zhat * basis).sum(dim=-1).to(torch.int32)
    boundary = (0.5 - (bounded - quantized).abs()).amin(dim=-1)
    out = codes.to(dt)
    if has_proj:
        out = F.linear(out, sd[prefix + "project_out.weight"].to(dt), sd[prefix + "project_out.bias"].to(dt))
    return out.transpose(1, 2), indices, boundary


def quantize(sd: SD, cfg: dict, z: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """BigCodecDecoder.forward(x, vq=True) -> ResidualVQ.forward (residual_vq.py:21-40), or, with ``fsq``, FSQ.forward
    (vq/codec_decoder.py:87-89).

    Returns (z_q [B,C,T'], indices int64 [n_q,B,T'], loss [n_q], margin [n_q,B,T']); with ``fsq``: indices int32
    [B,T'] (the reference's shape), loss [B] zeros, margin = boundary distance [B,T'].
    """
    if cfg.get("fsq", False):
        out, idx, boundary = fsq_forward(sd, "quantizer.", z, cfg["fsq_levels"])
        return out, idx, torch.zeros(z.shape[0], dtype=z.dtype, device=z.device), boundary
    out = torch.zeros_like(z)
    residual = z
    all_idx, all_loss, all_margin = [], [], []
    for q in range(cfg.get("vq_num_quantizers", 1)):
        zq, idx, loss, margin = vq_layer_forward(sd, f"quantizer.layers.{q}.", residual)
        residual = residual - zq
        out = out + zq
        all_idx.append(idx)
        all_loss.append(loss.mean())
        all_margin.append(margin)
    return out, torch.stack(all_idx), torch.stack(all_loss), torch.stack(all_margin)


def vq2emb(sd: SD, cfg: dict, codes: Tensor, proj: bool = True) -> Tensor:
    """ResidualVQ.vq2emb (residual_vq.py:42-48): codes [B,T,n_q] -> [B,T,C] (channel-last!)."""
    out = 0.0
    for q in range(cfg.get("vq_num_quantizers", 1)):
        p = f"quantizer.layers.{q}."
        emb = F.embedding(codes[:, :, q], sd[p + "_codebook.weight"])
        if proj and (p + "out_proj.weight_v") in sd:
            w_out, b_out = _wn(sd, p + "out_proj.", emb.dtype)
            emb = F.linear(emb, w_out, b_out)
        out = out + emb
    return out


def decoder_forward(sd: SD, cfg: dict, x: Tensor) -> Tensor:
    """BigCodecDecoder.forward(x, vq=False) (vq/codec_decoder.py:59-81,93): [B,C,T'] -> [B,1,T]."""
    causal, aa = bool(cfg.get("causal", False)), bool(cfg.get("antialias", False))
    h = wn_conv1d(sd, "model.0.", x, padding=3, causal=causal)
    idx = 1
    if cfg.get("use_rnn", True):
        h = res_lstm(sd, f"model.{idx}.", h, cfg.get("rnn_num_layers", 2), cfg.get("rnn_bidirectional", False))
        idx += 1
    for stride in cfg["up_ratios"]:
        p = f"model.{idx}."
        h = activation1d(sd, p + "block.0.", h, aa)
        h = wn_conv_transpose1d(sd, p + "block.1.", h, stride=stride, causal=causal) if stride != 1 else \
            wn_conv1d_as_transpose_k1(sd, p + "block.1.", h, causal)
        for r, d in enumerate(cfg["dilations"]):
            h = residual_unit(sd, f"{p}block.{2 + r}.", h, d, causal, aa)
        idx += 1
    h = activation1d(sd, f"model.{idx}.", h, aa)
    h = wn_conv1d(sd, f"model.{idx + 1}.", h, padding=3, causal=causal)
    return torch.tanh(h)


def wn_conv1d_as_transpose_k1(sd: SD, prefix: str, x: Tensor, causal: bool) -> Tensor:
    """stride==1 DecoderBlock: ConvTranspose1d(k=1, stride=1) (vq/module.py:129-136)."""
    if causal:
        w, b = _wn(sd, prefix + "conv.", x.dtype)
        return F.conv_transpose1d(x, w, b, stride=1)[..., :-1]
    w, b = _wn(sd, prefix, x.dtype)
    return F.conv_transpose1d(x, w, b, stride=1)


# ----------------------------------------------------------------------------
# driver-level forwards
# ----------------------------------------------------------------------------
def cast_sd(sd: SD, dtype) -> SD:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


@torch.no_grad()
def encode_to_indices(enc_sd: SD, dec_sd: SD, cfg: dict, x: Tensor) -> Dict[str, Tensor]:
    """extract_indices.BigCodecModel.forward (extract_indices.py:347-371) + the latents."""
    z = encoder_forward(enc_sd, cfg["codec_encoder"], x)
    z_q, idx, loss, margin = quantize(dec_sd, cfg["codec_decoder"], z)
    return {"z": z, "z_q": z_q, "indices": idx, "loss": loss, "margin": margin}


@torch.no_grad()
def round_trip(enc_sd: SD, dec_sd: SD, cfg: dict, x: Tensor) -> Dict[str, Tensor]:
    """inference_full.BigCodecModel.forward (inference_full.py:557-561)."""
    out = encode_to_indices(enc_sd, dec_sd, cfg, x)
    out["x_rec"] = decoder_forward(dec_sd, cfg["codec_decoder"], out["z_q"])
    return out


def indices_to_int16(indices: Tensor):
    """On-disk form written by extract_indices.py:520-532 for one utterance:
    [n_q,1,T'] -> squeeze(1) -> [n_q,T'] -> permute -> (T', n_q) int16."""
    import numpy as np
    idx = indices.squeeze(1)
    if idx.ndim == 2:
        idx = idx.permute(1, 0)
    arr = idx.cpu().numpy().astype(np.int16)
    return arr


# ----------------------------------------------------------------------------
# the steps on either side of the hot path (codebook statistics, on-disk layout)
# ----------------------------------------------------------------------------
def codebook_perplexity(indices, codebook_size: int) -> float:
    """CodebookPerplexity.update + compute (lightning_module.py:33-51): one-hot counts over all indices,
    p = counts / total, entropy over the non-zero p, exp."""
    import numpy as np
    idx = np.asarray(indices).astype(np.int64).reshape(-1)
    counts = np.bincount(idx, minlength=codebook_size).astype(np.float64)
    total = counts.sum()
    if total == 0:
        return 0.0
    p = counts / total
    p = p[p > 0]
    return float(np.exp(-(p * np.log(p)).sum()))


def codebook_utilization(indices, codebook_size: int) -> float:
    """CodebookUtilization.update + compute (lightning_module.py:62-69): used-code mask, used / K."""
    import numpy as np
    idx = np.asarray(indices).astype(np.int64).reshape(-1)
    used = np.zeros(codebook_size, dtype=bool)
    used[idx] = True
    return float(used.sum() / codebook_size)


def calculate_perplexity(counter, codebook_size: int):
    """inference_full.calculate_perplexity (inference_full.py:570-604): (normalised perplexity, perplexity) from a
    {index: count} mapping; out-of-range keys count towards the total but carry no probability; 0.0 when empty."""
    import numpy as np
    total = sum(counter.values())
    if total == 0:
        return 0.0
    probs = np.zeros(codebook_size)
    for i, c in counter.items():
        if i < codebook_size:
            probs[i] = c / total
    nz = probs[probs > 0]
    ent = -np.sum(nz * np.log(nz))
    return float(np.exp(ent / np.log(codebook_size))), float(np.exp(ent))


def index_file_path(output_dir: str, subset: str, fileid: str) -> str:
    """Where extract_indices.py:534-556 writes an utterance: <out>/<subset>/<speaker>/<chapter>/<fileid>.npy with
    speaker / chapter = the first two '_'-separated fields of the file id, else the first two '-'-separated
    fields, else 'unknown'."""
    import os
    try:
        if "_" in fileid:
            parts = fileid.split("_")
            speaker, chapter = parts[0], parts[1]
        elif "-" in fileid:
            parts = fileid.split("-")
            speaker, chapter = parts[0], parts[1]
        else:
            speaker, chapter = "unknown", "unknown"
    except IndexError:
        speaker, chapter = "unknown", "unknown"
    return os.path.join(output_dir, subset, speaker, chapter, f"{fileid}.npy")


def pad_to_stride(waveform: Tensor, stride: int) -> Tensor:
    """extract_indices.py:134-136: right-pad the last axis with zeros to a multiple of ``stride`` (no-op when aligned)."""
    t = waveform.shape[-1]
    if stride and t % stride != 0:
        waveform = F.pad(waveform, (0, stride - t % stride), mode="constant", value=0)
    return waveform
