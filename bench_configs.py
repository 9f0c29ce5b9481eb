"""Measurements of the BASELINE.json configurations other than the headline one, and of the stand-alone kernels
north_star names.  ``bench.py`` reports them under ``other_configs`` (one JSON line, driver-visible); the scripts
under ``scripts/bench_*.py`` are thin command-line wrappers around the same functions.

  configs[2]  round trip encode -> indices -> decode, batch 64 x 10 s            ``round_trip``
  configs[3]  one 10-minute recording, chunked with overlap, 1..N GPUs            ``longform``
  configs[4]  VQ sweep K = 8192 / 16384 / 32768 at D = 8                           ``vq_sweep``
  anti-aliased activation: base model with antialias=True, and the fused stencil kernel alone   ``antialias``
  library Blackwell path: the reference's arithmetic through PyTorch eager on the same GPU       ``gpu_eager``

Every function times with CUDA events on the current stream after warm-up and returns a plain dict.
"""
from __future__ import annotations

import json
import os

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
ROUND_TRIP_GFLOP_PER_AUDIO_S = {"base": 6.860 + 0.012 + 7.027}
ENC_GFLOP_PER_AUDIO_S = {"base": 6.860 + 0.012}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def _time(fn, steps, warmup=2):
    """Median of per-step CUDA-event times (a step that hits a fresh cudaMalloc of the caching allocator would otherwise
    move a 3-step mean by 10 %)."""
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        out = fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return ms[len(ms) // 2], out


# ----------------------------------------------------------------------------------------------
def round_trip(model, batch=64, seconds=10.0, steps=5, check=True, enc_sd=None, dec_sd=None, cfg=None):
    """configs[2]: ``BigCodecModel.forward(round_trip=True)`` over the whole batch (the LSTMs run once over it)."""
    from audiotokenization_b200 import synth
    T = int(seconds * 16000)
    x = synth.fast_synth_batch(2000, batch, T).cuda()
    ms, out = _time(lambda: model(x, round_trip=True), steps)
    pk = peaks()
    rate = batch * seconds / ms * 1e3
    line = {"workload": f"configs[2] round trip encode->indices->decode: batch {batch} x {seconds:g} s, base model",
            "precision": model.precision, "ms_per_step": ms, "audio_s_per_s": rate,
            "roofline_frac": rate * ROUND_TRIP_GFLOP_PER_AUDIO_S["base"] / 1e3 / pk["bf16_tflops_sustained"],
            "roofline": f"13.90 GFLOP per audio-second (enc + VQ + dec) vs {pk['bf16_tflops_sustained']} TFLOP/s ({pk['source']}, bf16 sustained)"}
    if check and enc_sd is not None:
        from oracle import bigcodec_oracle as oracle
        want = oracle.round_trip(enc_sd, dec_sd, cfg, x[:1].cpu())
        idx = out["indices"][:, :1].cpu()
        decided = want["margin"] > 1e-5
        agree = float((idx == want["indices"]).float().mean())
        line["index_agreement_clip0"] = agree
        line["exact_where_margin_gt_1e-5"] = bool(torch.equal(idx[decided], want["indices"][decided]))
        # reconstruction tolerance: decode the ORACLE's quantised latents so a sub-margin index flip cannot mask it
        y = model.decoder(want["z_q"].cuda(), vq=False).cpu()
        line["waveform_rel_err_clip0"] = float((y - want["x_rec"]).norm() / want["x_rec"].norm())
        line["waveform_tolerance"] = 1e-3
    return line


# ----------------------------------------------------------------------------------------------
def longform(model, minutes=10.0, chunk_seconds=30.0, micro_batch=8, steps=2, world=1, check_whole=False):
    """configs[3]: one long recording; the conv front end over ``world`` GPUs, LSTM + VQ on rank 0."""
    import torch.distributed as dist
    from audiotokenization_b200 import synth
    hop = int(model.encoder.hop_length)
    T = int(minutes * 60 * 16000) // hop * hop
    x = synth.fast_synth_batch(99, 1, T)[0, 0].cuda()

    def step():
        return model.indices_longform(x, chunk_seconds=chunk_seconds, micro_batch=micro_batch)

    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[torch.cuda.current_device()])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    line = {"workload": f"configs[3] long-form: one {minutes:g}-minute recording, {chunk_seconds:g} s chunks + halo, conv front end "
                        f"over {world} GPU(s), LSTM + final conv + VQ on rank 0", "precision": model.precision, "n_gpus": world,
            "ms_per_recording": ms, "audio_s_per_s": minutes * 60 / ms * 1e3,
            "frames": int(out.shape[1]) if out is not None else None,
            "bound": "latency of the sequential LSTM (2 layers x frames steps at batch 1)"}
    if check_whole and out is not None:
        whole = model.indices_device(x.view(1, 1, -1), micro_batch=1, rnn_batch=1)
        line["equals_unchunked"] = bool(torch.equal(whole, out))
    return line


# ----------------------------------------------------------------------------------------------
def vq_sweep(frames=1 << 20, C=512, D=8, sizes=(8192, 16384, 32768), reps=5, margin=False):
    """configs[4]: the fused factorized-VQ kernel alone (in_proj + L2 normalise + cosine argmax -> int32)."""
    import torch.nn.functional as F
    from audiotokenization_b200 import ops
    from audiotokenization_b200.vq import FactorizedVectorQuantize
    pk = peaks()
    g = torch.Generator().manual_seed(0)
    z = torch.randn(frames, C, generator=g).cuda()
    # fp32 FMA roof at the clock the part sustains: 148 SMs x 128 lanes x 2 FLOP; quoted at the max clock (1.965 GHz)
    fma_roof = 148 * 128 * 2 * 1.965e9 / 1e12
    out = []
    for K in sizes:
        layer = FactorizedVectorQuantize(dim=C, codebook_size=K, codebook_dim=D, commitment=0.25).eval()
        layer._codebook.weight.data = torch.randn(K, D, generator=g)
        layer = layer.cuda()
        w_in, b_in = layer._proj("in_proj")
        _, cbn = layer._codebooks()
        ms, (idx, mg, _) = _time(lambda: ops.vq_encode(z, w_in, b_in, cbn, want_margin=margin), reps, warmup=2)
        zs = z[:4096].cpu()
        e = F.normalize(F.linear(zs, w_in.cpu(), b_in.cpu()))
        c = F.normalize(layer._codebook.weight.data.cpu())
        dist = e.pow(2).sum(1, keepdim=True) - 2 * e @ c.t() + c.pow(2).sum(1, keepdim=True).t()
        ref = (-dist).max(1)[1]
        top2 = (e @ c.t()).topk(2, dim=1).values
        decided = (top2[:, 0] - top2[:, 1]) > 1e-5
        got = idx[:4096].cpu().long()
        bytes_per_frame = C * 4 + 4
        flop_per_frame = 2 * C * D + 2 * K * D
        tf = frames * flop_per_frame / ms / 1e9
        out.append({"K": K, "ms": ms, "frames_per_s": frames / ms * 1e3,
                    "achieved_gbs": frames * bytes_per_frame / ms / 1e6,
                    "hbm_frac": frames * bytes_per_frame / ms / 1e6 / pk["hbm_gbs"],
                    "fp32_tflops": tf, "fp32_fma_frac": tf / fma_roof,
                    "exact_where_margin_gt_1e-5": bool(torch.equal(got[decided], ref[decided])),
                    "agreement": float((got == ref).float().mean())})
    return {"workload": f"configs[4] VQ sweep: {frames} frames x {C} channels, D = {D}, index-only (margin={'on' if margin else 'off'})",
            "fp32_fma_roof_tflops": fma_roof, "hbm_peak_gbs": pk["hbm_gbs"],
            "bound": "fp32 FMA (2*K*D + 2*C*D FLOP vs 2052 B per frame); both fractions reported", "sizes": out}


# ----------------------------------------------------------------------------------------------
def antialias(enc_sd_fn, clips=16, seconds=10.0, steps=2, precision="bf16x3"):
    """Anti-aliased Activation1d (up-FIR -> SnakeBeta -> down-FIR): the stand-alone fused stencil kernel against the
    HBM roof (8 B per element), and the base encoder with antialias=True end to end."""
    from audiotokenization_b200 import configs, ops, synth
    from audiotokenization_b200.model import BigCodecModel
    pk = peaks()
    res = {}
    g = torch.Generator().manual_seed(5)
    for C, T, B in ((32, 480000, 8), (64, 240000, 8), (128, 60000, 8), (512, 2400, 64)):
        x = (torch.randn(B, T, C, generator=g) * 1.5).cuda()
        a = torch.exp(torch.randn(C, generator=g) * 0.3).cuda()
        ib = (1.0 / (torch.exp(torch.randn(C, generator=g) * 0.3) + 1e-9)).cuda()
        fir = synth.kaiser_sinc_filter12().reshape(-1).cuda()
        ms_aa, _ = _time(lambda: ops.snake(x, a, ib, antialias=True, fir=fir), 5)
        ms_pl, _ = _time(lambda: ops.snake(x, a, ib), 5)
        nbytes = 8.0 * B * T * C
        res[f"C{C}_T{T}_B{B}"] = {"aa_ms": ms_aa, "aa_gbs": nbytes / ms_aa / 1e6, "aa_hbm_frac": nbytes / ms_aa / 1e6 / pk["hbm_gbs"],
                                 "plain_ms": ms_pl, "plain_gbs": nbytes / ms_pl / 1e6, "plain_hbm_frac": nbytes / ms_pl / 1e6 / pk["hbm_gbs"]}
    cfg = configs.get_config("base", antialias=True)
    enc_sd, dec_sd = enc_sd_fn(cfg)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=precision)
    x = synth.fast_synth_batch(3000, clips, int(seconds * 16000)).cuda()
    ms, _ = _time(lambda: model.indices_device(x, micro_batch=8), steps)
    return {"workload": "anti-aliased Activation1d: stand-alone fused stencil kernel (8 B per element vs HBM copy bandwidth) and base "
                        f"encoder with antialias=True, {clips} x {seconds:g} s", "precision": precision,
            "kernel": res, "hbm_peak_gbs": pk["hbm_gbs"],
            "encode_antialias_ms": ms, "encode_antialias_audio_s_per_s": clips * seconds / ms * 1e3}


# ----------------------------------------------------------------------------------------------
def gpu_eager(cfg, enc_sd, dec_sd, x_dev, clips=8):
    """The library Blackwell path: the reference's arithmetic -- the same functional restatement
    the CPU oracle uses -- through PyTorch eager (cuDNN / cuBLAS) on this GPU, as extract_indices.py:399,503-510 runs it
    on a CUDA box; once with the default TF32 convolutions and once with ``cudnn.allow_tf32 = False``."""
    from oracle import bigcodec_oracle as oracle
    dev = x_dev.device
    e_sd = {k: v.to(dev) for k, v in enc_sd.items()}
    d_sd = {k: v.to(dev) for k, v in dec_sd.items()}
    x = x_dev[:clips]
    seconds = x.shape[0] * x.shape[2] / 16000.0
    out = {"sample": f"{x.shape[0]} x {x.shape[2] / 16000:g} s clips per pass (batched; the reference itself runs batch 1), best of 3 "
                     "after 2 warm-up passes, oracle/bigcodec_oracle.py functions on CUDA tensors (torch eager: cuDNN convs, "
                     "cuDNN LSTM, cuBLAS)", "unit": "audio-s/s", "torch": torch.__version__}
    prev = torch.backends.cudnn.allow_tf32
    try:
        for name, flag in (("tf32_default", True), ("fp32_allow_tf32_false", False)):
            torch.backends.cudnn.allow_tf32 = flag
            best = float("inf")
            with torch.no_grad():
                for i in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    got = oracle.encode_to_indices(e_sd, d_sd, cfg, x)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        best = min(best, e0.elapsed_time(e1))
            out[name] = {"value": seconds / best * 1e3, "ms_per_pass": best}
            out[name + "_indices"] = got["indices"][0].cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return out
