#!/usr/bin/env python
"""bench.py -- encode -> VQ-indices throughput of the BigCodec hot path on N B200s.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (N > 1 under torchrun, one rank per
GPU) prints ONE JSON line on rank 0.  A "step" is one pass of waveform -> encoder -> VQ indices
over this rank's shard of BASELINE.json configs[1]: 4096 synthetic 30 s 16 kHz clips sharded by
utterance over 8 GPUs = 512 clips (15 360 audio-seconds) per GPU, weak scaling.

  value     whole-job audio-seconds per wall-second, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same through the host-buffer API (BigCodecModel.extract_indices): pinned host waveforms
            -> H2D -> encode -> int16 indices -> D2H, copies inside the timed region
  roofline  the dense-contraction kernels (conv / transposed conv / LSTM input projection) vs the
            measured bf16 tensor peak: algorithmic FLOPs / CUDA-event time of those launches
  cpu_baseline   the CPU oracle port (torch CPU library calls, the reference's own arithmetic) on the
            host cores, bounded sample (rank 0, N = 1)

``--impl reference`` times the CPU implementation only (the reference's PyTorch-CPU arithmetic as
restated in oracle/; /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

METRIC = "audio_seconds_encoded_to_indices_per_second"
UNIT = "audio-s/s"
ENC_GFLOP_PER_AUDIO_S = {"base": 6.860 + 0.012, "debug": 3.523 + 0.007, "config9_base": 4.175 + 0.007,
                         "default": 50.983 + 0.013}   # BASELINE.md section 3 (encoder + VQ)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="base")
    ap.add_argument("--precision", default=os.environ.get("BC_PRECISION", "bf16x3"), choices=["fp32", "bf16", "bf16x3"],
                    help="arithmetic of the dense contractions for the headline number: bf16x3 (default) = tcgen05 with "
                         "hi/lo split operands, fp32-class accuracy (meets the <=1e-3 / bit-exact-index contract); "
                         "bf16 = single-pass tcgen05 (fast mode, ~1e-2 latent error); fp32 = CUDA-core FFMA")
    ap.add_argument("--also", default=os.environ.get("BC_ALSO", "bf16"),
                    help="comma-separated extra precisions measured device-resident and reported under 'modes'")
    ap.add_argument("--clips-per-gpu", type=int, default=512)
    ap.add_argument("--clip-seconds", type=float, default=30.0)
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("BC_MICRO_BATCH", "8")))
    ap.add_argument("--deep-batch", type=int, default=int(os.environ.get("BC_DEEP_BATCH", "64")),
                    help="clips per launch of the last strided stages of the conv stack (model._FrontPipeline)")
    ap.add_argument("--rnn-batch", type=int, default=int(os.environ.get("BC_RNN_BATCH", "512")))
    ap.add_argument("--cpu-sample-clips", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-table", default=None, help="write a per-layer timing table (markdown) here")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip `other_configs` (BASELINE configs[2..4], anti-aliased path) and the eager-GPU baseline")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        # "under load": drop samples far below the maximum seen while busy
        busy = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_encode_rate(cfg, enc_sd, dec_sd, clips, clip_samples, reps=1):
    """audio-s/s of the CPU oracle port on this host: ``clips`` x ``clip_samples`` per pass."""
    from audiotokenization_b200 import synth
    from oracle import bigcodec_oracle as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    x = synth.fast_synth_batch(0, clips, clip_samples)
    with torch.no_grad():
        oracle.encode_to_indices(enc_sd, dec_sd, cfg, x[:1, :, : min(clip_samples, 32000)])  # warm-up (thread pools)
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            out = oracle.encode_to_indices(enc_sd, dec_sd, cfg, x)
            best = min(best, time.perf_counter() - t0)
    assert out["indices"].shape[-1] > 0
    return clips * clip_samples / 16000.0 / best, best


# ----------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """CPU arm: the reference's PyTorch-CPU arithmetic (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    from audiotokenization_b200 import configs, synth
    cfg = configs.get_config(args.model)
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    clip_samples = int(round(args.clip_seconds * 16000))
    clips = args.cpu_sample_clips
    cores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        cpu_encode_rate(cfg, enc_sd, dec_sd, 1, clip_samples)
    times = []
    for _ in range(args.steps):
        _, t = cpu_encode_rate(cfg, enc_sd, dec_sd, clips, clip_samples)
        times.append(t)
    total = sum(times)
    value = args.steps * clips * args.clip_seconds / total
    sample = f"{clips} of the {args.clips_per_gpu} clips x {args.clip_seconds:g} s per step (bounded sample of the same workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, "fp32"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(args, world, precision):
    return {"workload": "configs[1]: batched encode->indices of 4096 synthetic 30 s 16 kHz clips sharded by utterance "
                        "across 8xB200 (512 clips per GPU, weak scaling)",
            "model": f"BigCodec {args.model} (cfgs/config11/model/base.yaml)" if args.model == "base" else args.model,
            "clips_per_gpu": args.clips_per_gpu, "clip_seconds": args.clip_seconds, "sample_rate": 16000,
            "global_clips": args.clips_per_gpu * world, "micro_batch": args.micro_batch, "deep_batch": args.deep_batch, "rnn_batch": args.rnn_batch,
            "precision": precision,
            "weights": "random-init, seed 0 (biases / snake alpha,beta / weight-norm gains randomised)",
            "l2_policy": "inputs_larger_than_l2 (983 MB of waveforms per step; every activation tensor > 126 MB)",
            "parallelism": f"utterance-sharded x{world}, no collective on the data path"}


def _emit(line: dict) -> None:
    """The ONE JSON line of the contract goes to the process's original stdout; everything else that libraries print to
    fd 1 during the run (e.g. NCCL's version banner at communicator creation) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    os.write(fd, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from audiotokenization_b200 import _cabi, configs, ops, sharding, synth
    from audiotokenization_b200.model import BigCodecModel
    import bench_configs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    cfg = configs.get_config(args.model)
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device=str(dev), precision=args.precision)
    clip_samples = int(round(args.clip_seconds * 16000))
    n_local = args.clips_per_gpu
    first = rank * n_local                       # this rank's shard of the global utterance list
    host = torch.empty((n_local, 1, clip_samples), dtype=torch.float32, pin_memory=True)
    synth.fast_synth_batch(first, n_local, clip_samples, out=host)
    x_dev = host.to(dev)
    audio_s_local = n_local * args.clip_seconds

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput ---------------------------------------------------------
    keep = {}

    def step_device():
        keep["idx"] = model.indices_device(x_dev, micro_batch=args.micro_batch, rnn_batch=args.rnn_batch, deep_batch=args.deep_batch)

    for _ in range(args.warmup):
        step_device()
    ops.STATS["launches"] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.STATS["launches"]
    value = world * audio_s_local * args.steps / (ms_total / 1000.0)

    # ---- end to end through the host-buffer API ----------------------------------------------
    host_group = dist.new_group(backend="gloo") if world > 1 else None   # CPU group: the host-side gather of the int16 indices

    def step_e2e():
        keep["i16"] = model.extract_indices(host, micro_batch=args.micro_batch, rnn_batch=args.rnn_batch, deep_batch=args.deep_batch)
        if world > 1:   # "results gathered on the host" (north_star): inside the timed region, no data-path collective on the GPUs
            keep["gathered"] = sharding.gather_indices_to_rank0(keep["i16"], group=host_group)

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * audio_s_local * args.steps / (ms_e2e / 1000.0)
    i16 = keep["i16"]
    same = bool((keep["idx"].cpu().numpy() == i16).all())

    # ---- roofline of the dense-contraction kernels (one instrumented step) ---------------------
    ops.PROFILE = []
    step_device()
    torch.cuda.synchronize(dev)
    prof, ops.PROFILE = ops.PROFILE, None
    by_kind, by_layer, by_kernel = {}, {}, {}
    _pk = load_peaks()
    RIDGE = _pk["bf16_tflops_sustained"] * 1e12 / (_pk["hbm_gbs"] * 1e9)
    for key, flops, a, b, kernel, nbytes in prof:
        ms = a.elapsed_time(b)
        for table, k in ((by_kind, key[0]), (by_layer, key), (by_kernel, kernel)):
            d = table.setdefault(k, [0.0, 0.0, 0, 0.0, 0.0])
            d[0] += flops
            d[1] += ms
            d[2] += 1
            d[3] += nbytes
            if nbytes > 0 and flops / nbytes < RIDGE:   # time spent in launches whose own FLOP/byte is below the ridge
                d[4] += ms
    if args.layer_table and rank == 0:
        with open(args.layer_table, "w") as f:
            f.write("| kind | C_in | C_out | K | stride | dil | T_out | B | prec | launches | ms/step | TFLOP/s |\n"
                    "|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|\n")
            for k, (fl, ms, n, _, _) in sorted(by_layer.items(), key=lambda kv: -kv[1][1]):
                f.write("| " + " | ".join(str(v) for v in k) + f" | {n} | {ms:.2f} | {fl / ms / 1e9 if ms else 0:.1f} |\n")
            f.write("\n| kernel | launches | ms/step | TFLOP/s | algorithmic GB/s |\n|---|---:|---:|---:|---:|\n")
            for k, (fl, ms, n, nb, _) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {n} | {ms:.2f} | {fl / ms / 1e9 if ms else 0:.1f} | {nb / ms / 1e6 if ms else 0:.0f} |\n")
    CONV_KINDS = ("conv1d", "convtr1d", "resunit")
    conv_flops = sum(v[0] for k, v in by_kind.items() if k in CONV_KINDS)
    conv_ms = sum(v[1] for k, v in by_kind.items() if k in CONV_KINDS)
    step_ms = ms_total / args.steps
    peaks = load_peaks()
    # dominant kernel = the tcgen05 convolution kernel family with the largest share of the step
    conv_kernels = {k: v for k, v in by_kernel.items() if k in ("conv_stream_kernel", "ru_persist_kernel", "ru_group_kernel", "ru_pair_kernel", "conv1d_tc_kernel", "conv1d_f32_kernel")}
    dom = max(conv_kernels, key=lambda k: conv_kernels[k][1]) if conv_kernels else None
    d_fl, d_ms, d_n, d_bytes, d_ms_hbm = conv_kernels[dom] if dom else (0.0, 0.0, 0, 0.0, 0.0)
    # which roof bounds it: arithmetic intensity of its launches against the ridge of the measured peaks
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    intensity = d_fl / d_bytes if d_bytes > 0 else float("inf")
    hbm_bound = d_ms_hbm > 0.5 * d_ms    # the roof under which most of the kernel's time is spent (per-launch FLOP/byte vs ridge)
    tflops = d_fl / (d_ms / 1000.0) / 1e12 if d_ms > 0 else 0.0
    gbs = d_bytes / (d_ms / 1000.0) / 1e9 if d_ms > 0 else 0.0
    traffic = None
    try:   # average DRAM bytes per launch of that kernel from the committed ncu pass (profiles/, same launch shapes)
        with open(os.path.join(REPO, "profiles", "kernel_traffic.json")) as f:
            traffic = json.load(f).get(args.precision, {}).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        traffic = None

    def kernel_entry(v):
        fl, ms, n, nb, ms_hbm = v
        return {"launches": n, "ms_per_step": ms, "tflops": fl / ms / 1e9 if ms else 0.0,
                "algorithmic_gbs": nb / ms / 1e6 if ms else 0.0, "flop_per_byte": fl / nb if nb else None,
                "share_of_step": ms / step_ms if step_ms else None,
                "time_share_in_hbm_bound_launches": ms_hbm / ms if ms else None}

    roofline = {"bound": "hbm" if hbm_bound else "tensor", "kernel": dom,
                "achieved": gbs if hbm_bound else tflops,
                "peak": peaks["hbm_gbs"] if hbm_bound else peaks["bf16_tflops_sustained"],
                "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": (gbs / peaks["hbm_gbs"]) if hbm_bound else (tflops / peaks["bf16_tflops_sustained"]),
                "traffic": traffic,
                "peak_source": f"{peaks['source']} ({'copy bandwidth' if hbm_bound else 'bf16 sustained'}, kernel timed inside a long step)",
                "definition": "dominant kernel = tcgen05 conv kernel family with the largest share of the step; algorithmic bytes "
                              "(fp32 activation read once + written once) or FLOPs (2*MAC, bf16x3 counted once) of its launches in "
                              "one instrumented step / their CUDA-event time on the launch stream; bound = the roof under which most of "
                              f"the kernel's time is spent (each launch's FLOP/byte against the ridge of the measured peaks, {ridge:.0f}; "
                              f"aggregate {intensity:.0f})",
                "launches_per_step": d_n, "avg_launch_ms": d_ms / d_n if d_n else None,
                "algorithmic_bytes_per_launch": d_bytes / d_n if d_n else None,
                "flop_per_launch": d_fl / d_n if d_n else None, "tflops": tflops,
                "tensor_frac": tflops / peaks["bf16_tflops_sustained"],
                # the split-precision mode issues three bf16 MMAs per product (hi*hi + hi*lo + lo*hi): the fraction of the tensor
                # roof its ISSUED arithmetic reaches (informational; `frac` counts the algorithmic FLOPs once, so its ceiling is 1/3)
                "mma_passes": 3 if args.precision == "bf16x3" else 1,
                "issued_tensor_frac": (3 if args.precision == "bf16x3" else 1) * tflops / peaks["bf16_tflops_sustained"],
                "kernel_ms_per_step": d_ms, "share_of_step": d_ms / step_ms if step_ms else None,
                "all_kernels": {k: kernel_entry(v) for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])},
                "dense_contractions": {"tflops": conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0,
                                       "ms_per_step": conv_ms, "share_of_step": conv_ms / step_ms if step_ms else None,
                                       "tensor_frac": conv_flops / (conv_ms / 1000.0) / 1e12 / peaks["bf16_tflops_sustained"] if conv_ms > 0 else 0.0},
                "lstm_ms_per_step": by_kind.get("lstm", [0, 0, 0])[1],
                "whole_step_frac": (value / world) * ENC_GFLOP_PER_AUDIO_S.get(args.model, 0.0) / 1e3
                / peaks["bf16_tflops_sustained"]}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        clips = args.cpu_sample_clips
        v, t = cpu_encode_rate(cfg, enc_sd, dec_sd, clips, clip_samples, reps=2)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{clips} of the {n_local} clips x {args.clip_seconds:g} s, best of 2 passes "
                                  f"({t:.1f} s per pass), oracle/bigcodec_oracle.py (torch CPU fp32)",
                        }
        # parity spot-check of the benchmarked run against the oracle on the sample
        from oracle import bigcodec_oracle as oracle
        with torch.no_grad():
            want = oracle.encode_to_indices(enc_sd, dec_sd, cfg, host[:1])
        got = torch.from_numpy(i16[:1, :, 0].astype("int64"))
        decided = want["margin"][0] > 1e-5
        cpu_baseline["parity_clip0"] = {
            "index_agreement": float((got == want["indices"][0]).float().mean()),
            "exact_where_margin_gt_1e-5": bool(torch.equal(got[decided], want["indices"][0][decided]))}

    # ---- the library Blackwell path: the reference's arithmetic through PyTorch eager on this GPU (rank 0, N = 1) ----
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_extras:
        try:
            ge = bench_configs.gpu_eager(cfg, enc_sd, dec_sd, x_dev, clips=8)
            for name in ("tf32_default", "fp32_allow_tf32_false"):
                got_e = ge.pop(name + "_indices")[:1]
                ge[name]["index_agreement_vs_cpu_oracle_clip0"] = float((got_e == want["indices"][0]).float().mean())
                ge[name]["speedup_of_this_repo"] = value / ge[name]["value"]
            gpu_eager = ge
        except Exception as e:   # a baseline that cannot run must not take the headline number down with it
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    # ---- the other BASELINE configurations and the stand-alone kernels north_star names ----------------------
    other = {}
    if not args.no_extras:
        def guarded(name, fn):
            try:
                other[name] = fn()
            except Exception as e:
                other[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

        # configs[3] runs at every N: the conv front end of ONE recording is dealt out over the ranks
        guarded("configs3_longform", lambda: bench_configs.longform(model, minutes=10.0, world=world, check_whole=(world == 1)))
        if rank == 0 and world == 1:
            guarded("configs2_round_trip", lambda: bench_configs.round_trip(model, enc_sd=enc_sd, dec_sd=dec_sd, cfg=cfg))
            guarded("configs4_vq_sweep", lambda: bench_configs.vq_sweep())
            guarded("antialias", lambda: bench_configs.antialias(lambda c: synth.make_state_dicts(c, seed=0), precision=args.precision))

    # ---- other arithmetic modes (device-resident only; reported, never substituted for the headline) ----
    modes = {}
    for mode in [m for m in args.also.split(",") if m and m != args.precision]:
        model.precision = mode
        for _ in range(max(1, min(args.warmup, 2))):
            step_device()
        ms_m = timed(step_device, args.steps)
        modes[mode] = {"value": world * audio_s_local * args.steps / (ms_m / 1000.0), "unit": UNIT,
                       "ms_per_step": ms_m / args.steps}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            got_m = keep["idx"][:1, :, 0].cpu().to(torch.int64)
            modes[mode]["index_agreement_clip0"] = float((got_m == want["indices"][0]).float().mean())
    model.precision = args.precision

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16x3(f32-class)"}[args.precision], "data": "synthetic",
            "config": workload_config(args, world, args.precision),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4),
                    "d2h_bytes_per_step": int(i16.nbytes), "ms_per_step": ms_e2e / args.steps,
                    "matches_device_path": same},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "modes": modes,
            "policy": _cabi.policy(),
        }
        if world > 1:
            g = keep.get("gathered")
            line["e2e"]["host_gather"] = {"backend": "gloo", "inside_timed_region": True,
                                          "gathered_shape": list(g.shape) if g is not None else None}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if gpu_eager is not None:
            line["gpu_eager_baseline"] = gpu_eager
        if other:
            line["other_configs"] = other
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
