#!/usr/bin/env python
"""BASELINE.json configs[2]: full round trip (inference_full.py): encode -> indices -> decode waveform,
batch 64 x 10 s, with a reconstruction tolerance check against the CPU oracle on a sample (bench_configs.round_trip)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_configs
from audiotokenization_b200 import configs, ops, synth
from audiotokenization_b200.model import BigCodecModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--layer-table", default=None)
    args = ap.parse_args()
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=args.precision)
    line = bench_configs.round_trip(model, args.batch, args.seconds, args.steps, enc_sd=enc_sd, dec_sd=dec_sd, cfg=cfg)
    if args.layer_table:
        x = synth.fast_synth_batch(2000, args.batch, int(args.seconds * 16000)).cuda()
        ops.PROFILE = []
        model(x, round_trip=True)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        by = {}
        for key, fl, a, b, _kernel, _nbytes in prof:
            d = by.setdefault(key, [0.0, 0.0, 0]); d[0] += fl; d[1] += a.elapsed_time(b); d[2] += 1
        with open(args.layer_table, "w") as f:
            f.write("| kind | C_in | C_out | K | stride | dil | T_out | B | prec | launches | ms/step | TFLOP/s |\n|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|\n")
            for k, (fl, m, n) in sorted(by.items(), key=lambda kv: -kv[1][1]):
                f.write("| " + " | ".join(str(v) for v in k) + f" | {n} | {m:.2f} | {fl / m / 1e9 if m else 0:.1f} |\n")
        line["timed_ms_in_table"] = sum(v[1] for v in by.values())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
