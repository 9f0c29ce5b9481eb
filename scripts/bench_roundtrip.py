#!/usr/bin/env python
"""BASELINE.json configs[2]: full round trip (inference_full.py): encode -> indices -> decode waveform,
batch 64 x 10 s, with a reconstruction tolerance check against the CPU oracle on a sample."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import configs, ops, synth
from audiotokenization_b200.model import BigCodecModel
from oracle import bigcodec_oracle as oracle

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--micro-batch", type=int, default=64, help="clips per model() call; the whole batch by default so the LSTM runs once over it")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--layer-table", default=None)
    args = ap.parse_args()
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=args.precision)
    T = int(args.seconds * 16000)
    x = synth.fast_synth_batch(0, args.batch, T).cuda()

    def step():
        outs = []
        for b0 in range(0, args.batch, args.micro_batch):
            outs.append(model(x[b0:b0 + args.micro_batch], round_trip=True))
        return outs

    for _ in range(3):
        outs = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        outs = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # tolerance check on clip 0 against the oracle
    want = oracle.round_trip(enc_sd, dec_sd, cfg, x[:1].cpu())
    got = outs[0]
    idx = got["indices"][:, :1].cpu()
    agree = float((idx == want["indices"]).float().mean())
    decided = want["margin"] > 1e-5
    yr = got["x_rec"][:1].cpu()
    rel_y = float((yr - want["x_rec"]).norm() / want["x_rec"].norm())
    line = {"workload": f"configs[2] round trip: batch {args.batch} x {args.seconds:g} s, base model", "precision": args.precision,
            "ms_per_step": ms, "audio_s_per_s": args.batch * args.seconds / ms * 1e3,
            "index_agreement_clip0": agree, "exact_where_margin_gt_1e-5": bool(torch.equal(idx[decided], want["indices"][decided])),
            "waveform_rel_err_clip0": rel_y, "waveform_rel_err_valid": agree == 1.0}
    if args.layer_table:
        ops.PROFILE = []
        step()
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
