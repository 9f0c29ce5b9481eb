#!/usr/bin/env python
"""Parity statistics of SURVEY.md section 8d on the benchmark workload: indices vs the CPU oracle (fp32 and fp64) over
several 30 s clips, overall and bucketed by the oracle's top-1/top-2 cosine margin; latent relative error."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import configs, synth
from audiotokenization_b200.model import BigCodecModel
from oracle import bigcodec_oracle as oracle

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--precisions", default="bf16x3,fp32,bf16")
args = ap.parse_args()
cfg = configs.get_config("base")
enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
T = int(args.seconds * 16000)
x = synth.fast_synth_batch(0, args.clips, T)
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    want = oracle.encode_to_indices(enc_sd, dec_sd, cfg, x)
    want64 = oracle.encode_to_indices(oracle.cast_sd(enc_sd, torch.float64), oracle.cast_sd(dec_sd, torch.float64), cfg, x[:2].double())
ref_idx, margin, z_ref = want["indices"][0], want["margin"][0], want["z"]
edges = [("margin>1e-2", 1e-2, float("inf")), ("1e-3..1e-2", 1e-3, 1e-2), ("1e-5..1e-3", 1e-5, 1e-3), ("<1e-5", -1.0, 1e-5)]
report = {"workload": f"{args.clips} synthetic clips x {args.seconds:g} s, base model, seed-0 weights",
          "frames": int(ref_idx.numel()), "oracle_fp32_vs_fp64_index_agreement_2clips": float((want["indices"][0][:2] == want64["indices"][0]).float().mean()),
          "margin_histogram": {n: int(((margin > lo) & (margin <= hi)).sum()) for n, lo, hi in edges}}
for prec in args.precisions.split(","):
    model = BigCodecModel(cfg, enc_sd, dec_sd, device="cuda", precision=prec)
    idx, _, z_cl = model.encode_indices_cl(x.cuda().reshape(args.clips, T, 1))
    got = idx[0].cpu().long()
    z = z_cl.permute(0, 2, 1).cpu()
    ok = got == ref_idx
    r = {"index_agreement": float(ok.float().mean()), "latent_rel_l2": float((z - z_ref).norm() / z_ref.norm()),
         "latent_max_abs_over_rms": float((z - z_ref).abs().max() / z_ref.pow(2).mean().sqrt()),
         "exact_where_margin_gt_1e-5": bool(ok[margin > 1e-5].all()), "mismatches": int((~ok).sum()),
         "largest_margin_of_a_mismatch": float(margin[~ok].max()) if (~ok).any() else 0.0,
         "by_margin": {n: {"frames": int(((margin > lo) & (margin <= hi)).sum()),
                           "agree": float(ok[(margin > lo) & (margin <= hi)].float().mean()) if ((margin > lo) & (margin <= hi)).any() else None}
                       for n, lo, hi in edges}}
    report[prec] = r
print(json.dumps(report, indent=1))
