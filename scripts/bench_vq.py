#!/usr/bin/env python
"""BASELINE.json configs[4]: VQ-heavy sweep -- codebook sizes 8192/16384/32768 at codebook_dim 8, isolating the
fused factorized-VQ kernel (in_proj + L2 normalise + cosine argmax).  Prints one JSON line per K with the
achieved HBM GB/s (algorithmic: 2048 B in + 4 B out per frame) and FP32 TFLOP/s (2*C*D + 2*K*D per frame)
against the measured peaks, plus exactness against the CPU oracle on a sample."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops
from audiotokenization_b200.vq import FactorizedVectorQuantize

def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    N, C, D = 1 << 20, 512, 8
    g = torch.Generator().manual_seed(0)
    z = torch.randn(N, C, generator=g).cuda()
    for K in (8192, 16384, 32768):
        layer = FactorizedVectorQuantize(dim=C, codebook_size=K, codebook_dim=D, commitment=0.25).eval()
        layer._codebook.weight.data = torch.randn(K, D, generator=g)
        layer = layer.cuda()
        w_in, b_in = layer._proj("in_proj")
        _, cbn = layer._codebooks()
        for _ in range(3):
            idx, _, _ = ops.vq_encode(z, w_in, b_in, cbn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            idx, margin, _ = ops.vq_encode(z, w_in, b_in, cbn, want_margin=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # oracle check on a sample (torch CPU, reference formula)
        import torch.nn.functional as F
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
        print(json.dumps({
            "workload": f"configs[4] VQ sweep: {N} frames x {C} ch, K={K}, D={D}", "ms": ms,
            "frames_per_s": N / ms * 1e3, "achieved_gbs": N * bytes_per_frame / ms / 1e6,
            "hbm_frac": N * bytes_per_frame / ms / 1e6 / peaks["hbm_gbs"],
            "fp32_tflops": N * flop_per_frame / ms / 1e9,
            "exact_where_margin_gt_1e-5": bool(torch.equal(got[decided], ref[decided])),
            "agreement": float((got == ref).float().mean()),
            "frac_margin_lt_1e-5": float((margin < 1e-5).float().mean())}))

if __name__ == "__main__":
    main()
