#!/usr/bin/env python
"""Times the fused ResidualUnits at the bench's launch shapes (8 x 30 s clips for C = 32 / 64 / 128, 64 clips for C = 256)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200.vq import module as M

M.set_precision(os.environ.get("PREC", "bf16x3"))
torch.manual_seed(0)
for C, dil, T, clips in ((32, 1, 480000, 8), (32, 9, 480000, 8), (64, 1, 240000, 8), (64, 9, 240000, 8), (128, 9, 60000, 8), (256, 9, 12000, 64)):
    m = M.ResidualUnit(C, dilation=dil).cuda()
    x = torch.randn(clips, T, C, device="cuda")
    for _ in range(3):
        y = m.forward_cl(x)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n):
        y = m.forward_cl(x)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) / n * 1e3
    ref = m.forward_cl(x)
    print(f"ResidualUnit C={C:3d} dil={dil} T={T:6d} x{clips:2d}: {us:8.1f} us  {4.0 * 2 * x.numel() / us * 1e-3:7.0f} GB/s  finite={bool(torch.isfinite(ref).all())} sum={float(ref.double().sum()):.6e}", flush=True)
    del x, y, m
