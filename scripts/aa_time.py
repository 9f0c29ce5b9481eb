#!/usr/bin/env python
"""Debug: time the anti-aliased activation kernel alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops, synth
g = torch.Generator().manual_seed(5)
out = []
for C, T, B in ((32, 480000, 8), (64, 240000, 8), (512, 2400, 64)):
    x = (torch.randn(B, T, C, generator=g) * 1.5).cuda()
    a = torch.exp(torch.randn(C, generator=g) * 0.3).cuda()
    ib = (1.0 / (torch.exp(torch.randn(C, generator=g) * 0.3) + 1e-9)).cuda()
    fir = synth.kaiser_sinc_filter12().reshape(-1).cuda()
    for _ in range(3):
        ops.snake(x, a, ib, antialias=True, fir=fir)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.snake(x, a, ib, antialias=True, fir=fir)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out.append(f"C={C}: {8.0 * B * T * C / ms / 1e6:6.0f} GB/s")
print(os.environ.get("BC_LIB_PATH", "default"), " | ".join(out))
