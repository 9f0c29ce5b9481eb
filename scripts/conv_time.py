#!/usr/bin/env python
"""Times the un-fused convs of the encoder at the bench's launch shapes (8 x 30 s clips per launch for blocks 1-2, 64 clips for
blocks 3-4, 512 x 2400 frames for the LSTM input projection) with CUDA events: us per launch, effective TFLOP/s, algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200.vq import module as M, activations

M.set_precision(os.environ.get("PREC", "bf16x3"))
torch.manual_seed(0)
CASES = ((32, 64, 4, 2, 480000, 8), (64, 128, 8, 4, 240000, 8), (128, 256, 10, 5, 60000, 64), (256, 512, 10, 5, 12000, 64),
         (512, 512, 3, 1, 2400, 512), (512, 2048, 1, 1, 2400, 512))
for ci, co, k, s, T, clips in CASES:
    pad = (s // 2 + s % 2) if s > 1 else (k - 1) // 2
    m = M.WNConv1d(ci, co, kernel_size=k, stride=s, padding=pad).cuda()
    act = activations.SnakeBeta(ci, alpha_logscale=True).cuda() if k > 1 else None
    x = torch.randn(clips, T, ci, device="cuda")
    for _ in range(3):
        y = m.forward_cl(x, act=act)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n):
        y = m.forward_cl(x, act=act)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) / n * 1e3
    fl = 2.0 * clips * y.shape[1] * ci * co * k
    by = 4.0 * (x.numel() + y.numel())
    print(f"{ci:4d}->{co:4d} k{k:2d} s{s} T_in={T:6d} x{clips:3d}: {us:8.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s  {by / us * 1e-3:7.0f} GB/s", flush=True)
    del x, y, m
