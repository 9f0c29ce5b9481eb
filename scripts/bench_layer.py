#!/usr/bin/env python
"""Per-layer timing of the wide encoder layers (CUDA events, inputs larger than L2 across the rotation).
usage: bench_layer.py [precision] [clips]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200.vq import module as M, activations

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 8
only = sys.argv[3] if len(sys.argv) > 3 else ""
M.set_precision(prec)
torch.manual_seed(0)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
cases = [("ru", 64, 1, 240000), ("ru", 64, 3, 240000), ("ru", 128, 9, 60000), ("ru", 256, 9, 12000), ("ru", 128, 1, 60000), ("ru", 64, 9, 240000), ("ru", 32, 9, 480000),
         ("conv", 32, 64, 4, 2, 480000), ("conv", 64, 128, 8, 4, 240000), ("conv", 128, 256, 10, 5, 60000),
         ("conv", 256, 512, 10, 5, 12000), ("conv", 512, 2048, 1, 1, 2400 * 8), ("conv", 512, 512, 3, 1, 2400 * 8)]
for c in cases:
    if only and only not in "_".join(str(v) for v in c):
        continue
    if c[0] == "ru":
        _, C, dil, T = c
        m = M.ResidualUnit(C, dilation=dil).cuda()
        xs = [torch.randn(clips, T, C, device="cuda") for _ in range(3)]
        i = [0]

        def fn():
            i[0] = (i[0] + 1) % 3
            return m.forward_cl(xs[i[0]])
        ms = timeit(fn)
        fl = 2.0 * clips * T * C * C * 8
        name = f"resunit C={C} dil={dil} T={T}"
    else:
        _, ci, co, k, s, T = c
        pad = (s // 2 + s % 2) if s > 1 else (k - 1) // 2
        m = M.WNConv1d(ci, co, kernel_size=k, stride=s, padding=pad).cuda()
        act = activations.SnakeBeta(ci, alpha_logscale=True).cuda() if s > 1 else None   # EncoderBlock: snake -> strided conv
        xs = [torch.randn(clips, T, ci, device="cuda") for _ in range(3)]
        i = [0]

        def fn():
            i[0] = (i[0] + 1) % 3
            return m.forward_cl(xs[i[0]], act=act)
        ms = timeit(fn)
        fl = 2.0 * clips * m.out_length(T) * ci * co * k
        name = f"conv {ci}->{co} k={k} s={s} T_in={T}"
    print(f"{prec:7s} {name:40s} {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
