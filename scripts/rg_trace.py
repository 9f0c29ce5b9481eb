#!/usr/bin/env python
"""Debug: phase timeline of one warpgroup of ru_group_kernel (needs a BC_TRACE build).  usage: rg_trace.py [C] [dil]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import _cabi
from audiotokenization_b200.vq import module as M

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dil = int(sys.argv[2]) if len(sys.argv) > 2 else 9
B, T = 8, 480000 * 32 // C
ru = M.ResidualUnit(C, dilation=dil).cuda()
x = torch.randn(B, T, C, device="cuda")
M.set_precision("bf16x3")
lib = _cabi.load_library()
for _ in range(3):
    y = ru.forward_cl(x)
torch.cuda.synchronize()
trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.bc_debug_set_ru_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ru.forward_cl(x); e1.record()
torch.cuda.synchronize()
lib.bc_debug_set_ru_trace(None)
t = trace.cpu().view(64, 16).double()
names = ["stage: loads + convert", "  group sync", "mma7 issue", "wait mma7 (residual loads issued)", "mid: acc1 -> A2", "  group sync", "mma1 issue", "wait mma1", "store: acc2 -> staging", "  group sync", "store: staging -> y", "  group sync / next tile"]
print(f"C={C} dil={dil}: kernel {e0.elapsed_time(e1)*1e3:.0f} us; per tile of ONE group: {float((t[11:60, 0] - t[10:59, 0]).mean()):.0f} cycles")
for j, n in enumerate(names):
    nxt = t[10:59, j + 1] if j < 11 else t[11:60, 0]
    print(f"   {n:36s} {float((nxt - t[10:59, j]).mean()):8.0f}")
