#!/usr/bin/env python
"""Debug: per-stage timeline of the streamed-weight kernel (clock64 stamps of CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import _cabi
from audiotokenization_b200.vq import module as M

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
dil = int(sys.argv[3]) if len(sys.argv) > 3 else 9
B, T = 8, 60000 * 128 // C
if C == 256:
    T = 12000 * 4
M.STREAM_RU_MIN_C[0] = 32
ru = M.ResidualUnit(C, dilation=dil).cuda()
x = torch.randn(B, T, C, device="cuda")
M.set_precision(prec)
lib = _cabi.load_library()
for _ in range(3):
    y = ru.forward_cl(x)
torch.cuda.synchronize()
trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.bc_debug_set_stream_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ru.forward_cl(x); e1.record()
torch.cuda.synchronize()
lib.bc_debug_set_stream_trace(None)
t = trace.cpu().view(64, 16)
names = ["p_start", "p_done", "m7_go", "m7_iss", "mid_go", "mid_done", "m1_go", "m1_iss", "st_acc", "st_done", "waitA", "waitB", "p_blk"]
base = int(t[0, 0])
ntile = B * ((T + 127) // 128)
print(f"C={C} {prec} dil={dil}: kernel {e0.elapsed_time(e1)*1e3:.0f} us, tiles/CTA {ntile/148:.1f}")
print("tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(2, 14):
    print(f"{i:4d} " + " ".join((f"{int(t[i, j]) - base:9d}" if j < 10 else f"{int(t[i, j]):9d}") if int(t[i, j]) else "        -" for j in range(13)))
n = min(60, ntile // 148 - 1)
d = (t[6:n, 9] - t[5:n - 1, 9]).float()
print("steady-state cycles per tile (store_done deltas):", float(d.mean()))
for a, b, label in [(0, 1, "PRODUCE tile"), (2, 3, "MMA conv issue span"), (4, 5, "MID"), (6, 7, "MMA 1x1 issue span"), (8, 9, "STORE after acc")]:
    print(f"  {label:24s} {float((t[5:n, b] - t[5:n, a]).float().mean()):8.0f} cycles")
print(f"  MMA warp waits per tile: A {float(t[5:n, 10].float().mean()):8.0f}  B {float(t[5:n, 11].float().mean()):8.0f}   producer team 0 blocked on free slot: {float(t[5:n, 12].float().mean()):8.0f}")
