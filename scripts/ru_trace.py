#!/usr/bin/env python
"""Debug: per-stage timeline of the persistent ResidualUnit kernel (clock64 stamps of CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import _cabi, ops
from audiotokenization_b200.vq import module as M

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dil = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B, T = 8, 480000 * 32 // C
ru = M.ResidualUnit(C, dilation=dil).cuda()
x = torch.randn(B, T, C, device="cuda")
M.set_precision(prec)
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
t = trace.cpu().view(64, 16)
names = ["ld_start", "ld_slot", "ld_done", "m7_go", "m7_iss", "m1_go", "m1_iss", "mid_go", "mid_done", "st_start", "st_acc", "st_done"]
base = int(t[0, 0])
print(f"C={C} {prec} dil={dil}: kernel {e0.elapsed_time(e1)*1e3:.0f} us, tiles/CTA {B*((T+127)//128)/148:.1f}")
print("tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(2, 26):
    print(f"{i:4d} " + " ".join(f"{int(t[i, j]) - base:9d}" if int(t[i, j]) else "        -" for j in range(12)))
d = (t[20:60, 11] - t[19:59, 11]).float()
print("steady-state cycles per tile (store_done deltas):", float(d.mean()))
for a, b, label in [(0, 2, "LOAD total"), (0, 1, "LOAD until slot+data"), (1, 2, "LOAD convert"), (3, 4, "MMA7 issue"), (7, 8, "MID"), (9, 11, "STORE total"), (10, 11, "STORE after acc"), (10, 12, "STORE tmem_load"), (12, 11, "STORE transpose+stores")]:
    print(f"  {label:24s} {float((t[10:60, b] - t[10:60, a]).float().mean()):8.0f} cycles")
