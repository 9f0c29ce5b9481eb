#!/usr/bin/env python
"""Debug: per-stage timeline of the streamed-weight kernel on a plain (strided) conv with the SnakeBeta prologue,
as the EncoderBlock calls it (clock64 stamps of CTA 0; needs a BC_TRACE=1 build).
usage: stream_trace_conv.py C_in C_out K stride T_in [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import _cabi
from audiotokenization_b200.vq import module as M, activations

ci, co, k, s, T = (int(v) for v in sys.argv[1:6])
prec = sys.argv[6] if len(sys.argv) > 6 else "bf16x3"
B = 8
M.set_precision(prec)
pad = (s // 2 + s % 2) if s > 1 else (k - 1) // 2
conv = M.WNConv1d(ci, co, kernel_size=k, stride=s, padding=pad).cuda()
act = activations.SnakeBeta(ci, alpha_logscale=True).cuda()
xs = [torch.randn(B, T, ci, device="cuda") for _ in range(3)]
lib = _cabi.load_library()
for i in range(3):
    y = conv.forward_cl(xs[i], act=act)
torch.cuda.synchronize()
trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.bc_debug_set_stream_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = conv.forward_cl(xs[0], act=act); e1.record()
torch.cuda.synchronize()
lib.bc_debug_set_stream_trace(None)
t = trace.cpu().view(64, 16)
names = ["p_start", "p_done", "m7_go", "m7_iss", "mid_go", "mid_done", "m1_go", "m1_iss", "st_acc", "st_done", "waitA", "waitB", "p_blk"]
base = int(t[0, 0])
T_out = y.shape[1]
ntile = B * ((T_out + 127) // 128) * max(1, co // 256)
print(f"conv {ci}->{co} k{k} s{s} {prec}: kernel {e0.elapsed_time(e1)*1e3:.0f} us, tiles/CTA {ntile/148:.1f}")
print("tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(2, 14):
    print(f"{i:4d} " + " ".join((f"{int(t[i, j]) - base:9d}" if j < 10 else f"{int(t[i, j]):9d}") if int(t[i, j]) else "        -" for j in range(13)))
n = min(60, ntile // 148 - 1)
d = (t[6:n, 9] - t[5:n - 1, 9]).float()
print("steady-state cycles per tile (store_done deltas):", float(d.mean()))
for a, b, label in [(0, 1, "PRODUCE tile"), (2, 3, "MMA conv issue span"), (8, 9, "STORE after acc"), (1, 2, "p_done -> m_go"), (3, 8, "m_iss -> st_acc")]:
    print(f"  {label:24s} {float((t[5:n, b] - t[5:n, a]).float().mean()):8.0f} cycles")
print(f"  producer warp 0, per tile it produced a group of: load wait {float(t[5:n, 13].float().mean()):8.0f}  math+stores {float(t[5:n, 14].float().mean()):8.0f}  incl. warp sync {float(t[5:n, 15].float().mean()):8.0f}  of which LDS batch {float(t[5:n, 11].float().mean()):8.0f}")
print(f"  MMA warp waits per tile: A {float(t[5:n, 10].float().mean()):8.0f}  B {float(t[5:n, 11].float().mean()):8.0f}   producer team 0 blocked on free slot: {float(t[5:n, 12].float().mean()):8.0f}")
