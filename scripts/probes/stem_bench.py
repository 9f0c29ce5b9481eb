import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiotokenization_b200.vq import module as M
conv = M.WNConv1d(1, 32, kernel_size=7, padding=3).cuda()
xs = [torch.randn(8, 480000, 1, device='cuda') for _ in range(3)]
for _ in range(3): y = conv.forward_cl(xs[0])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(9): y = conv.forward_cl(xs[i % 3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 9
print(f"stem 1->32 k7, 8 x 30 s: {ms*1e3:.1f} us, {8*480000*33*4/ms/1e6:.0f} GB/s")
