#!/usr/bin/env python
"""Profiling workload: ONE launch (after one warm-up) of each hot kernel at the bench's launch shapes (8 x 30 s clips per
launch; LSTM: 512 utterances, 200 steps).  Meant to run under `ncu --set full` (scripts/gpu_profile_units.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops
from audiotokenization_b200.vq import module as M, activations

M.set_precision("bf16x3")
torch.manual_seed(0)
clips = 8
for C, dil, T in ((32, 9, 480000), (64, 9, 240000), (128, 9, 60000), (256, 9, 12000 * 8)):
    m = M.ResidualUnit(C, dilation=dil).cuda()
    x = torch.randn(clips, T, C, device="cuda")
    for _ in range(2):
        y = m.forward_cl(x)
    torch.cuda.synchronize()
    del x, y, m
for ci, co, k, s, T in ((32, 64, 4, 2, 480000), (512, 512, 3, 1, 2400 * 64), (512, 2048, 1, 1, 2400 * 64)):
    pad = (s // 2 + s % 2) if s > 1 else (k - 1) // 2
    m = M.WNConv1d(ci, co, kernel_size=k, stride=s, padding=pad).cuda()
    act = activations.SnakeBeta(ci, alpha_logscale=True).cuda() if k > 1 else None
    x = torch.randn(clips if s > 1 else 1, T, ci, device="cuda")
    for _ in range(2):
        y = m.forward_cl(x, act=act)
    torch.cuda.synchronize()
    del x, y, m
H, T, B = 512, 200, 512
lstm = M.ResLSTM(H, num_layers=1).cuda()
img = lstm.lstm.recurrent_image_for(0, "bf16x3")
pre = torch.randn(B, T, 4 * H, device="cuda") * 0.5
for _ in range(2):
    y = ops.lstm_recurrent_tc(pre, img, None, "bf16x3", ops.lstm_tc_max_batch(H, "bf16x3"))
torch.cuda.synchronize()
print("ok")
