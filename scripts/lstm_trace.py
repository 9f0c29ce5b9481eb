#!/usr/bin/env python
"""Debug: time the tensor-core LSTM recurrence alone for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops
from audiotokenization_b200.vq import module as M

H, T = 512, 2400
for prec in ("bf16x3",):
    lstm = M.ResLSTM(H, num_layers=1).cuda()
    img = lstm.lstm.recurrent_image_for(0, prec)
    mb = ops.lstm_tc_max_batch(H, prec)
    for B in (1, 64, 256):
        if B > mb:
            continue
        pre = torch.randn(B, T, 4 * H, device="cuda") * 0.5
        for _ in range(2):
            y = ops.lstm_recurrent_tc(pre, img, None, prec, mb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = ops.lstm_recurrent_tc(pre, img, None, prec, mb); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{prec:7s} B={B:4d}: {ms:7.2f} ms  {ms / T * 1e3:6.2f} us/step  max_batch {mb}")
        del pre, y

# per-phase timeline of CTA (0,0), steps 100..163 (B = 256, split precision)
from audiotokenization_b200 import _cabi
lib = _cabi.load_library()
for prec, B in (("bf16x3", 1), ("bf16x3", 64), ("bf16x3", 256)):
    lstm = M.ResLSTM(H, num_layers=1).cuda()
    img = lstm.lstm.recurrent_image_for(0, prec)
    mb = ops.lstm_tc_max_batch(H, prec)
    pre = torch.randn(B, T, 4 * H, device="cuda") * 0.5
    ops.lstm_recurrent_tc(pre, img, None, prec, mb)
    torch.cuda.synchronize()
    trace = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.bc_debug_set_lstm_trace(trace.data_ptr())
    ops.lstm_recurrent_tc(pre, img, None, prec, mb)
    torch.cuda.synchronize()
    lib.bc_debug_set_lstm_trace(None)
    t = trace.cpu().view(64, 8).double()
    names = ["counter seen", "h copies issued", "MMAs issued", "gates start (acc ready)", "h stored+fenced", "published", "(pair: polls of chunk 0)", "(pair: first arrival of chunk 0 seen)"]
    step = (t[1:, 5] - t[:-1, 5]).mean()
    print(f"{prec} B={B}: {step:.0f} cycles per step; phase offsets from the previous step's publish:")
    for j, n in enumerate(names):
        print(f"   {n:28s} {float((t[1:, j] - t[:-1, 5]).mean()):8.0f}" if j != 6 else f"   {n:28s} {float(t[1:, j].mean()):8.1f}")
