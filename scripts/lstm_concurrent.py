#!/usr/bin/env python
"""Experiment: 512 rows of the LSTM recurrence as ONE launch (two tiles per CTA) against TWO concurrent launches of 256 rows on two
streams (run with BC_LSTM_PAIR=2: the CTA-pair kernel takes 64 CTAs per 256 rows, so two launches fit the chip side by side)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops, _cabi
from audiotokenization_b200.vq import module as M

H, T, prec = 512, 2400, "bf16x3"
print(_cabi.policy())
lstm = M.ResLSTM(H, num_layers=1).cuda()
img = lstm.lstm.recurrent_image_for(0, prec)
pre = torch.randn(512, T, 4 * H, device="cuda") * 0.5
mb = ops.lstm_tc_max_batch(H, prec)

def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

ms, y512 = timed(lambda: ops.lstm_recurrent_tc(pre, img, None, prec, mb))
print(f"one launch of 512 rows ({ops.lstm_tc_ctas(512, H, prec)} CTAs): {ms:.2f} ms = {ms / T * 1e3:.2f} us/step")
ms, y256 = timed(lambda: ops.lstm_recurrent_tc(pre[:256], img, None, prec, 256))
print(f"one launch of 256 rows ({ops.lstm_tc_ctas(256, H, prec)} CTAs): {ms:.2f} ms = {ms / T * 1e3:.2f} us/step")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        a = ops.lstm_recurrent_tc(pre[:256], img, None, prec, 256)
    with torch.cuda.stream(s2):
        b = ops.lstm_recurrent_tc(pre[256:], img, None, prec, 256)
    cur.wait_stream(s1); cur.wait_stream(s2)
    return a, b
ms, (a, b) = timed(two)
print(f"two concurrent launches of 256 rows: {ms:.2f} ms = {ms / T * 1e3:.2f} us/step of 512 rows; identical to the single launch: {torch.equal(torch.cat([a, b]), y512)}")
