#!/usr/bin/env python
"""Debug: time the tensor-core LSTM recurrence alone (H = 512, T = 2400) for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops
from audiotokenization_b200.vq import module as M

H, T = 512, 2400
for prec in sys.argv[1:] or ("bf16x3",):
    lstm = M.ResLSTM(H, num_layers=1).cuda()
    img = lstm.lstm.recurrent_image_for(0, prec)
    mb = ops.lstm_tc_max_batch(H, prec)
    out = []
    for B in (1, 128, 256, 512):
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
        out.append(f"B={B}: {ms / T * 1e3:6.2f} us/step")
        del pre, y
    print(os.environ.get("BC_LIB_PATH", "default"), prec, " | ".join(out))
