#!/usr/bin/env python
"""BASELINE.json configs[4]: VQ-heavy sweep -- codebook sizes 8192/16384/32768 at codebook_dim 8, isolating the
fused factorized-VQ kernel (bench_configs.vq_sweep): achieved HBM GB/s (2048 B in + 4 B out per frame) and FP32 TFLOP/s
(2*C*D + 2*K*D per frame) against the measured peaks, plus exactness against the reference formula on a sample."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs

if __name__ == "__main__":
    for margin in (False, True):
        print(json.dumps(bench_configs.vq_sweep(margin=margin)))
