#!/usr/bin/env python
"""Debug / profiling: one VQ sweep point (K = 8192, 1 Mi frames) of the fused factorized-VQ kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import bench_configs
print(json.dumps(bench_configs.vq_sweep(frames=1 << 20, sizes=(8192,), reps=3)))
