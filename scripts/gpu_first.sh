#!/bin/bash
# First GPU session: parity tests, smoke, a small bench, then the default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --clips-per-gpu 16 --steps 2 --warmup 3 --cpu-sample-clips 1 > gpurun_out/bench_small.log 2>&1; echo "bench small exit $?" >> gpurun_out/bench_small.log
tail -3 gpurun_out/bench_small.log
