#!/bin/bash
# Probe the tcgen05 conv kernel: run the tensor-core tests under each descriptor variant.
mkdir -p gpurun_out
for V in 0 1 2 3; do
  echo "=== BC_TC_VARIANT=$V" 
  BC_TC_VARIANT=$V timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tensor_core_modes and 32-32-7-1-1-3-777" > gpurun_out/tc_probe_v$V.log 2>&1
  echo "exit $?"; tail -5 gpurun_out/tc_probe_v$V.log | cut -c1-400
done
