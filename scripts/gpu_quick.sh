#!/bin/bash
# quick TC iteration: ops tests (tensor-core + fused) then layer tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tensor_core or fused or lstm" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/pytest_quick.log | cut -c1-400
bash scripts/gpu_layers.sh "$@"
