#!/bin/bash
# build + run the activation-throughput probe (on the GPU box: gpurun -- 'bash scripts/probes/act_probe.sh')
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o act_probe act_probe.cu && ./act_probe
