#!/bin/bash
# 2-GPU check of the long-form peer hand-off: bit-identical to the 1-GPU result, timing.  usage: gpurun --gpus 2 -- bash scripts/gpu_longform2.sh
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_longform.py --minutes 2 --chunk-seconds 15 --check-whole 2>&1 | grep -v "^W\|warn" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_longform.py --minutes 10 2>&1 | grep -v "^W\|warn" | tail -3
