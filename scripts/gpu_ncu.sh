#!/bin/bash
# ncu --set full on one kernel family of the small bench.  usage: gpu_ncu.sh <tag> <precision> <kernel-regex> [skip] [count]
TAG=$1; PREC=$2; KREG=$3; SKIP=${4:-6}; CNT=${5:-3}
mkdir -p gpurun_out
SMALL="python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-cpu-baseline --precision $PREC"
timeout 600 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$KREG -s $SKIP -c $CNT \
    -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
