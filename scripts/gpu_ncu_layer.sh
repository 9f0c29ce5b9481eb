#!/bin/bash
# ncu --set full of one layer micro-benchmark.  usage: gpu_ncu_layer.sh <tag> <precision> <bench_layer filter> <kernel regex> [env...]
TAG=$1; PREC=$2; FILT=$3; KREG=$4
mkdir -p gpurun_out
CMD="python scripts/bench_layer.py $PREC 8 $FILT"
timeout 300 $CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
