#!/bin/bash
# Default bench, then ncu launch list and one --set full capture of the dominant kernel
# (B200_PROFILING.md recipe).  $1 = tag for output names, $2 = precision (default fp32).
TAG=${1:-r01}
PREC=${2:-fp32}
mkdir -p gpurun_out
export BC_PRECISION=$PREC
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_${TAG}.json
SMALL="python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-cpu-baseline --precision $PREC"
timeout 600 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches exit $?"
KREG=${3:-conv1d_f32_kernel}
timeout 600 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 20 -c 3 \
    -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/
