#!/bin/bash
# ncu launch list over ONE full-size step of the bench (about 1250 launches, skipped past the warm-up steps), so the per-kernel shares
# can be compared with bench.py's own.  $1 = tag, $2 = precision
TAG=${1:-r01g}; PREC=${2:-bf16x3}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --also= --precision $PREC --layer-table gpurun_out/layers_${TAG}.md"
timeout 600 $CMD > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
timeout 1700 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1250 --csv \
    --log-file gpurun_out/launches_full_${TAG}.csv $CMD > gpurun_out/ncu_launch_full_${TAG}.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launch_full_${TAG}.log | cut -c1-200
