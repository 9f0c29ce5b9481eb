#!/bin/bash
# Round profile: full bench, then (each only after the plain run exited 0) the ncu launch list, the DRAM-bytes pass over every
# launch of the small bench and one --set full capture of the two persistent conv kernels.  $1 = tag, $2 = precision.
TAG=${1:-r01e}
PREC=${2:-bf16x3}
mkdir -p gpurun_out
timeout 900 python bench.py --precision $PREC --layer-table gpurun_out/layers_${TAG}.md > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
SMALL="python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-cpu-baseline --also= --precision $PREC"
timeout 600 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches exit $?"
timeout 600 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 && \
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/dram_${TAG}.csv $SMALL > gpurun_out/ncu_dram_${TAG}.log 2>&1
echo "ncu dram exit $?"
timeout 600 $SMALL > gpurun_out/plain3_${TAG}.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_stream_kernel|ru_persist_kernel|ru_group_kernel" -s 40 -c 4 \
    -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/ | grep ${TAG}
