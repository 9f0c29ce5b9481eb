#!/bin/bash
# Round-2 profile: full bench, then (each only after the plain run of the same command exited 0) the ncu launch list, the
# DRAM-bytes pass over every launch of the small bench, and --set full captures of the hot kernels.  $1 = tag.
TAG=${1:-r02}
PREC=bf16x3
mkdir -p gpurun_out
timeout 900 python bench.py --precision $PREC --layer-table gpurun_out/layers_${TAG}.md > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
SMALL="env BC_LSTM_WAVEFRONT=0 python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --also= --precision $PREC"
timeout 600 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches exit $?"
timeout 600 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 && \
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv \
    --log-file gpurun_out/dram_${TAG}.csv $SMALL > gpurun_out/ncu_dram_${TAG}.log 2>&1
echo "ncu dram exit $?"
timeout 600 $SMALL > gpurun_out/plain3_${TAG}.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_stream|ru_pair_kernel|ru_group_kernel" -s 30 -c 8 \
    -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit $?"
timeout 300 python scripts/aa_time.py > gpurun_out/plain_aa_${TAG}.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"snake_aa2" -s 3 -c 1 \
    -f -o gpurun_out/prof_aa_${TAG} python scripts/aa_time.py > gpurun_out/ncu_aa_${TAG}.log 2>&1
echo "ncu aa exit $?"
timeout 300 python scripts/vq_time.py > gpurun_out/plain_vq_${TAG}.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vq_scan" -s 2 -c 1 \
    -f -o gpurun_out/prof_vq_${TAG} python scripts/vq_time.py > gpurun_out/ncu_vq_${TAG}.log 2>&1
echo "ncu vq exit $?"
ls -la gpurun_out/ | grep ${TAG}
