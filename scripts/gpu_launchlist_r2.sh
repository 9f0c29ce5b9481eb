#!/bin/bash
# Round-2 evidence, ONE ncu pass per call (each only after the same command exited 0 without ncu):  $1 = tag, $2 = launches | dram | full | units
TAG=${1:-r02}; WHAT=${2:-launches}
PREC=bf16x3
mkdir -p gpurun_out
SMALL="env BC_LSTM_WAVEFRONT=0 python bench.py --clips-per-gpu 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --also= --precision $PREC"
case $WHAT in
launches)
  timeout 600 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv \
      --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launch_${TAG}.log 2>&1
  echo "ncu launches exit $?";;
dram)
  timeout 600 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 && \
  timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv \
      --log-file gpurun_out/dram_${TAG}.csv $SMALL > gpurun_out/ncu_dram_${TAG}.log 2>&1
  echo "ncu dram exit $?";;
units)
  timeout 300 python scripts/profile_units.py > gpurun_out/plain_units_${TAG}.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ru_group_kernel|ru_pair_kernel|conv_stream|lstm_tc_kernel" \
      --launch-skip 0 -c 16 -f -o gpurun_out/prof_units_${TAG} python scripts/profile_units.py > gpurun_out/ncu_units_${TAG}.log 2>&1
  echo "ncu units exit $?"; tail -3 gpurun_out/ncu_units_${TAG}.log;;
esac
ls -la gpurun_out/ | grep ${TAG}
