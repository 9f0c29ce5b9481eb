#!/bin/bash
# ncu --set full of one launch of each hot kernel at bench shapes (after the same command exited 0 without ncu).  $1 = tag
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python scripts/profile_units.py > gpurun_out/plain_units_${TAG}.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ru_group_kernel|ru_pair_kernel|conv_stream|lstm_tc_kernel" \
    --launch-skip 0 -c 14 -f -o gpurun_out/prof_units_${TAG} python scripts/profile_units.py > gpurun_out/ncu_units_${TAG}.log 2>&1
echo "ncu units exit $?"; tail -3 gpurun_out/ncu_units_${TAG}.log
