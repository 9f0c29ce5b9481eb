#!/bin/bash
# Full GPU test suite + benches in the given precisions.  usage: gpu_round.sh <tag> <prec...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/pytest_${TAG}.log | cut -c1-300
for P in "$@"; do
  timeout 900 python bench.py --precision $P > gpurun_out/bench_${TAG}_$P.json 2> gpurun_out/bench_${TAG}_$P.err; echo "bench $P exit $?"
  tail -c 2500 gpurun_out/bench_${TAG}_$P.json; tail -3 gpurun_out/bench_${TAG}_$P.err
done
