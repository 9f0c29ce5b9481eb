#!/bin/bash
# per-layer timing tables at a reduced shard (64 clips) for the given precisions
mkdir -p gpurun_out
for P in "$@"; do
  timeout 600 python bench.py --precision $P --clips-per-gpu 64 --steps 2 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/layers_$P.md > gpurun_out/bench_layers_$P.json 2> gpurun_out/bench_layers_$P.err; echo "bench $P exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_layers_$P.json')); print('$P', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'conv ms', round(d['roofline']['kernel_ms_per_step'],1), 'lstm ms', round(d['roofline']['lstm_ms_per_step'],1), 'step ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4))"
  cat gpurun_out/layers_$P.md
done
