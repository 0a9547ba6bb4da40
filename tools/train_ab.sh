#!/bin/bash
# A/B of the pretraining step variants on one GPU: tools/train_ab.sh <tag>
tag=${1:-tab}
mkdir -p gpurun_out
for v in 0 1; do
  M2H_TRAIN_TWO_STREAMS=$v timeout 600 python bench.py --no-cpu-baseline --no-procs --steps 10 > gpurun_out/bench_${tag}_ts$v.json 2> gpurun_out/bench_${tag}_ts$v.err
  tail -2 gpurun_out/bench_${tag}_ts$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_ts$v.json').read().strip().splitlines()[-1])
t = d['train_step']
print('TWO_STREAMS=$v', t['mode'], 'ms/step', round(t['ms_per_step'], 3), 'eager', round(t['ms_per_step_eager_this_rank'], 3), 'losses', t['bin_mono_loss_first_last'], t['graph_capture_error'])
PY
done
