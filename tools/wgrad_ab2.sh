#!/bin/bash
tag=${1:-wg}
mkdir -p gpurun_out
for v in ${2:-50 148 296}; do
  M2H_WGRAD_TARGET_CTAS_32=$v timeout 600 python bench.py --no-cpu-baseline --no-procs --steps 10 > gpurun_out/bench_${tag}_t$v.json 2> gpurun_out/bench_${tag}_t$v.err
  tail -1 gpurun_out/bench_${tag}_t$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_t$v.json').read().strip().splitlines()[-1])
t = d['train_step']
print('TARGET_CTAS_32=$v', t['mode'], 'ms/step', round(t['ms_per_step'], 3), 'losses', t['bin_mono_loss_first_last'])
PY
done
