#!/bin/bash
# same-box A/B of an environment switch on the pretraining step: tools/train_env_ab.sh VAR "v1 v2 ..." [repeats]
var=$1; vals=$2; rep=${3:-2}; mkdir -p gpurun_out
for r in $(seq 1 $rep); do for v in $vals; do
  env $var=$v timeout 600 python bench.py --no-cpu-baseline --no-procs --steps 10 > gpurun_out/bench_envab.json 2> gpurun_out/bench_envab.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_envab.json').read().strip().splitlines()[-1])
print('$var=$v', 'train ms/step', round(d['train_step']['ms_per_step'], 3), 'value', round(d['value']))
PY
done; done
