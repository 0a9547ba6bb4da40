#!/bin/bash
# what the driver runs at round end, on one GPU: tools/final_check.sh <tag>
tag=${1:-final}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err ) 2>&1 | grep real
tail -2 gpurun_out/bench_${tag}.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'roofline', d['roofline']['frac'], 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'], 'train', (d.get('train_step') or {}).get('ms_per_step'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print({k: (v if not isinstance(v, (dict, list)) else '...') for k, v in d.items()})
PY
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/bench_${tag}_ref.json
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-procs --no-train > gpurun_out/launches_${tag}.log 2>&1
  tail -1 gpurun_out/launches_${tag}.log | cut -c1-200
fi
