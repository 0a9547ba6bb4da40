#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_train.json'))
    print(d['value'], d['ms_per_step'], d['train_step'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_train.err').read()[-3000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1200 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log
