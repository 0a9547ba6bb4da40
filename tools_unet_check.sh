#!/bin/bash
# GPU check of the U-Net kernels: parity tests under both A-descriptor base-offset conventions, then a bench line.
mkdir -p gpurun_out
for mode in 1 0; do
  echo "=== M2H_SLAB_BASE_OFFSET=$mode"
  M2H_SLAB_BASE_OFFSET=$mode timeout 180 python -m pytest tests/test_unet_gpu.py -x -q 2>&1 | tail -15
done
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_quick.json'))
    print(d['value'], d['ms_per_step'], [(r['kernel'][:20], round(r.get('ms', r.get('ms_per_launch', 0)), 3)) for r in [d['roofline']] + d['roofline_other']], d['e2e'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
