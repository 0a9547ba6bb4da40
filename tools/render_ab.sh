#!/bin/bash
# A/B of run-time selectable render variants on the GPU box (no ncu): tools/render_ab.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 3 4; do
  M2H_STFT_CTAS=$v timeout 600 python bench.py --no-train --no-cpu-baseline --no-procs > gpurun_out/bench_${tag}_stft$v.json 2> gpurun_out/bench_${tag}_stft$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_stft$v.json').read().strip().splitlines()[-1])
ro = d['roofline_other']
print('STFT_CTAS=$v value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'conv ms', round(d['roofline']['ms_per_launch'], 4), 'stft ms', round(ro[0]['ms_per_launch'], 4), 'unet ms', round(ro[2]['ms'], 4))
PY
done
M2H_STFT_CTAS=4 timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -x -q 2>&1 | tail -2
