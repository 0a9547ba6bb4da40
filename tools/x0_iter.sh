#!/bin/bash
tag=${1:-x0}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_shapes_gpu.py tests/test_ops.py tests/test_unet_gpu.py tests/test_graphs_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in "" 1; do
  M2H_BENCH_FP32_OBS=$v timeout 600 python bench.py --no-train --no-cpu-baseline --no-procs > gpurun_out/bench_${tag}_fp32obs$v.json 2> gpurun_out/bench_${tag}_fp32obs$v.err
  tail -2 gpurun_out/bench_${tag}_fp32obs$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_fp32obs$v.json').read().strip().splitlines()[-1])
ro = d['roofline_other']
print('FP32_OBS=$v value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'conv ms', round(d['roofline']['ms_per_launch'], 4), 'stft ms', round(ro[0]['ms_per_launch'], 4), 'unet ms', round(ro[2]['ms'], 4), 'e2e', round(d['e2e']['value']))
PY
done
