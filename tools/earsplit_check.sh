#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_ingest.py tests/test_graphs_gpu.py tests/test_transport.py tests/test_bench_shapes_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/shim_profile.py 2>&1 | head -2
M2H_CONV_EAR_SPLIT=0 timeout 300 python tools/shim_profile.py 2>&1 | head -2
for v in 1 0; do
  M2H_CONV_EAR_SPLIT=$v timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/bench_es$v.json 2> gpurun_out/bench_es$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_es$v.json').read().strip().splitlines()[-1])
print('EAR_SPLIT=$v value', round(d['value']), 'rollout_step graph ms', round(d['rollout_step']['gpu_ms_cuda_graph'], 4), 'eager', round(d['rollout_step']['gpu_ms_eager'], 4), 'transport ms', round(d['rollout_transport']['latency_ms_per_step'], 4), 'procs env-steps/s', round(d['rollout_processes']['env_steps_per_s']), 'lat', round(d['rollout_processes']['latency_ms_per_step_mean'], 3))
PY
done
