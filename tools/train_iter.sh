#!/bin/bash
# pretraining-step iteration on one GPU: the train tests + the two-stream bench leg.  tools/train_iter.sh <tag>
tag=${1:-ti}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_shapes_gpu.py tests/test_unet_train_gpu.py tests/test_render_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in 1; do
  M2H_TRAIN_TWO_STREAMS=$v timeout 600 python bench.py --no-cpu-baseline --no-procs --steps 10 > gpurun_out/bench_${tag}_ts$v.json 2> gpurun_out/bench_${tag}_ts$v.err
  tail -2 gpurun_out/bench_${tag}_ts$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_ts$v.json').read().strip().splitlines()[-1])
t = d['train_step']
print('TWO_STREAMS=$v', t['mode'], 'ms/step', round(t['ms_per_step'], 3), 'eager', round(t['ms_per_step_eager_this_rank'], 3), 'losses', t['bin_mono_loss_first_last'], t['graph_capture_error'])
print('value', d['value'], 'ms/step', d['ms_per_step'])
PY
done
