#!/bin/bash
# data-parallel pretraining step on N GPUs, both reduction modes: tools/ddp_iter.sh <N> <tag>
N=${1:-2}; tag=${2:-ddp}
mkdir -p gpurun_out
for mode in ${3:-after_backward one_group}; do
  M2H_DDP_MODE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-procs > gpurun_out/bench_${tag}_${N}gpu_$mode.json 2> gpurun_out/bench_${tag}_${N}gpu_$mode.err
  tail -2 gpurun_out/bench_${tag}_${N}gpu_$mode.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_${N}gpu_$mode.json').read().strip().splitlines()[-1])
t = d['train_step']
print('$mode N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), 'train ms/step', round(t['ms_per_step'], 3), 'median', round(t['ms_per_step_median_this_rank'], 3), 'losses', t['bin_mono_loss_first_last'], 'pcie', d['pcie_probe'])
PY
done
