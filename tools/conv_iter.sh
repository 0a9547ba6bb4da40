#!/bin/bash
# quick conv/STFT iteration on the GPU box: render parity tests, bench line without the training leg, ncu of the render kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_iter.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'conv ms', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'], 'stft ms', d['roofline_other'][0]['ms_per_launch'], 'unet ms', d['roofline_other'][2]['ms'], 'e2e', d['e2e']['value'])
PY
if [ "$1" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_rir_conv -s 2 -c 1 -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_stft -s 2 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_stft.log 2>&1
fi
