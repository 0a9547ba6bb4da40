#!/bin/bash
# render-kernel iteration on the GPU box: parity tests, bench line (no train / cpu legs), optional ncu of conv + stft
# usage: tools/render_iter.sh <tag> [ncu]
tag=${1:-iter}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_ingest.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
ro = d['roofline_other']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'conv ms', round(d['roofline']['ms_per_launch'], 4), 'stft ms', round(ro[0]['ms_per_launch'], 4), 'unet ms', round(ro[2]['ms'], 4), 'e2e', round(d['e2e']['value']))
PY
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_rir_conv -s 2 -c 1 -o gpurun_out/prof_conv_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_stft -s 2 -c 1 -o gpurun_out/prof_stft_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_stft.log 2>&1
fi
