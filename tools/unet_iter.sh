#!/bin/bash
# U-Net iteration on the GPU box: parity tests, bench line, launch list (+ optional full capture of the U-Net kernels)
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_unet_gpu.py tests/test_unet_train_gpu.py -x -q 2>&1 | tail -4
timeout -k 10 300 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_iter.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'conv ms', d['roofline']['ms_per_launch'], 'stft ms', d['roofline_other'][0]['ms_per_launch'], 'unet ms', d['roofline_other'][2]['ms'], 'e2e', d['e2e']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"unet_|pack_input|clear_halo|head" -s 28 -c 30 --csv --log-file gpurun_out/launches_unet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_unet.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); ii=h.index('ID'); ig=h.index('Grid Size')
cur={}
for r in rows[1:]:
    cur.setdefault(r[ii],[r[ik][:28],r[ig]]).append(r[iv])
for k,v in list(cur.items())[:26]: print(k,v)
PY
if [ "$1" == "full" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unet_ -s 22 -c 11 -o gpurun_out/prof_unet -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_unet.log 2>&1
fi
