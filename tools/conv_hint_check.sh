#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_ingest.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-train --no-cpu-baseline --no-procs --steps 40 > gpurun_out/bench_hint.json 2> gpurun_out/bench_hint.err; tail -2 gpurun_out/bench_hint.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_hint.json').read().strip().splitlines()[-1])
ro = d['roofline_other']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'conv ms', round(d['roofline']['ms_per_launch'], 4), 'stft ms', round(ro[0]['ms_per_launch'], 4), 'unet ms', round(ro[2]['ms'], 4))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:"m2h_rir_conv24|m2h_stft" -s 4 -c 4 --csv --log-file gpurun_out/hint_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-procs > gpurun_out/hint_traffic.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/hint_traffic.csv')))
hdr = None
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    print(d['ID'], d['Kernel Name'][:30], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
