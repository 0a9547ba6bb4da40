#!/bin/bash
# U-Net forward iteration on one GPU: tools/unet_iter.sh <tag> [ncu]
tag=${1:-ui}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/unet_fwd_probe.py 1024 50
M2H_SLAB_OLD_RESIDENT_RULE=1 timeout 300 python tools/unet_fwd_probe.py 1024 50
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none --launch-skip 33 --launch-count 11 --csv --log-file gpurun_out/unet_layers_${tag}.csv python tools/unet_fwd_probe.py 1024 1 > gpurun_out/unet_layers_${tag}.log 2>&1
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/unet_layers_${tag}.csv')))
hdr = None; cur = {}
out = {}
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    out.setdefault(d['ID'], {'k': d['Kernel Name'][:34], 'g': d['Grid Size']})[d['Metric Name'].split('.')[0][-28:]] = d['Metric Value']
for i, v in out.items(): print(i, v)
PY
fi
