#!/bin/bash
# one `ncu --set full` capture each of the three dominant kernels of the final build (bench workload): tools/final_ncu.sh
mkdir -p gpurun_out
for k in m2h_rir_conv24 m2h_stft; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/final_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-procs > gpurun_out/final_ncu_$k.log 2>&1
  tail -1 gpurun_out/final_ncu_$k.log | cut -c1-120
done
ls -la gpurun_out/final_*.ncu-rep
