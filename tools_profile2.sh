#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:m2h_rir_conv -s 2 -c 1 -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:m2h_stft -s 2 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:unet_gemm -s 20 -c 10 -o gpurun_out/prof_unet -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_unet.log 2>&1
ls -la gpurun_out
