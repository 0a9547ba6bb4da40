#!/bin/bash
# launch list + full capture of the U-Net kernels of one bench step
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_unet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:unet_ -s 22 -c 11 -o gpurun_out/prof_unet -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_unet.log 2>&1
tail -3 gpurun_out/ncu_unet.log
