#!/bin/bash
# Run on the GPU box through gpurun: GPU tests, bench line, ncu launch list, ncu --set full of the main kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:m2h_rir_conv -s 2 -c 1 -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:m2h_stft -s 2 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:unet_gemm -s 20 -c 10 -o gpurun_out/prof_unet -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_unet.log 2>&1
ls -la gpurun_out
