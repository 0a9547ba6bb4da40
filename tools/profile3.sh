#!/bin/bash
# Full round pass on the GPU box: GPU tests, smoke, bench line, ncu launch list, ncu --set full of render + U-Net kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_rir_conv -s 2 -c 1 -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2h_stft -s 2 -c 1 -o gpurun_out/prof_stft -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_stft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unet_ -s 22 -c 11 -o gpurun_out/prof_unet -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_unet.log 2>&1
ls -la gpurun_out
