#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 0 100 60 16; do
  echo "MIN_CTAS=$v"; M2H_TMA_BN256_MIN_CTAS=$v timeout 300 python tools/unet_fwd_probe.py 1024 60
done
M2H_TMA_BN256_MIN_CTAS=16 timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -m gpu -x -q 2>&1 | tail -2
