#!/bin/bash
# GPU check of the U-Net inference + training paths
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_gpu.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_unet_train_gpu.py -q -s 2>&1 > gpurun_out/train_tests.log; grep -E "^E  |passed|failed|^{" gpurun_out/train_tests.log | cut -c1-3000 | head -40
