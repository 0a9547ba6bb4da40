#!/bin/bash
# config-4 per-process mode with and without CUDA MPS: tools/procs_probe.sh
mkdir -p gpurun_out
timeout 120 python tools/procs_probe.py "no MPS :"
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
if timeout 20 nvidia-cuda-mps-control -d 2> gpurun_out/mps_start.err; then
  sleep 1
  timeout 120 python tools/procs_probe.py "with MPS:" || echo "with MPS: failed or timed out"
  echo quit | timeout 20 nvidia-cuda-mps-control
else
  echo "MPS daemon unavailable:"; cat gpurun_out/mps_start.err
fi
