#!/bin/bash
# full ncu capture of one U-Net kernel launch: tools/unet_ncu.sh <tag> <kernel-regex> <launch-skip>
tag=$1; mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$2 --launch-skip $3 --launch-count 1 -f -o gpurun_out/unet_${tag} python tools/unet_fwd_probe.py 1024 1 > gpurun_out/unet_${tag}.log 2>&1
tail -3 gpurun_out/unet_${tag}.log
ls -la gpurun_out/unet_${tag}.ncu-rep
