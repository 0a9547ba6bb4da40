#!/bin/bash
# compute-sanitizer over the GPU tests: tools/sanitize.sh [memcheck|racecheck] [pytest -k expression]
tool=${1:-memcheck}; mkdir -p gpurun_out
if [ -n "$2" ]; then sel=(-k "$2"); else sel=(); fi
timeout 2400 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 30 python -m pytest tests -m gpu -q "${sel[@]}" > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"
tail -6 gpurun_out/sanitize_$tool.log | cut -c1-300
grep -m5 -n "Invalid\|out of bounds\|misaligned\|hazard\|Race" gpurun_out/sanitize_$tool.log | cut -c1-300
true
