#!/bin/bash
# Runs every bring-up group in its own process with a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for g in "$@"; do
  timeout 300 python tests/gpu_bringup.py $g > gpurun_out/bringup_$g.log 2>&1
  echo "group $g exit $?" | tee -a gpurun_out/bringup_summary.log
  tail -n 40 gpurun_out/bringup_$g.log
done
