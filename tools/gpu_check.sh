#!/bin/bash
# run each bring-up section in its own process with a watchdog; collect logs in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/check_gpu.txt
for s in "$@"; do
  timeout 300 python tools/gpu_check.py $s > gpurun_out/check_$s.log 2>&1
  echo "section $s exit $?" | tee -a gpurun_out/check_summary.txt
  tail -n 40 gpurun_out/check_$s.log
done
