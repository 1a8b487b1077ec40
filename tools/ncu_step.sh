#!/bin/bash
# Final-evidence captures for profiles/ (run on the GPU box):  bash tools/ncu_step.sh <tag> [frames per step]
#  1. launch list (gpu__time_duration.sum) of the bench command itself -> <tag>_launches_bench.csv
#  2. ncu --set full of ONE eager step (tools/profile_step.py, cudaProfilerStart/Stop) exported to CSV on the box (the
#     .ncu-rep exceeds the 64 MiB gpurun_out limit) -> <tag>_ncu_step_{raw,details}.csv + the op list of the step
mkdir -p gpurun_out
TAG=${1:-r02f}
B=${2:-32}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-parity --batch $B > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/${TAG}_step \
    python tools/profile_step.py --batch $B > gpurun_out/${TAG}_profile_step_ops.txt 2> gpurun_out/${TAG}_ncu_step.log
echo "full capture rc=$? $(stat -c %s /tmp/${TAG}_step.ncu-rep 2>/dev/null) bytes"
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_step_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_step.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_step_details.csv 2>/dev/null
ls -la gpurun_out | tail -8
