#!/bin/bash
# Evidence captures for profiles/ (run on the GPU box):  bash tools/ncu_capture.sh <tag> [frames per step]
#  1. launch list (gpu__time_duration) of the bench command itself
#  2. ncu --set full of one eager step's conv_tc / conv_halo / decode launches, exported to CSV on the box
#     (the .ncu-rep files can exceed the 64 MiB gpurun_out limit)
mkdir -p gpurun_out
TAG=${1:-r1}
B=${2:-32}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
run() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 \
      python tools/profile_step.py $B 2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_$1_details.csv 2>/dev/null
  echo "$1: report $(stat -c %s /tmp/$1.ncu-rep 2>/dev/null) bytes"
}
run tc "conv_tc_kernel" 40 40
run halo "conv_halo_kernel" 34 34
run decode "decode_kernel" 1 1
ls -la gpurun_out | tail -12
