#!/bin/bash
# ncu --set full captures of one eager step (B=16), exported to CSV on the box (reports can exceed the 64 MiB
# gpurun_out limit); per-instruction source pages for selected kernel instances.
mkdir -p gpurun_out
TAG=${1:-r1}
run() {  # name, kernel regex, skip, count, "ids for source export"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 \
      python tools/profile_step.py 16 2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/${TAG}_$1_details.csv 2>/dev/null
  for id in $5; do
    ncu -i /tmp/$1.ncu-rep --page source --csv --kernel-id ::$2:$id 2>/dev/null | cut -d, -f1-12 > gpurun_out/${TAG}_$1_source_$id.csv
  done
  echo "$1: report $(stat -c %s /tmp/$1.ncu-rep 2>/dev/null) bytes"
}
run halo "conv_halo_kernel" 34 34 "35 36 41 49 67 68"
run tc "conv_tc_kernel" 40 40 "41 61 73"
run decode "decode_kernel" 1 1 "2"
ls -la gpurun_out | tail -25
