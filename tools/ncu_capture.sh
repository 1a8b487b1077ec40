#!/bin/bash
# ncu --set full captures of one eager step (B=16), exported to CSV on the box (reports can exceed the 64 MiB
# gpurun_out limit); keeps the .ncu-rep only when small.
mkdir -p gpurun_out
TAG=${1:-r1}
run() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 \
      python tools/profile_step.py 16 2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/${TAG}_$1_details.csv 2>/dev/null
  sz=$(stat -c %s /tmp/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -lt 30000000 ] && [ "$sz" -gt 0 ]; then cp /tmp/$1.ncu-rep gpurun_out/${TAG}_$1.ncu-rep; fi
  echo "$1: report $sz bytes"
}
run halo "conv_halo_kernel" 13 13
run tc_dcn "conv_tc_kernel" ${2:-100} 12      # a window of gather-engine launches in the DCN/neck part of step 2
run decode "decode_kernel" 1 1
ls -la gpurun_out | tail -12
