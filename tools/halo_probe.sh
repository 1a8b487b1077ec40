#!/bin/bash
# bring-up of the swizzled halo mode: find which descriptor convention is right, then bench with it
mkdir -p gpurun_out
for v in "1" "0"; do
  CTB_HALO_BASEOFF=$v timeout 300 python tools/gpu_check.py conv_halo > gpurun_out/halo_baseoff_$v.log 2>&1
  rc=$?
  echo "swizzled halo, base_offset=$v -> rc=$rc" | tee -a gpurun_out/halo_probe.txt
  grep -c " OK" gpurun_out/halo_baseoff_$v.log; grep "FAIL" gpurun_out/halo_baseoff_$v.log | cut -c1-150 | head -5
  if [ $rc -eq 0 ]; then export CTB_HALO_BASEOFF=$v; break; fi
done
if [ $rc -ne 0 ]; then echo "falling back to planes" | tee -a gpurun_out/halo_probe.txt; export CTB_HALO_MODE=planes; fi
timeout 300 python tools/gpu_check.py net_bf16_emu > gpurun_out/check_net_bf16_emu.log 2>&1; tail -22 gpurun_out/check_net_bf16_emu.log | cut -c1-150
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json; tail -3 gpurun_out/bench_r1d.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
