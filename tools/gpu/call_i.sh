#!/bin/bash
mkdir -p gpurun_out
for per in 170 20; do
echo "=== per_id $per DEMO_COUNT255_ALL=1"
env DEMO_COUNT255_ALL=1 timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"count_matrix(63|255)" -c 2 \
  python tools/profile_count_mid.py 4096 262144 $per 2>&1 | grep -E "count_matrix|gpu__time|wavefronts|inst_executed|issue_active|warps_active" | sed 's/(const float.*//' | tail -6
done > gpurun_out/r2i_count_mid.log 2>&1
cat gpurun_out/r2i_count_mid.log
