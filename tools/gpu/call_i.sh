#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q -x -k "64_to_255 or edge_cases or rgbnt100 or near_duplicate" 2>&1 | tail -5) > gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log
for per in 170 20; do for all in "" 1; do
echo "=== per_id $per DEMO_COUNT255_ALL=$all"
env ${all:+DEMO_COUNT255_ALL=1} timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:count_matrix -c 3 \
  python tools/profile_count_mid.py 4096 262144 $per 2>&1 | grep -E "count_matrix|gpu__time|bank_conflicts|wavefronts|inst_executed|issue_active|iter 2" | sed 's/(const float.*//'
done; done > gpurun_out/r2i_count_mid.log 2>&1
cat gpurun_out/r2i_count_mid.log
