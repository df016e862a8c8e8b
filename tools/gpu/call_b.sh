#!/bin/bash
# box-side: round-2 profiles (launch list of the bench, ncu --set full of the count GEMM, re-ranking launch list, HBM kernels)
mkdir -p gpurun_out
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels_r2a.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2a.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/launches_r2a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sqdist_gemm2_kernel -s 1 -c 1 -f -o gpurun_out/prof_count_r2a \
  python tools/exp_count.py 20000 1000000 1536 1 > gpurun_out/prof_count_r2a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank_r2a.csv python tools/profile_rerank.py > gpurun_out/launches_rerank_r2a.log 2>&1
cat gpurun_out/hbm_kernels_r2a.txt; tail -3 gpurun_out/prof_count_r2a.log; tail -3 gpurun_out/launches_rerank_r2a.log
