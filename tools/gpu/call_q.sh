#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sqdist_gemm2_kernel" -s 2 -c 1 -f -o gpurun_out/prof_rerank_r2q \
  python tools/profile_rerank.py > gpurun_out/prof_rerank_r2q.log 2>&1
tail -2 gpurun_out/prof_rerank_r2q.log
