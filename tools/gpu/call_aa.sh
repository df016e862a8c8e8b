#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/profile_rerank.py > gpurun_out/r2aa_rerank.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'topk_rows|krecip|jaccard|expand_kernel' -s 8 -c 4 \
  -o gpurun_out/prof_rerank_r2aa -f python tools/profile_rerank.py > gpurun_out/r2aa_ncu.log 2>&1
tail -3 gpurun_out/r2aa_rerank.log; tail -3 gpurun_out/r2aa_ncu.log
