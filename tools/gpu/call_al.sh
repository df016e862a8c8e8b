#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/profile_topk.py > gpurun_out/r2al_topk.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'topk_rows' -s 3 -c 1 \
  -o gpurun_out/prof_topk_r2al_long -f python tools/profile_topk.py > gpurun_out/r2al_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'topk_rows' -s 7 -c 1 \
  -o gpurun_out/prof_topk_r2al_mid -f python tools/profile_topk.py > gpurun_out/r2al_ncu2.log 2>&1
cat gpurun_out/r2al_topk.log; tail -2 gpurun_out/r2al_ncu2.log
