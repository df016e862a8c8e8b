#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank5015_r2.csv python tools/profile_rerank_50_15.py > gpurun_out/launches_rerank5015_r2.log 2>&1
timeout 100 python tools/profile_rerank_50_15.py 2>&1 | tail -2
