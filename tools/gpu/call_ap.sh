#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank_r2ap.csv python tools/profile_rerank.py > /dev/null 2>&1
timeout 100 python tools/profile_rerank.py | tail -1
