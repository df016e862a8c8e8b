#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2y_pytest.log
cat gpurun_out/r2y_pytest.log
timeout 200 python tools/profile_rerank.py > gpurun_out/r2y_rerank.log 2>&1; tail -2 gpurun_out/r2y_rerank.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank5015_r2y.csv python tools/profile_rerank_50_15.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank_r2y.csv python tools/profile_rerank.py > /dev/null 2>&1
