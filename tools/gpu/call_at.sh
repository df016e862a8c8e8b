#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2at_pytest.log
cat gpurun_out/r2at_pytest.log
for sp in 1 2 4; do
DEMO_JC_SPLIT=$sp timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank5015_r2at$sp.csv python tools/profile_rerank_50_15.py > /dev/null 2>&1
done
DEMO_JC_SPLIT=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank_r2at2.csv python tools/profile_rerank.py > /dev/null 2>&1
