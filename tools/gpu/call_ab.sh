#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_rerank.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2ab_pytest.log
cat gpurun_out/r2ab_pytest.log
timeout 300 python tools/bench_hbm_kernels.py 2>&1 | grep topk | tee gpurun_out/r2ab_hbm.log
DEMO_TOPK_NOREG=1 timeout 300 python tools/bench_hbm_kernels.py 2>&1 | grep "topk_rows 10290" | tee gpurun_out/r2ab_hbm_noreg.log
timeout 200 python tools/profile_rerank.py 2>&1 | tail -1
