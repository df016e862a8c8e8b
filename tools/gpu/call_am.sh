#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_msrv.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2am_pytest.log
cat gpurun_out/r2am_pytest.log
timeout 300 python tools/bench_hbm_kernels.py 2>&1 | grep topk | tee gpurun_out/r2am_hbm.log
