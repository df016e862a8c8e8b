#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_eval.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py tests/test_gpu_msrv.py tests/test_gpu_rerank.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2ao_pytest.log
cat gpurun_out/r2ao_pytest.log
timeout 300 python tools/bench_hbm_kernels.py 2>&1 | grep -A1 "evaluate_matrix" | tee gpurun_out/r2ao_hbm.log
