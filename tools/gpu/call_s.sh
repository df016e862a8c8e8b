#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_sharded.py tests/test_gpu_rerank.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2s_pytest.log
cat gpurun_out/r2s_pytest.log
