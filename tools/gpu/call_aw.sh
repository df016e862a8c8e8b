#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -2)
timeout 100 python tools/profile_rerank.py | tail -1
timeout 100 python tools/time_rerank_params.py 2>&1 | tail -6
