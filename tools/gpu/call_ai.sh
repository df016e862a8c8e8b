#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_triplet.py tests/test_gpu_batch_losses.py tests/test_gpu_center_loss.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2ai_pytest.log
cat gpurun_out/r2ai_pytest.log
