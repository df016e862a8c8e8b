#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -4) > gpurun_out/r2an_pytest.log
cat gpurun_out/r2an_pytest.log
