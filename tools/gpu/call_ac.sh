#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2ac_pytest.log
cat gpurun_out/r2ac_pytest.log
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r2ac_hbm.log 2>&1; grep -c . gpurun_out/r2ac_hbm.log
timeout 900 python bench.py > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; tail -c 600 gpurun_out/r2ac_bench.err
