#!/bin/bash
# box-side: GPU tests + bench (round 2, call a)
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2a_pytest.log
timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -8 gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
