#!/bin/bash
# box-side: GPU tests, stream probe, bench (round 2, call a)
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
(timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2a_pytest.log
(timeout 300 python tools/probe_stream.py --large) > gpurun_out/r2a_probe.log 2>&1
timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -20 gpurun_out/r2a_probe.log; head -c 600 gpurun_out/r2a_bench.json
