#!/bin/bash
# box-side: GPU tests + bench (full JSON line incl. cpu baseline and other workloads)
tag=${1:-r2l}
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -8 gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json
