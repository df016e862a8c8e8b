#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ax_smoke.log 2>&1; tail -1 gpurun_out/r2ax_smoke.log
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2ax_pytest.log; cat gpurun_out/r2ax_pytest.log
timeout 900 python bench.py > gpurun_out/r2ax_bench.json 2> gpurun_out/r2ax_bench.err; echo bench rc=$?
