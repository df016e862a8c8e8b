#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ak_smoke.log 2>&1; tail -1 gpurun_out/r2ak_smoke.log
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2ak_pytest.log; cat gpurun_out/r2ak_pytest.log
timeout 900 python bench.py > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; echo bench rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ak_ref.json 2> gpurun_out/r2ak_ref.err; echo ref rc=$?; head -c 400 gpurun_out/r2ak_ref.json
