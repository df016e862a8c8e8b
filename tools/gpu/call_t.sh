#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; tail -2 gpurun_out/r2t_smoke.log
(time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_ref.json 2> gpurun_out/r2t_ref.err) 2>&1 | tail -3
head -c 900 gpurun_out/r2t_ref.json
