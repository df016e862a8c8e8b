#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r2n_other.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda')
peaks = bench.load_peaks()
for rep in range(2):
    o = bench.other_workloads(dev, peaks)
    print(rep, {k: round(v['ms'], 4) for k, v in o.items() if isinstance(v, dict) and 'ms' in v})
PY
cat gpurun_out/r2n_other.log | tail -5
