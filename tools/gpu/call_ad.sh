#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r2ad_other.log 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
peaks = bench.load_peaks() if hasattr(bench, 'load_peaks') else json.load(open('MEASURED_PEAKS.json'))
ow = bench.other_workloads(torch.device('cuda'), peaks)
for k, v in ow.items():
    if 'rerank' in k:
        print(k, json.dumps({a: b for a, b in v.items() if a != 'cpu_baseline'}))
PY
tail -5 gpurun_out/r2ad_other.log | cut -c1-1500
