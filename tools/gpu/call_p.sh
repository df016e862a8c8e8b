#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_pytest.log
timeout 200 python tools/profile_rerank.py > gpurun_out/r2p_rerank.log 2>&1; tail -2 gpurun_out/r2p_rerank.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_rerank_r2p.csv python tools/profile_rerank.py > /dev/null 2>&1
timeout 200 python tools/diag_eval_auto.py 2>&1 | tail -5
timeout 300 python tools/bench_hbm_kernels.py 2>&1 | tail -6
timeout 300 python - 2>&1 <<'PY' | tail -2
import sys, torch
sys.path.insert(0, '.')
import bench
def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out
print(bench.large_r171(torch.device('cuda'), timed))
PY
