#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_msrv.py tests/test_gpu_rerank.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2h_pytest.log
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels_r2h.txt 2>&1
timeout 300 python - > gpurun_out/r2h_r171.log 2>&1 <<'PY'
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
cat gpurun_out/r2h_pytest.log gpurun_out/hbm_kernels_r2h.txt gpurun_out/r2h_r171.log
