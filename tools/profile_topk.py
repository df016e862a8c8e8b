"""topk_rows on the shapes of profiles/hbm_kernels_*.txt, for ncu: python tools/profile_topk.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import reranking  # noqa: E402

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
for R, C, pitch, k in ((4096, 262144, 262144, 50), (10290, 10290, 10304, 21), (10290, 10290, 10304, 51)):
    m = torch.rand(R, pitch, device=dev, generator=gen)[:, :C]
    for _ in range(3):
        idx = reranking.topk_rows(m, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx = reranking.topk_rows(m, k)
    e1.record()
    torch.cuda.synchronize()
    print("topk_rows %d x %d k=%d: %.3f ms, %.0f GB/s" % (R, C, k, e0.elapsed_time(e1), R * C * 4 / e0.elapsed_time(e1) * 1e-6))
    del m
