"""Small fused evaluation for ncu captures (count GEMM ~ few ms):  python tools/profile_eval.py [Q G d]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics  # noqa: E402

Q, G, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 131072, 1536)
nid = max(2, G // 20)
rng = np.random.default_rng(0)
qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
gen = torch.Generator(device="cuda").manual_seed(0)
centers = torch.randn(nid, d, device="cuda", generator=gen)
qf = centers[torch.from_numpy(qp).cuda()] + 4 * torch.randn(Q, d, device="cuda", generator=gen)
gf = centers[torch.from_numpy(gp).cuda()] + 4 * torch.randn(G, d, device="cuda", generator=gen)
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    e1.record()
    torch.cuda.synchronize()
    print("iter %d: %.3f ms  mAP %.5f R1 %.4f  (%.1f TFLOP/s algorithmic incl. all stages)"
          % (it, e0.elapsed_time(e1), res.mAP, res.cmc[0], 2.0 * Q * G * d / e0.elapsed_time(e1) * 1e-9))
