"""evaluate_matrix on short rows: arithmetic-bin kernel vs bisection kernel (DEMO_CM_SMALL_G=<columns>)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics  # noqa: E402

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)


def timed(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("DEMO_CM_SMALL_G =", os.environ.get("DEMO_CM_SMALL_G"))
for Q, G, per_id in ((836, 836, 4), (1715, 8575, 20), (1715, 8575, 170), (4096, 16384, 20), (4096, 32768, 20),
                     (4096, 32768, 170), (4096, 65536, 20)):
    rng = np.random.default_rng(0)
    nid = max(2, G // per_id)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
    dist = torch.rand(Q, G, device=dev, generator=gen)
    plan = metrics.RankPlan(qp, gp, qc, gc)
    ms_u = timed(lambda: metrics.evaluate_matrix(dist, plan=plan))
    r_u = metrics.evaluate_matrix(dist, plan=plan)
    same = torch.from_numpy(qp).to(dev)[:, None] == torch.from_numpy(gp).to(dev)[None, :]
    dist = torch.where(same, dist * (2.0 * per_id / G), dist)
    ms_r = timed(lambda: metrics.evaluate_matrix(dist, plan=plan))
    r_r = metrics.evaluate_matrix(dist, plan=plan)
    print("%5d x %6d ~%3d per id: uniform %.4f ms (mAP %.6f)  retrieval-like %.4f ms (mAP %.6f)"
          % (Q, G, per_id, ms_u, r_u.mAP, ms_r, r_r.mAP))
