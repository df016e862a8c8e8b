"""evaluate_matrix on a materialised matrix with ~per_id positives per query, for ncu launch lists:
python tools/profile_count_mid.py [Q G per_id]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics  # noqa: E402

Q, G, per_id = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 262144, 170)
rng = np.random.default_rng(0)
nid = max(2, G // per_id)
qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
dist = torch.rand(Q, G, device="cuda")
plan = metrics.RankPlan(qp, gp, qc, gc)
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = metrics.evaluate_matrix(dist, plan=plan)
    e1.record()
    torch.cuda.synchronize()
    print("iter %d: %.3f ms  mAP %.5f  max positives %d" % (it, e0.elapsed_time(e1), res.mAP, plan.max_cnt))
