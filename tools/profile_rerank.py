"""Re-ranking + materialised evaluation at RGBNT100 scale for ncu launch lists:
python tools/profile_rerank.py [shape]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics, reranking, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "rgbnt100"
s = synth.make_named(shape, sigma=5.0, seed=0)
qf, gf = s.qf.cuda(), s.gf.cuda()
plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
for it in range(3):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
    e[1].record()
    res = metrics.evaluate_matrix(dist, plan=plan)
    e[2].record()
    torch.cuda.synchronize()
    print("iter %d: re_ranking %.3f ms, eval_func(matrix) %.3f ms, mAP %.5f" %
          (it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), res.mAP))
