"""Where does the time of metrics.evaluate_auto go at RGBNT100 size?  (stage timings, CUDA events + host clock)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demo2_b200 import metrics, synth, _lib

dev = torch.device("cuda")
for key in ("rgbnt201", "rgbnt100"):
    s = synth.make_named(key, sigma=4.0, seed=0)
    qf, gf = s.qf.to(dev), s.gf.to(dev)

    def t(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - h0) / n * 1e3, r

    print(key)
    print("  RankPlan            %.3f ms gpu / %.3f ms host" % t(lambda: metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids))[:2])
    plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    print("  plan: T=%d max_cnt=%d queried=%d" % (plan.T, plan.max_cnt, plan.n_queried))
    print("  sqdist_device       %.3f ms gpu / %.3f ms host" % t(lambda: metrics.sqdist_device(qf, gf, _lib.DIST_SQ, normalize=True))[:2])
    dm = metrics.sqdist_device(qf, gf, _lib.DIST_SQ, normalize=True)
    print("  evaluate_matrix     %.3f ms gpu / %.3f ms host" % t(lambda: metrics.evaluate_matrix(dm, plan=plan))[:2])
    print("  evaluate_features   %.3f ms gpu / %.3f ms host" % t(lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True))[:2])
    print("  evaluate_auto       %.3f ms gpu / %.3f ms host" % t(lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))[:2])
    print("  evaluate_auto(plan) %.3f ms gpu / %.3f ms host" % t(lambda: metrics.evaluate_auto(qf, gf, plan=plan, normalize=True))[:2])
