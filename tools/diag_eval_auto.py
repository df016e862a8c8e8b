import sys, time, torch
sys.path.insert(0, '.')
from demo2_b200 import metrics, synth
s = synth.make_named("rgbnt100", sigma=4.0, seed=0)
qf, gf = s.qf.cuda(), s.gf.cuda()
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t)/n*1e3
print("auto", ev(lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True)))
plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
print("plan", ev(lambda: metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)))
print("sqdist", ev(lambda: metrics.sqdist_device(qf, gf, normalize=True)))
d = metrics.sqdist_device(qf, gf, normalize=True)
print("matrix", ev(lambda: metrics.evaluate_matrix(d, plan=plan)))
print("fused", ev(lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True)))
