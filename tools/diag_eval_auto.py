import sys, time, torch, numpy as np, gc
sys.path.insert(0, '.')
from demo2_b200 import metrics, reranking, synth
dev = torch.device('cuda')
def timed_detail(name, fn, iters=10, busy_s=0.2):
    t0 = time.perf_counter(); n = 0
    while n < 3 or time.perf_counter() - t0 < busy_s:
        fn(); n += 1
        if n % 4 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    walls = []
    evs[0].record()
    for i in range(iters):
        t = time.perf_counter(); fn(); walls.append((time.perf_counter() - t) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    g = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    print('%-28s warm calls %4d | gpu ms/iter %s | cpu ms/iter %s' % (name, n, [round(x, 2) for x in g], [round(x, 2) for x in walls]))
for key in ("rgbnt201", "rgbnt100"):
    s = synth.make_named(key, sigma=4.0, seed=0)
    qf, gf = s.qf.to(dev), s.gf.to(dev)
    plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    timed_detail(key + '_eval', lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))
    timed_detail(key + '_eval_fused', lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True))
    def rr():
        dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
        return metrics.evaluate_matrix(dist, plan=plan)
    timed_detail(key + '_rerank', rr, iters=5)
    timed_detail(key + '_eval again', lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))
print('gc counts', gc.get_count(), 'mem', torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20)
