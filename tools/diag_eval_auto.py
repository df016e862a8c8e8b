import sys, time, torch, numpy as np, gc
sys.path.insert(0, '.')
from demo2_b200 import metrics, reranking, synth
dev = torch.device('cuda')
def seg():
    st = torch.cuda.memory_stats()
    return st['segment.all.allocated'], st['segment.all.freed'], st['reserved_bytes.all.current'] >> 20
def timed(name, fn, iters=10, busy_s=0.2):
    t0 = time.perf_counter(); n = 0
    while n < 3 or time.perf_counter() - t0 < busy_s:
        out = fn(); n += 1
        if n % 4 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize()
    s0 = seg()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    walls = []
    evs[0].record()
    for i in range(iters):
        t = time.perf_counter(); out = fn(); walls.append((time.perf_counter() - t) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    s1 = seg()
    g = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    print('%-30s mean %.3f max %.2f | cpu max %.2f | segments alloc %d->%d freed %d->%d reserved %d MB' % (name, np.mean(g), max(g), max(walls), s0[0], s1[0], s0[1], s1[1], s1[2]))
    return out
for rep in range(2):
  for key in ("rgbnt201", "rgbnt100"):
    s = synth.make_named(key, sigma=4.0, seed=0)
    qf, gf = s.qf.to(dev), s.gf.to(dev)
    plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
    Q, G = qf.shape[0], gf.shape[0]
    r = timed(key + '_eval', lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))
    r = timed(key + '_eval_fused', lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True))
    def rr():
        dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
        return metrics.evaluate_matrix(dist, plan=plan)
    r = timed(key + '_rerank', rr, iters=5)
    allp = metrics.sqdist_device(torch.cat([qf, gf]), torch.cat([qf, gf]), normalize=True)
    timed(key + ' topk', lambda: reranking.topk_rows(allp, 21), iters=20)
    dist_m = allp[:Q, Q:].contiguous()
    timed(key + ' eval_matrix', lambda: metrics.evaluate_matrix(dist_m, plan=plan), iters=20)
    del allp, dist_m
