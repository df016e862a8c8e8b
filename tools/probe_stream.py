"""Probe of the streamed host evaluation (one GPU):

1. zero-copy prepare: GB/s of demo_eval_prepare reading PINNED HOST rows in place over PCIe,
   against cudaMemcpy H2D of the same bytes and against the device-resident prepare;
2. the same prepare while a count GEMM is resident on every SM (does it make progress?);
3. evaluate_host vs copy-then-evaluate at a medium and at the headline size, with stage times.

    python tools/probe_stream.py [--large]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demo2_b200 import metrics, parallel  # noqa: E402

dev = torch.device("cuda")
large = "--large" in sys.argv


def ev_ms(fn, iters=3, warm=1):
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


def make(Q, G, d, nid, seed=0):
    rng = np.random.default_rng(seed)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
    gen = torch.Generator(device=dev).manual_seed(seed)
    centers = torch.randn(nid, d, device=dev, generator=gen)
    qf = centers[torch.from_numpy(qp).to(dev)] + 4.0 * torch.randn(Q, d, device=dev, generator=gen)
    gf = torch.empty(G, d, device=dev)
    gpd = torch.from_numpy(gp).to(dev)
    for s in range(0, G, 131072):
        e = min(G, s + 131072)
        gf[s:e] = centers[gpd[s:e]] + 4.0 * torch.randn(e - s, d, device=dev, generator=gen)
    return qf, gf, qp, gp, qc, gc


Q, G, d, nid = (20000, 1000000, 1536, 50000) if large else (8192, 262144, 1536, 13000)
qf, gf, qp, gp, qc, gc = make(Q, G, d, nid)
q_host, g_host = qf.cpu().pin_memory(), gf.cpu().pin_memory()
nbytes = G * d * 4
print("problem %d x %d x %d, gallery %.2f GB" % (Q, G, d, nbytes * 1e-9))

# 1. raw transfer numbers
g_dev = torch.empty_like(gf)
ms = ev_ms(lambda: g_dev.copy_(g_host, non_blocking=True))
print("cudaMemcpy H2D (pinned)            %8.2f ms  %6.1f GB/s" % (ms, nbytes / ms * 1e-6))
eng = parallel.CudaEngine()
plan = eng.plan(qp, gp, qc, gc)
plan.finish(plan.info.cpu())
w = eng.workspace(plan, d, plan.max_cnt)
print("plan: T=%d max_cnt=%d queried rows=%d (%.1f %%)" % (plan.T, plan.max_cnt, plan.n_queried, 100.0 * plan.n_queried / G))
ms = ev_ms(lambda: eng.prepare(plan, w, gf, 1, 0, G, True))
print("prepare, device-resident rows      %8.2f ms  %6.1f GB/s (read)" % (ms, nbytes / ms * 1e-6))
ms = ev_ms(lambda: eng.prepare(plan, w, g_host, 1, 0, G, True, host_input=True))
print("prepare, zero-copy from pinned host %7.2f ms  %6.1f GB/s (PCIe read, sorted-order gather)" % (ms, nbytes / ms * 1e-6))

# 2. zero-copy prepare of a slab while the count GEMM owns the SMs
eng.prepare(plan, w, qf, 0, 0, Q, True)
eng.prepare(plan, w, gf, 1, 0, G, True)
recs = eng.extract(plan, w, 0)
thr = eng.thresholds(plan.rec_ofs, recs, Q)
counts = torch.zeros(max(plan.T, 1), dtype=torch.int32, device=dev)
half = (G // 2 // 256) * 256
ms_count = ev_ms(lambda: eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, plan.max_cnt, 0, half))
print("count GEMM over %d rows alone      %8.2f ms" % (half, ms_count))
side = torch.cuda.Stream()
w2 = eng.workspace(plan, d, plan.max_cnt)   # scratch target for the concurrent prepare


def both():
    main = torch.cuda.current_stream()
    e = torch.cuda.Event()
    e.record(main)
    side.wait_event(e)
    with torch.cuda.stream(side):
        eng.prepare(plan, w2, g_host, 1, half, G - half, True, host_input=True)
        done = torch.cuda.Event()
        done.record(side)
    eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, plan.max_cnt, 0, half)
    main.wait_event(done)


def both_gemm_first():
    main = torch.cuda.current_stream()
    eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, plan.max_cnt, 0, half)
    e = torch.cuda.Event()
    e.record(main)   # NOTE: recorded after the GEMM launch -> the prepare starts after it; use the plain variant for overlap
    with torch.cuda.stream(side):
        eng.prepare(plan, w2, g_host, 1, half, G - half, True, host_input=True)
        done = torch.cuda.Event()
        done.record(side)
    main.wait_event(done)


ms_both = ev_ms(both)
ms_prep_half = ev_ms(lambda: eng.prepare(plan, w2, g_host, 1, half, G - half, True, host_input=True))
print("zero-copy prepare of the other half alone %6.2f ms; both concurrently %8.2f ms (sum %.2f, max %.2f)"
      % (ms_prep_half, ms_both, ms_count + ms_prep_half, max(ms_count, ms_prep_half)))
del w2
if "--overlap-only" in sys.argv:
    sys.exit(0)

# 3. end to end
ev = parallel.ShardedEvaluator()
base = ev.evaluate(qf, gf, qp, gp, qc, gc, normalize=True)
ms_dev = ev_ms(lambda: ev.evaluate(qf, gf, qp, gp, qc, gc, normalize=True))


def staged():
    g_dev.copy_(g_host, non_blocking=True)
    return ev.evaluate(q_host.to(dev, non_blocking=True), g_dev, qp, gp, qc, gc, normalize=True)


ms_staged = ev_ms(staged)
for slab in (65536, 131072, 262144):
    t = {}
    r = ev.evaluate_host(q_host, g_host, qp, gp, qc, gc, normalize=True, slab_rows=slab, timers=t)
    ok = bool(r.mAP == base.mAP and np.array_equal(r.cmc, base.cmc))
    ms_host = ev_ms(lambda: ev.evaluate_host(q_host, g_host, qp, gp, qc, gc, normalize=True, slab_rows=slab))
    stages = {k: round(v[0].elapsed_time(v[1]), 2) for k, v in t.items() if isinstance(v, tuple)}
    print("evaluate_host slab_rows=%-7d %8.2f ms (identical=%s) stages %s" % (slab, ms_host, ok, stages))
print("device-resident evaluate           %8.2f ms" % ms_dev)
print("copy everything, then evaluate     %8.2f ms" % ms_staged)
