"""Timeline of one streamed host evaluation (ShardedEvaluator.evaluate_host, one GPU, headline
size): every engine stage (prepare = PCIe pull of a gallery piece, extract, thresholds, count) is
bracketed by CUDA events on the stream it is issued to, and printed as start / end offsets from the
start of the call -- where the tensor cores wait for data and where the link waits for nothing.

    python tools/e2e_timeline.py [--medium] [--groups K]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demo2_b200 import parallel  # noqa: E402

dev = torch.device("cuda")
Q, G, d, nid = (8192, 262144, 1536, 13000) if "--medium" in sys.argv else (20000, 1000000, 1536, 50000)
groups = int(sys.argv[sys.argv.index("--groups") + 1]) if "--groups" in sys.argv else None

rng = np.random.default_rng(0)
qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
gen = torch.Generator(device=dev).manual_seed(0)
centers = torch.randn(nid, d, device=dev, generator=gen)
q_host = (centers[torch.from_numpy(qp).to(dev)] + 4.0 * torch.randn(Q, d, device=dev, generator=gen)).cpu().pin_memory()
g_host = torch.empty(G, d).pin_memory()
gpd = torch.from_numpy(gp).to(dev)
for s in range(0, G, 131072):
    e = min(G, s + 131072)
    g_host[s:e].copy_(centers[gpd[s:e]] + 4.0 * torch.randn(e - s, d, device=dev, generator=gen))
del centers
torch.cuda.synchronize()

ev = parallel.ShardedEvaluator()
eng = ev.engine
log = []


def wrap(name, describe):
    inner = getattr(eng, name)

    def f(*a, **k):
        s = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        out = inner(*a, **k)
        e1.record(s)
        log.append((name, describe(*a, **k), e0, e1))
        return out
    setattr(eng, name, f)


wrap("prepare", lambda plan, w, x, which, row0, nrows, *a, **k: "%s rows %d..%d" % ("gallery" if which else "query", row0, row0 + nrows))
wrap("extract", lambda plan, w, base, g_index=None, q_row0=0, q_nrows=None: "queries %d+%s" % (q_row0, q_nrows))
wrap("thresholds_into", lambda ro, recs, thr, q0, qn: "queries %d+%d" % (q0, qn))
wrap("count", lambda w, plan, *a, g_row0=0, g_nrows=None, q_row0=0, q_nrows=None, **k:
     "gallery %d+%s x queries %d+%s" % (g_row0, g_nrows, q_row0, q_nrows))

kw = dict(normalize=True)
if groups is not None:
    kw["query_groups"] = groups
for it in range(3):
    log.clear()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    r = ev.evaluate_host(q_host, g_host, qp, gp, qc, gc, **kw)
    t1.record()
    torch.cuda.synchronize()
print("evaluate_host %.2f ms, mAP %.6f" % (t0.elapsed_time(t1), r.mAP))
busy = {"prepare": 0.0, "count": 0.0}
for name, what, e0, e1 in sorted(log, key=lambda x: t0.elapsed_time(x[2])):
    a, b = t0.elapsed_time(e0), t0.elapsed_time(e1)
    if name in busy:
        busy[name] += b - a
    print("%8.2f .. %8.2f  (%7.2f ms)  %-16s %s" % (a, b, b - a, name, what))
print("sum of prepare intervals %.2f ms, sum of count intervals %.2f ms" % (busy["prepare"], busy["count"]))
