"""HBM-bound kernels of the path at sizes beyond L2 (126 MB): achieved GB/s on ALGORITHMIC bytes
against the measured HBM copy peak (MEASURED_PEAKS.json).  CUDA events, 3 warm-ups, 10 timed runs.

    python tools/bench_hbm_kernels.py > profiles/hbm_kernels_<round>.txt
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demo2_b200 import metrics, reranking  # noqa: E402

peak = 6546.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timed(fn, iters=10, warm=3):
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


def report(name, ms, nbytes, note=""):
    gbs = nbytes / ms * 1e-6
    print("%-58s %9.3f ms  %8.1f GB/s  %5.1f %% of %.0f GB/s  %s" % (name, ms, gbs, 100 * gbs / peak, peak, note))


dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
print("# algorithmic bytes / CUDA-event time; HBM peak = measured copy bandwidth (MEASURED_PEAKS.json)")

# operand preparation: 4*d read + 4*d written per row
rows, d = 1000000, 1536
x = torch.randn(rows, d, device=dev, generator=gen)
other = torch.randn(256, d, device=dev, generator=gen)
out = torch.empty(256, 8, device=dev)
# prep runs inside sqdist; isolate it by a tiny second operand (GEMM 256 x rows is negligible next to 12 GB of traffic?)
# -> time the evaluation records stage instead: prep of both operands dominates
ms = timed(lambda: metrics.sqdist_device(other[:8], x[:262144], normalize=True), iters=5)
report("prep_rows (262144 x 1536, via sqdist 8 x 262144)", ms, 262144 * d * 8, "(includes an 8-row GEMM + 8 MB store)")
del x

# top-k of a materialised matrix: 4 B per entry read once
for R, C, k in ((10290, 10290, 21), (4096, 262144, 50), (16384, 65536, 21)):
    m = torch.rand(R, C, device=dev, generator=gen)
    ms = timed(lambda: reranking.topk_rows(m, k))
    report("topk_rows %d x %d, k=%d" % (R, C, k), ms, R * C * 4)
    if C == 10290:   # the all-pairs matrix of the re-ranking path has a 128-byte row pitch: every row 16-byte aligned
        for kk in (21, 51):
            mp = torch.rand(R, 10304, device=dev, generator=gen)[:, :C]
            ms = timed(lambda: reranking.topk_rows(mp, kk))
            report("topk_rows %d x %d (row pitch 10304), k=%d" % (R, C, kk), ms, R * C * 4)
        del mp
    del m

# rank counts over a materialised matrix (eval_func on a distance matrix): 4 B per pair read once
for Q, G, per_id in ((1715, 8575, 20), (4096, 262144, 20), (8192, 131072, 20), (4096, 262144, 170), (1715, 8575, 170)):
    rng = np.random.default_rng(0)
    nid = max(2, G // per_id)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
    dist = torch.rand(Q, G, device=dev, generator=gen)
    plan = metrics.RankPlan(qp, gp, qc, gc)
    ms = timed(lambda: metrics.evaluate_matrix(dist, plan=plan))
    report("evaluate_matrix (records+thresholds+count+finalize) %d x %d, ~%d per id" % (Q, G, per_id), ms, Q * G * 4,
           "(whole call incl. D2H of the metrics; uniform random distances = thresholds anywhere in the row)")
    if G >= 131072:
        # retrieval-like: the positives of a query are among its nearest gallery items (mAP ~ 0.9)
        qpd, gpd = torch.from_numpy(qp).to(dev), torch.from_numpy(gp).to(dev)
        for s0 in range(0, Q, 512):
            same = qpd[s0:s0 + 512, None] == gpd[None, :]
            dist[s0:s0 + 512] = torch.where(same, dist[s0:s0 + 512] * (2.0 * per_id / G), dist[s0:s0 + 512])
        del same
        ms = timed(lambda: metrics.evaluate_matrix(dist, plan=plan))
        r = metrics.evaluate_matrix(dist, plan=plan)
        report("  the same, retrieval-like distances (mAP %.2f)" % r.mAP, ms, Q * G * 4, "(positives among the nearest items)")
    del dist
