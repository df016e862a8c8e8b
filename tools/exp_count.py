"""Timing experiment: count-GEMM stage of the full-size evaluation (20k x 1M, d=1536), a few
iterations, prints the per-stage CUDA-event times and the SM clocks sampled meanwhile.  Tuning /
diagnostic switches of the library (read once per process): DEMO_DEBUG_NOEPI=1 mainloop only
(results are garbage), DEMO_COUNT_1CTA=1 one CTA per tile instead of CTA pairs, DEMO_GROUP_M /
DEMO_CHUNK_TILES / DEMO_PAIRS unit grouping, DEMO_PACE=<window>|-1 / DEMO_PACE_TILES pacing of the CTA
pairs, DEMO_DEBUG_TIES=1 prints the tie-list fill.
    python tools/exp_count.py [Q G d iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import parallel  # noqa: E402

Q, G, d, iters = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (20000, 1000000, 1536, 4)
nid = max(2, G // 20)
rng = np.random.default_rng(0)
qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
qc, gc = rng.integers(0, 8, Q), rng.integers(0, 8, G)
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
centers = torch.randn(nid, d, device=dev, generator=gen)
qf = centers[torch.from_numpy(qp).to(dev)] + 4 * torch.randn(Q, d, device=dev, generator=gen)
gf = torch.empty(G, d, device=dev)
gpd = torch.from_numpy(gp).to(dev)
for s in range(0, G, 131072):
    e = min(G, s + 131072)
    gf[s:e] = centers[gpd[s:e]] + 4 * torch.randn(e - s, d, device=dev, generator=gen)
del centers
lab = [torch.from_numpy(x).int().to(dev) for x in (qp, gp, qc, gc)]
ev = parallel.ShardedEvaluator()
out = []
import bench  # noqa: E402
sampler = bench.ClockSampler(0)
sampler.start()
for it in range(iters + 2):
    t = {}
    res = ev.evaluate(qf, gf, lab[0], lab[1], lab[2], lab[3], normalize=True, timers=t)
    torch.cuda.synchronize()
    if it >= 2:
        out.append({k: v[0].elapsed_time(v[1]) for k, v in t.items() if isinstance(v, tuple)})
clk = sampler.stop()
keys = out[0].keys()
print("NOEPI=%s 1CTA=%s  Q=%d G=%d d=%d  mAP %.5f" % (os.environ.get("DEMO_DEBUG_NOEPI"),
                                                     os.environ.get("DEMO_COUNT_1CTA"), Q, G, d, res.mAP))
print("  " + "  ".join("%s %.2f" % (k, float(np.mean([o[k] for o in out]))) for k in keys))
cm = float(np.mean([o["count"] for o in out]))
print("  clocks", clk)
print("  count: %.2f ms -> %.1f TFLOP/s algorithmic, %.1f executed" % (cm, 2.0 * Q * G * d / cm * 1e-9,
                                                                     6.0 * Q * G * d / cm * 1e-9))
