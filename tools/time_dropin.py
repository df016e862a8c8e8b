import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics, reranking, synth
s = synth.make_named("rgbnt100", sigma=5.0, seed=0)
feats = torch.cat([s.qf, s.gf]).cuda()
pids = np.concatenate([s.q_pids, s.g_pids]); cams = np.concatenate([s.q_camids, s.g_camids])
for rr in (False, True):
    for it in range(4):
        ev = metrics.R1_mAP_eval(s.num_query, feat_norm=True, reranking=rr)
        ev.update((feats, pids, torch.from_numpy(cams), ["x"] * len(pids)))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cmc, mAP, distmat, *_ = ev.compute()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("R1_mAP_eval.compute(reranking=%s) rgbnt100 (host wall, incl. D2H of the %d x %d distmat): %.2f ms  mAP %.4f" % (rr, *distmat.shape, dt * 1e3, mAP))
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = metrics.euclidean_distance(s.qf, s.gf)
    dt = time.perf_counter() - t0
print("euclidean_distance(host in, numpy out): %.2f ms" % (dt * 1e3))
