import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import reranking
torch.manual_seed(0)
Q, G, d = 4096, 28672, 512
nid = 1500
centers = torch.randn(nid, d, device="cuda")
qp = torch.randint(0, nid, (Q,), device="cuda"); gp = torch.randint(0, nid, (G,), device="cuda")
qf = centers[qp] + 3 * torch.randn(Q, d, device="cuda"); gf = centers[gp] + 3 * torch.randn(G, d, device="cuda")
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True); e1.record()
    torch.cuda.synchronize()
    print("iter %d: re_ranking N=%d: %.3f ms" % (it, Q + G, e0.elapsed_time(e1)))
