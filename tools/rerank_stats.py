import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import parallel, synth
def stats(qf, gf, k1, k2, tag):
    eng = parallel.CudaRerankEngine()
    Q, G = qf.shape[0], gf.shape[0]; N = Q + G
    feat = torch.cat([qf, gf]).float().cuda()
    dev = eng.begin(feat, Q, k1, k2, 0, N, N, True)
    K, cap, capq = eng.dims(N, k1, k2)
    rank_all = torch.zeros((N, K), dtype=torch.int32, device=dev)
    v_idx = torch.zeros((N, cap), dtype=torch.int32, device=dev); v_val = torch.zeros((N, cap), dtype=torch.float16, device=dev); v_cnt = torch.zeros(N, dtype=torch.int32, device=dev)
    eng.topk(rank_all); eng.krecip(rank_all, v_idx, v_val, v_cnt)
    q_idx = torch.zeros((N, capq), dtype=torch.int32, device=dev); q_val = torch.zeros((N, capq), dtype=torch.float16, device=dev); q_cnt = torch.zeros(N, dtype=torch.int32, device=dev)
    eng.expand(rank_all, v_idx, v_val, v_cnt, q_idx, q_val, q_cnt)
    torch.cuda.synchronize()
    vc, qc = v_cnt.cpu().numpy(), q_cnt.cpu().numpy()
    cols = torch.cat([q_idx[i, :qc[i]] for i in range(0, N, max(1, N // 2000))]).cpu().numpy()
    col_len = np.bincount(q_idx.cpu().numpy()[np.arange(capq)[None, :] < qc[:, None]], minlength=N)
    work = np.array([col_len[q_idx[i, :qc[i]].cpu().numpy()].sum() for i in range(0, Q, max(1, Q // 300))])
    print("%s: N=%d cap=%d capq=%d | V nnz mean %.1f max %d | V_qe nnz mean %.1f max %d | inverted list mean %.1f max %d | entries per query mean %.0f max %d"
          % (tag, N, cap, capq, vc.mean(), vc.max(), qc.mean(), qc.max(), col_len.mean(), col_len.max(), work.mean(), work.max()))
s = synth.make_named("rgbnt100", sigma=5.0, seed=0)
stats(s.qf, s.gf, 20, 6, "rgbnt100 k20/6")
stats(s.qf, s.gf, 50, 15, "rgbnt100 k50/15")
torch.manual_seed(0)
Q, G, d, nid = 4096, 28672, 512, 1500
centers = torch.randn(nid, d); qp = torch.randint(0, nid, (Q,)); gp = torch.randint(0, nid, (G,))
stats(centers[qp] + 3 * torch.randn(Q, d), centers[gp] + 3 * torch.randn(G, d), 20, 6, "large k20/6")
