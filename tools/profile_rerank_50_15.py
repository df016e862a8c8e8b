import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import reranking, synth
s = synth.make_named("rgbnt100", sigma=5.0, seed=0)
qf, gf = s.qf.cuda(), s.gf.cuda()
for it in range(3):
    d = reranking.re_ranking_device(qf, gf, 50, 15, 0.3, normalize=True)
torch.cuda.synchronize()
