import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import reranking, synth
for shape in ("rgbnt201", "rgbnt100"):
    s = synth.make_named(shape, sigma=5.0, seed=0)
    qf, gf = s.qf.cuda(), s.gf.cuda()
    for k1, k2 in ((20, 6), (50, 15)):
        for it in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); d = reranking.re_ranking_device(qf, gf, k1, k2, 0.3, normalize=True); e1.record()
            torch.cuda.synchronize()
        print("%s re_ranking(k1=%d, k2=%d): %.3f ms" % (shape, k1, k2, e0.elapsed_time(e1)))
