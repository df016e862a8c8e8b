"""Multi-GPU (NCCL) correctness check, launched with torchrun on one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multi_gpu.py

Gallery-sharded evaluation and row-sharded re-ranking must reproduce the single-GPU results bit
for bit on every rank (rank counts are additive over gallery shards; the re-ranking stages are
row-local given the gathered neighbour lists and sparse V rows)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demo2_b200 import metrics, parallel, reranking, synth  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ok = True
for shape, sigma in (("rgbnt201", 4.0), ("rgbnt100", 4.0)):
    s = synth.make_named(shape, sigma=sigma, seed=0)
    qf, gf = s.qf.to(dev), s.gf.to(dev)
    one = metrics.evaluate_features(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True)
    lo, hi = parallel.shard_range(len(s.g_pids), world, rank)
    ev = parallel.ShardedEvaluator(world=world, rank=rank, group=dist.group.WORLD)
    res = ev.evaluate(qf, gf[lo:hi], s.q_pids, s.g_pids[lo:hi], s.q_camids, s.g_camids[lo:hi], g_index_base=lo,
                      normalize=True)
    same = (torch.equal(res.first.cpu(), one.first.cpu()) and torch.equal(res.ap.cpu(), one.ap.cpu())
            and res.mAP == one.mAP and np.array_equal(res.cmc, one.cmc))
    print("[rank %d] sharded eval %s: mAP %.6f R1 %.4f  identical to 1 GPU: %s" % (rank, shape, res.mAP, res.cmc[0], same))
    ok &= same
    k1, k2 = (20, 6)
    whole = reranking.re_ranking_device(qf, gf, k1, k2, 0.3, normalize=True)
    shard = parallel.ShardedReranker(world=world, rank=rank, group=dist.group.WORLD).re_ranking(qf, gf, k1, k2, 0.3,
                                                                                               normalize=True)
    same = torch.equal(whole, shard)
    print("[rank %d] sharded re-ranking %s (k1=%d, k2=%d): identical to 1 GPU: %s" % (rank, shape, k1, k2, same))
    ok &= same
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
dist.destroy_process_group()
if rank == 0:
    print("MULTI-GPU CHECK", "PASS" if int(t.item()) == 0 else "FAIL")
sys.exit(0 if int(t.item()) == 0 else 1)
