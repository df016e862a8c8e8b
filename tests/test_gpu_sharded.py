"""GPU tests of the multi-GPU data flows with the CUDA engine (through the C ABI), on ONE device:

* two real processes (torch.distributed, gloo group, payloads staged through the host) that share
  cuda:0 -- gallery-sharded evaluation, evaluation under DDP (DistributedR1mAP) and row-sharded
  re-ranking (ShardedReranker) must equal the single-process result bit for bit;
* the streamed host evaluation (pinned host features pulled over PCIe slab by slab, queried rows
  first) == the device-resident evaluation;
* query blocks with more than one window of thresholds (slab path) == the oracle.

The NCCL runs of the same flows are part of bench.py (`multi_gpu_checks` in the N > 1 line)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import make_case, oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_per_query(dist_m, qp, gp, qc, gc):
    ofs, idx, r, c = oracle.rank_counts(dist_m, qp, gp, qc, gc)
    Q = len(ofs) - 1
    ap = np.full(Q, -1.0)
    first = np.zeros(Q, np.int64)
    for q in range(Q):
        s, e = ofs[q], ofs[q + 1]
        if e > s:
            ap[q] = (c[s:e] / r[s:e]).sum() / (e - s)
            first[q] = r[s:e].min()
    return ap, first


def _case():
    qf, gf, qp, gp, qc, gc = make_case("rgbnt201", 0, 4.0)
    qf, gf = qf[:300, :512].copy(), gf[:701, :512].copy()
    qp, gp, qc, gc = qp[:300].copy(), gp[:701].copy(), qc[:300], gc[:701]
    gf[40:50] = gf[600:610]      # exact ties across shards
    qp[3] = 999                  # identity absent -> skipped
    return qf, gf, qp, gp, qc, gc


def _case_big():
    """>= 2048 queries: the streamed evaluation runs its grouped flow (one record exchange per
    query block); a fifth of the gallery has ids no query asks for, some ids have > 63 images."""
    rng = np.random.default_rng(5)
    Q, G, d, nid = 2300, 6001, 128, 150
    centers = rng.standard_normal((nid, d)).astype(np.float32)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    gp[gp < 8] = 3                               # one identity with ~300 gallery images (slab path)
    gp[::5] += 4000
    qp[:5] = 88888
    qc, gc = rng.integers(0, 4, Q), rng.integers(0, 4, G)
    qf = centers[qp % nid] + 2.0 * rng.standard_normal((Q, d)).astype(np.float32)
    gf = centers[gp % nid] + 2.0 * rng.standard_normal((G, d)).astype(np.float32)
    gf[100:140] = gf[5000:5040]                  # exact ties across shards
    return qf, gf, qp, gp, qc, gc


# ---------------------------------------------------------------------------------------------
# two processes, one GPU
# ---------------------------------------------------------------------------------------------
def _worker(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from demo2_b200 import parallel
        torch.cuda.set_device(0)
        qf, gf, qp, gp, qc, gc = _case_big() if mode == "sharded_host_groups" else _case()
        if mode == "sharded_host_groups":
            parallel.MIN_PIECE_ROWS = 64          # the library only groups when the pieces are large; force it here
            lo, hi = parallel.shard_range(len(gp), world, rank)
            ev = parallel.ShardedEvaluator(world=world, rank=rank, group=dist.group.WORLD)
            timers = {}
            res = ev.evaluate_host(torch.from_numpy(qf).pin_memory(), torch.from_numpy(gf[lo:hi].copy()).pin_memory(),
                                   qp, gp[lo:hi], qc, gc[lo:hi], g_index_base=lo, normalize=True, slab_rows=512,
                                   query_groups=2, timers=timers)
            assert timers.get("query_groups") == 2
            out[rank] = (res.cmc, float(res.mAP), res.num_valid, res.ap.cpu().numpy(), res.first.cpu().numpy())
        elif mode in ("sharded", "sharded_host"):
            lo, hi = parallel.shard_range(len(gp), world, rank)
            ev = parallel.ShardedEvaluator(world=world, rank=rank, group=dist.group.WORLD)
            assert isinstance(ev.engine, parallel.CudaEngine)
            if mode == "sharded":
                res = ev.evaluate(torch.from_numpy(qf).cuda(), torch.from_numpy(gf[lo:hi]).cuda(), qp, gp[lo:hi], qc,
                                  gc[lo:hi], g_index_base=lo, normalize=True)
            else:
                res = ev.evaluate_host(torch.from_numpy(qf).pin_memory(), torch.from_numpy(gf[lo:hi].copy()).pin_memory(),
                                       qp, gp[lo:hi], qc, gc[lo:hi], g_index_base=lo, normalize=True, slab_rows=128)
            out[rank] = (res.cmc, float(res.mAP), res.num_valid, res.ap.cpu().numpy(), res.first.cpu().numpy())
        elif mode == "ddp":
            from torch.utils.data import DistributedSampler
            feats = np.concatenate([qf, gf])
            pids, cams = np.concatenate([qp, gp]), np.concatenate([qc, gc])
            ev = parallel.DistributedR1mAP(len(qp), world=world, rank=rank, group=dist.group.WORLD, feat_norm=True)
            mine = np.asarray(list(DistributedSampler(range(len(pids)), num_replicas=world, rank=rank, shuffle=False)))
            for s in range(0, len(mine), 64):
                b = mine[s:s + 64]
                ev.update((torch.from_numpy(feats[b]).cuda(), pids[b], torch.from_numpy(cams[b]), torch.from_numpy(b)))
            cmc, mAP = ev.compute()
            r = ev.last_result
            out[rank] = (cmc, float(mAP), r.num_valid, r.ap.cpu().numpy(), r.first.cpu().numpy())
        elif mode == "rerank":
            rr = parallel.ShardedReranker(world=world, rank=rank, group=dist.group.WORLD)
            res = rr.re_ranking(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda(), 20, 6, 0.3)
            out[rank] = res.cpu().numpy()
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def _spawn(mode, world=2):
    mgr = mp.get_context("spawn").Manager()   # never fork a process that holds a CUDA context
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert len(out) == world
    return out


@pytest.mark.parametrize("mode", ["sharded", "sharded_host", "sharded_host_groups", "ddp"])
def test_two_process_evaluation_equals_single_gpu(mode):
    """Gallery sharded over two processes (CUDA engine, real process group) == one GPU, bit for bit:
    distances are position-independent and rank counts are additive over gallery shards."""
    from demo2_b200 import metrics
    qf, gf, qp, gp, qc, gc = _case_big() if mode == "sharded_host_groups" else _case()
    assert (len(qp) + len(gp)) % 2 == 1          # the DDP sampler really pads
    single = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    out = _spawn(mode)
    for r in range(2):
        cmc, mAP, nv, ap, first = out[r]
        np.testing.assert_array_equal(cmc, single.cmc)
        assert mAP == float(single.mAP) and nv == single.num_valid
        np.testing.assert_array_equal(ap, single.ap.cpu().numpy())
        np.testing.assert_array_equal(first, single.first.cpu().numpy())
    # and the single-GPU result is the oracle's on our own matrix
    ours = metrics.sqdist_device(qf, gf, normalize=True).cpu().numpy()
    ap_o, first_o = _oracle_per_query(ours, qp, gp, qc, gc)
    np.testing.assert_array_equal(single.first.cpu().numpy(), first_o)
    np.testing.assert_allclose(single.ap.cpu().numpy(), ap_o, atol=1e-12)


def test_two_process_reranking_equals_single_gpu():
    from demo2_b200 import reranking
    qf, gf, *_ = _case()
    whole = reranking.re_ranking_device(qf, gf, 20, 6, 0.3).cpu().numpy()
    out = _spawn("rerank")
    for r in range(2):
        np.testing.assert_array_equal(out[r], whole)


# ---------------------------------------------------------------------------------------------
# streamed host evaluation, one process
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slab_rows", [256, 1000, 1 << 20])
def test_streamed_host_evaluation_equals_device_evaluation(slab_rows):
    from demo2_b200 import metrics, parallel
    qf, gf, qp, gp, qc, gc = make_case("rgbnt100", 0, 4.0)
    gp = gp.copy()
    gp[::3] += 1000                              # a third of the gallery has ids no query asks for
    single = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    ev = parallel.ShardedEvaluator()
    timers = {}
    res = ev.evaluate_host(torch.from_numpy(qf).pin_memory(), torch.from_numpy(gf).pin_memory(), qp, gp, qc, gc,
                           normalize=True, slab_rows=slab_rows, timers=timers)
    assert timers["queried_rows"] < len(gp) and timers["slabs"] >= 1
    np.testing.assert_array_equal(res.cmc, single.cmc)
    assert float(res.mAP) == float(single.mAP)
    np.testing.assert_array_equal(res.ap.cpu().numpy(), single.ap.cpu().numpy())
    # pageable host input and device input go through the same call
    res2 = ev.evaluate_host(torch.from_numpy(qf), torch.from_numpy(gf).cuda(), qp, gp, qc, gc, normalize=True,
                            slab_rows=slab_rows)
    assert float(res2.mAP) == float(single.mAP)


@pytest.mark.parametrize("Q,G,nid,groups,force", [(2600, 12000, 400, 4, True), (4200, 9000, 700, 3, True),
                                                   (2100, 8000, 30, 2, True), (2600, 12000, 400, 4, False)])
def test_streamed_host_evaluation_query_groups(Q, G, nid, groups, force, monkeypatch):
    """One GPU, >= 2048 queries: the queried gallery rows are pulled in per block of pid-sorted
    queries and the count GEMM starts on the first block's rectangle while the others are still in
    flight (ShardedEvaluator._evaluate_host_grouped).  The rectangles tile Q x G exactly once:
    same integers as the device-resident evaluation (nid = 30: ~270 gallery images per id, slab path)."""
    from demo2_b200 import metrics, parallel
    if force:
        monkeypatch.setattr(parallel, "MIN_PIECE_ROWS", 64)   # small test galleries: force the grouped flow
    rng = np.random.default_rng(Q + G)
    d = 128
    centers = rng.standard_normal((nid, d)).astype(np.float32)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    gp[::4] += 5000                              # a quarter of the gallery has ids no query asks for
    qp[:7] = 77777                               # and a few queries ask for an id the gallery lacks
    qc, gc = rng.integers(0, 4, Q), rng.integers(0, 4, G)
    qf = centers[qp % nid] + 2.0 * rng.standard_normal((Q, d)).astype(np.float32)
    gf = centers[gp % nid] + 2.0 * rng.standard_normal((G, d)).astype(np.float32)
    gf[::13] = qf[rng.integers(0, Q, len(gf[::13]))]        # exact duplicates: ties
    single = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    ev = parallel.ShardedEvaluator()
    timers = {}
    res = ev.evaluate_host(torch.from_numpy(qf).pin_memory(), torch.from_numpy(gf).pin_memory(), qp, gp, qc, gc,
                           normalize=True, slab_rows=2048, timers=timers, query_groups=groups)
    assert timers["query_groups"] >= 2 if force else timers["query_groups"] == 1   # small pieces are merged
    np.testing.assert_array_equal(res.cmc, single.cmc)
    assert float(res.mAP) == float(single.mAP)
    np.testing.assert_array_equal(res.ap.cpu().numpy(), single.ap.cpu().numpy())
    np.testing.assert_array_equal(res.first.cpu().numpy(), single.first.cpu().numpy())
    r1, r2 = res.positive_ranks(), single.positive_ranks()
    for a, b in zip(r1, r2):
        np.testing.assert_array_equal(a, b)


def test_streamed_host_evaluation_no_query_id_in_gallery(monkeypatch):
    """Grouped flow with T = 0 (no query identity appears in the gallery): nothing to rank, zero
    valid queries -- the same answer as the device-resident evaluation."""
    from demo2_b200 import metrics, parallel
    monkeypatch.setattr(parallel, "MIN_PIECE_ROWS", 0)
    rng = np.random.default_rng(0)
    Q, G, d = 2100, 3000, 64
    qf = rng.standard_normal((Q, d)).astype(np.float32)
    gf = rng.standard_normal((G, d)).astype(np.float32)
    qp, gp = rng.integers(0, 50, Q), rng.integers(100, 150, G)
    qc, gc = rng.integers(0, 3, Q), rng.integers(0, 3, G)
    single = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    assert single.num_valid == 0
    timers = {}
    res = parallel.ShardedEvaluator().evaluate_host(torch.from_numpy(qf).pin_memory(), torch.from_numpy(gf).pin_memory(),
                                                    qp, gp, qc, gc, normalize=True, timers=timers, query_groups=2)
    assert timers["query_groups"] == 2 and res.num_valid == 0
    np.testing.assert_array_equal(res.cmc, single.cmc)
    np.testing.assert_array_equal(res.first.cpu().numpy(), single.first.cpu().numpy())


# ---------------------------------------------------------------------------------------------
# more thresholds than one window: slab path
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Q,G,nid,d", [(300, 3000, 4, 128), (700, 9000, 9, 256), (130, 5000, 2, 64)])
def test_fused_eval_long_threshold_lists(Q, G, nid, d):
    """Identities with hundreds to thousands of gallery images (R up to ~2 500 per query, mixed with
    short lists): the fused evaluation counts the flagged 256-row query blocks from a stored slab
    and everything else in the GEMM epilogue; integer counts == oracle on our own matrix."""
    from demo2_b200 import metrics
    rng = np.random.default_rng(Q + G)
    qf = rng.standard_normal((Q, d)).astype(np.float32)
    gf = rng.standard_normal((G, d)).astype(np.float32)
    gf[::11] = qf[rng.integers(0, Q, len(gf[::11]))]        # exact duplicates: ties
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    qp[Q // 2:] += 100                                       # half of the queries: rare identities ...
    gp[:40] = qp[Q // 2:Q // 2 + 40]                         # ... with a single gallery image each
    qc, gc = rng.integers(0, 3, Q), rng.integers(0, 3, G)
    plan = metrics.RankPlan(qp, gp, qc, gc)
    assert plan.max_cnt > 63
    res = metrics.evaluate_features(qf, gf, plan=plan, normalize=True)
    ours = metrics.sqdist_device(qf, gf, normalize=True).cpu().numpy()
    ap_o, first_o = _oracle_per_query(ours, qp, gp, qc, gc)
    np.testing.assert_array_equal(res.first.cpu().numpy(), first_o)
    np.testing.assert_allclose(res.ap.cpu().numpy(), ap_o, atol=1e-12)
    # the window fallback (workspace without a slab) gives the same integers
    from demo2_b200 import _lib
    from demo2_b200.metrics import _EvalWorkspace, check, ptr, stream_ptr
    w = _EvalWorkspace(Q, G, d, plan.T, matrix=False, max_cnt=0)
    q, g = metrics._features(qf), metrics._features(gf)
    check(_lib.load().demo_eval_features(ptr(q), ptr(g), Q, G, d, q.stride(0), g.stride(0), _lib.FLAG_L2NORM,
                                         ptr(plan.q_cam), ptr(plan.g_cam), ptr(plan.buf), plan.nbytes, plan.T,
                                         plan.max_cnt, 50, ptr(w.buf), w.nbytes, None, None, None, None, None, None,
                                         None, stream_ptr()))
    np.testing.assert_array_equal(w.view("first", torch.int32, Q).cpu().numpy(), res.first.cpu().numpy())
    np.testing.assert_array_equal(w.view("ap", torch.float64, Q).cpu().numpy(), res.ap.cpu().numpy())
