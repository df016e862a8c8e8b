"""World-size-2 (and 3) gloo tests of the gallery-sharded evaluation host logic on CPU.

The CUDA stages are replaced by a numpy engine built on the oracle (test infrastructure);
what is under test is demo2_b200.parallel: shard ranges, the all-gather / merge of per-rank
records, threshold CSR handling, the all-reduce of counts and the finalisation flow."""
from __future__ import annotations

import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import make_case, oracle

from demo2_b200 import parallel


class NumpyEngine:
    """Same stage interface as parallel.CudaEngine, computed with numpy on the host."""

    def plan(self, q_pid, g_pid, q_cam, g_cam):
        q_pid, g_pid = np.asarray(q_pid), np.asarray(g_pid)
        q_perm = np.argsort(q_pid, kind="stable")
        g_perm = np.argsort(g_pid, kind="stable")
        gs = g_pid[g_perm]
        lo = np.searchsorted(gs, q_pid[q_perm], "left")
        hi = np.searchsorted(gs, q_pid[q_perm], "right")
        cnt = hi - lo
        rec_ofs = np.concatenate([[0], np.cumsum(cnt)])
        return SimpleNamespace(Q=len(q_pid), G=len(g_pid), T=int(rec_ofs[-1]), max_cnt=int(cnt.max()),
                               q_perm=torch.from_numpy(q_perm.astype(np.int32)), g_perm=g_perm, g_lo=lo,
                               rec_ofs=torch.from_numpy(rec_ofs.astype(np.int32)),
                               q_cam=np.asarray(q_cam), g_cam=np.asarray(g_cam))

    def records(self, plan, qf, gf, g_index_base, normalize, g_index=None):
        qf, gf = np.asarray(qf, np.float32), np.asarray(gf, np.float32)
        gidx = np.asarray(g_index, np.int64) if g_index is not None else g_index_base + np.arange(len(gf))
        if normalize:
            qf, gf = oracle.l2_normalize(qf), oracle.l2_normalize(gf)
        distmat = oracle.euclidean_distance(qf, gf)
        recs = np.zeros((3, plan.T), np.int32)
        ofs = plan.rec_ofs.numpy()
        for i in range(plan.Q):
            q = plan.q_perm[i].item()
            n = ofs[i + 1] - ofs[i]
            cols = plan.g_perm[plan.g_lo[i]:plan.g_lo[i] + n]
            recs[0, ofs[i]:ofs[i + 1]] = distmat[q, cols].view(np.int32)
            recs[1, ofs[i]:ofs[i + 1]] = gidx[cols]
            recs[2, ofs[i]:ofs[i + 1]] = plan.g_cam[cols] == plan.q_cam[q]
        return SimpleNamespace(distmat=distmat, gidx=gidx), torch.from_numpy(recs)

    def thresholds(self, rec_ofs, recs, Q):
        ofs, recs = rec_ofs.numpy(), recs.numpy()
        T = recs.shape[1]
        thr_cnt = np.zeros(Q, np.int32)
        thr_val = np.zeros(max(T, 1), np.float32)
        thr_gidx = np.zeros(max(T, 1), np.int32)
        thr_junk = np.zeros(max(T, 1), np.int32)
        for i in range(Q):
            s, e = ofs[i], ofs[i + 1]
            d, g, j = recs[0, s:e].view(np.float32), recs[1, s:e], recs[2, s:e]
            order = np.lexsort((g, d))
            junk_before = np.cumsum(j[order]) - j[order]
            keep = order[j[order] == 0]
            n = len(keep)
            thr_cnt[i] = n
            thr_val[s:s + n], thr_gidx[s:s + n] = d[keep], g[keep]
            thr_junk[s:s + n] = junk_before[j[order] == 0]
        return tuple(torch.from_numpy(a) for a in (thr_cnt, thr_val, thr_gidx, thr_junk))

    def count(self, w, plan, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt):
        ofs, cnt = thr_ofs.numpy(), thr_cnt.numpy()
        tv, tg, out = thr_val.numpy(), thr_gidx.numpy(), counts.numpy()
        gidx = w.gidx
        for i in range(plan.Q):
            row = w.distmat[plan.q_perm[i].item()]
            for k in range(cnt[i]):
                t, p = tv[ofs[i] + k], tg[ofs[i] + k]
                out[ofs[i] + k] += int(np.sum((row < t) | ((row == t) & (gidx < p))))

    def finalize(self, thr_ofs, thr_cnt, thr_junk, counts, q_perm, Q, max_rank):
        ofs, cnt, junk, c = thr_ofs.numpy(), thr_cnt.numpy(), thr_junk.numpy(), counts.numpy()
        ap = np.full(Q, -1.0)
        first = np.zeros(Q, np.int32)
        for i in range(Q):
            n = cnt[i]
            if n == 0:
                continue
            r = 1 + c[ofs[i]:ofs[i] + n] - junk[ofs[i]:ofs[i] + n]
            q = q_perm[i].item()
            ap[q] = np.sum(np.arange(1, n + 1) / r) / n
            first[q] = r[0]
        valid = first > 0
        nv = int(valid.sum())
        cmc = np.array([(first[valid] <= k + 1).sum() for k in range(max_rank)], np.float32) / np.float32(max(nv, 1))
        scal = torch.zeros(4, dtype=torch.float64)
        scal[0] = ap[valid].mean() if nv else 0.0
        scal[1:2].view(torch.int32)[0] = nv
        return torch.from_numpy(cmc), scal, torch.from_numpy(ap), torch.from_numpy(first)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        qf, gf, qp, gp, qc, gc = case
        lo, hi = parallel.shard_range(len(gp), world, rank)
        ev = parallel.ShardedEvaluator(world=world, rank=rank, group=dist.group.WORLD, engine=NumpyEngine())
        res = ev.evaluate(qf, gf[lo:hi], qp, gp[lo:hi], qc, gc[lo:hi], g_index_base=lo, max_rank=50)
        out[rank] = (res.cmc, float(res.mAP), res.num_valid, res.ap.numpy(), res.first.numpy())
    finally:
        dist.destroy_process_group()


def _small_case():
    qf, gf, qp, gp, qc, gc = make_case("rgbnt201", 0, 4.0)
    qf, gf, qp, gp, qc, gc = qf[:60, :128].copy(), gf[:157, :128].copy(), qp[:60].copy(), gp[:157].copy(), qc[:60], gc[:157]
    gf[40:50] = gf[100:110]      # exact ties across shards
    qp[3] = 999                  # identity absent -> skipped
    return qf, gf, qp, gp, qc, gc


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_eval_matches_single_process(world):
    case = _small_case()
    qf, gf, qp, gp, qc, gc = case
    dist_full = oracle.euclidean_distance(qf, gf)
    cmc_o, mAP_o = oracle.eval_func(dist_full, qp, gp, qc, gc)
    ofs_o = oracle.rank_counts(dist_full, qp, gp, qc, gc)[0]
    nv_o = int((np.diff(ofs_o) > 0).sum())
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        cmc, mAP, nv, ap, first = out[r]
        np.testing.assert_allclose(cmc, cmc_o, atol=1e-7)
        assert abs(mAP - mAP_o) < 1e-12
        assert nv == nv_o and first[3] == 0
    # identical on every rank (bitwise)
    for r in range(1, world):
        np.testing.assert_array_equal(out[0][3], out[r][3])
        np.testing.assert_array_equal(out[0][4], out[r][4])


def test_single_rank_flow_and_shard_ranges():
    for G, world in [(10, 3), (8575, 8), (5, 8), (1000000, 8)]:
        spans = [parallel.shard_range(G, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == G
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    qf, gf, qp, gp, qc, gc = _small_case()
    ev = parallel.ShardedEvaluator(engine=NumpyEngine())
    res = ev.evaluate(qf, gf, qp, gp, qc, gc)
    cmc_o, mAP_o = oracle.eval_func(oracle.euclidean_distance(qf, gf), qp, gp, qc, gc)
    np.testing.assert_allclose(res.cmc, cmc_o, atol=1e-7)
    assert abs(res.mAP - mAP_o) < 1e-12


def test_merge_records_layout():
    cnt = torch.tensor([[2, 0, 1], [1, 3, 0]], dtype=torch.int32)
    recs = torch.zeros((2, 3, 4), dtype=torch.int32)
    recs[0, 1, :3] = torch.tensor([10, 11, 12])
    recs[1, 1, :4] = torch.tensor([20, 21, 22, 23])
    ofs, merged, T, max_cnt = parallel.merge_records(cnt, recs)
    assert ofs.tolist() == [0, 3, 6, 7] and T == 7 and max_cnt == 3
    assert merged[1].tolist() == [10, 11, 20, 21, 22, 23, 12]


# ---------------------------------------------------------------------------------------------
# row-sharded re-ranking (parallel.ShardedReranker): gloo, numpy stage engine
# ---------------------------------------------------------------------------------------------
class NumpyRerankEngine:
    """Stage interface of parallel.CudaRerankEngine on the host, from the oracle's pieces.
    (Every rank forms the full all-pairs matrix -- cheap at test size -- and keeps its rows of
    E = D^T, so that E[i][j] / rowmax_i is exactly the oracle's od[i][j].)"""
    F16, F32 = np.float16, np.float32

    def dims(self, N, k1, k2):
        kh = int(np.around(k1 / 2)) + 1
        cap = (k1 + 1) * (kh + 1)
        return max(k1 + 1, k2), cap, min(N, max(k2, 1) * cap)

    def begin(self, feat, Q, k1, k2, row0, nrows, rows_cap, normalize):
        f = np.asarray(feat, np.float32)
        if normalize:
            f = oracle.l2_normalize(f)
        D = oracle.euclidean_distance(f, f)
        self.N, self.Q, self.k1, self.k2, self.row0, self.nrows = len(f), Q, k1, k2, row0, nrows
        E = np.ascontiguousarray(D[:, row0:row0 + nrows].T)
        self.od = (E / E.max(axis=1, keepdims=True)).astype(np.float32) if nrows else E
        return torch.device("cpu")

    def topk(self, rank_rows):
        if self.nrows:
            rank_rows.numpy()[:self.nrows] = oracle.stable_topk(self.od, max(self.k1 + 1, self.k2))

    def krecip(self, rank_all, v_idx, v_val, v_cnt):
        rows = oracle.k_reciprocal_rows(self.od, rank_all.numpy(), self.k1, row0=self.row0)
        for li, (idx, val) in enumerate(rows):
            i = self.row0 + li
            keep = val != 0
            n = int(keep.sum())
            v_idx.numpy()[i, :n], v_val.numpy()[i, :n], v_cnt.numpy()[i] = idx[keep], val[keep], n

    def expand(self, rank_all, v_idx, v_val, v_cnt, q_idx, q_val, q_cnt):
        vi, vv, vc, rk = v_idx.numpy(), v_val.numpy(), v_cnt.numpy(), rank_all.numpy()
        for i in range(self.row0, self.row0 + self.nrows):
            acc = {}
            for nb in rk[i, :self.k2]:
                for c, v in zip(vi[nb, :vc[nb]].tolist(), vv[nb, :vc[nb]].astype(np.float32).tolist()):
                    acc[c] = np.float32(acc.get(c, np.float32(0.0)) + np.float32(v))
            cols = np.fromiter(sorted(acc), dtype=np.int64, count=len(acc))
            vals = (np.array([acc[c] for c in cols.tolist()], np.float32) / np.float32(self.k2)).astype(np.float16)
            keep = vals != 0
            n = int(keep.sum())
            q_idx.numpy()[i, :n], q_val.numpy()[i, :n], q_cnt.numpy()[i] = cols[keep], vals[keep], n

    def jaccard(self, f_idx, f_val, f_cnt, lambda_value, out_rows):
        fi, fv, fc = f_idx.numpy(), f_val.numpy(), f_cnt.numpy()
        n, Q = self.N, self.Q
        inv_rows = [[] for _ in range(n)]
        inv_vals = [[] for _ in range(n)]
        for t in range(n):
            for c, v in zip(fi[t, :fc[t]].tolist(), fv[t, :fc[t]].tolist()):
                inv_rows[c].append(t)
                inv_vals[c].append(v)
        inv_rows = [np.asarray(a, np.int64) for a in inv_rows]
        inv_vals = [np.asarray(a, np.float16) for a in inv_vals]
        for i in range(self.row0, min(self.row0 + self.nrows, Q)):
            temp_min = np.zeros(n, dtype=np.float16)
            for j, vij in zip(fi[i, :fc[i]].tolist(), fv[i, :fc[i]]):
                t = inv_rows[j]
                temp_min[t] = temp_min[t] + np.minimum(vij, inv_vals[j])
            jac = 1 - temp_min / (2 - temp_min)
            final = jac * (1 - lambda_value) + self.od[i - self.row0] * lambda_value
            out_rows.numpy()[i - self.row0] = final[Q:].astype(np.float32)


def _rr_worker(rank, world, port, qf, gf, k1, k2, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rr = parallel.ShardedReranker(world=world, rank=rank, group=dist.group.WORLD, engine_factory=NumpyRerankEngine)
        out[rank] = rr.re_ranking(torch.from_numpy(qf), torch.from_numpy(gf), k1, k2, 0.3).numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k1,k2", [(2, 8, 3), (3, 6, 1)])
def test_sharded_reranking_matches_oracle(world, k1, k2):
    qf, gf, *_ = make_case("rgbnt201", 1, 5.0)
    qf, gf = qf[:37, :96].copy(), gf[:90, :96].copy()      # N = 127: ragged row shards
    ref = oracle.re_ranking(qf, gf, k1, k2, 0.3)
    # the same flow without a process group (ranks emulated one after another)
    emu = parallel.ShardedReranker(engine_factory=NumpyRerankEngine).re_ranking(
        torch.from_numpy(qf), torch.from_numpy(gf), k1, k2, 0.3, emulate_ranks=world).numpy()
    np.testing.assert_array_equal(emu, ref)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rr_worker, args=(world, _free_port(), qf, gf, k1, k2, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        np.testing.assert_array_equal(out[r], ref)


# ---------------------------------------------------------------------------------------------
# evaluation under DDP: every rank keeps the features it extracted (parallel.DistributedR1mAP)
# ---------------------------------------------------------------------------------------------
def _ddp_worker(rank, world, port, feats, pids, cams, num_query, out, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ev = parallel.DistributedR1mAP(num_query, world=world, rank=rank, group=dist.group.WORLD, feat_norm=False,
                                       engine=NumpyEngine())
        n = len(pids)
        if mode == "sampler":
            # the real thing: DistributedSampler(shuffle=False) pads the index list with its head when
            # len(dataset) % world != 0, so some samples are extracted twice
            from torch.utils.data import DistributedSampler
            mine = np.asarray(list(DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=False)))
        elif mode == "queries_only_rank":
            # rank 0 extracted nothing but queries (no gallery shard of its own)
            mine = np.arange(0, num_query // 2) if rank == 0 else np.arange(num_query // 2, n)
        else:
            mine = np.arange(rank, n, world)                 # interleaving without padding
        for s in range(0, len(mine), 16):
            b = mine[s:s + 16]
            ev.update((torch.from_numpy(feats[b]), pids[b], torch.from_numpy(cams[b]), torch.from_numpy(b)))
        cmc, mAP = ev.compute()
        out[rank] = (cmc, float(mAP), ev.last_result.ap.numpy(), ev.last_result.first.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "interleave"), (3, "sampler"), (2, "queries_only_rank")])
def test_ddp_evaluator_matches_rank0_evaluation(world, mode):
    """Every rank keeps what it extracted; the result equals the single-process evaluation of the
    concatenated features EXACTLY (ties follow the dataset order of the gallery), also when the
    sampler repeated samples (len % world != 0) and when a rank holds no gallery item."""
    qf, gf, qp, gp, qc, gc = _small_case()
    feats = np.concatenate([qf, gf])
    pids, cams = np.concatenate([qp, gp]), np.concatenate([qc, gc])
    assert len(pids) % 3 != 0                                # the sampler case really pads
    dist_full = oracle.euclidean_distance(qf, gf)
    cmc_o, mAP_o = oracle.eval_func(dist_full, qp, gp, qc, gc)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), feats, pids, cams, len(qp), out, mode), nprocs=world, join=True)
    ofs, idx, r, c = oracle.rank_counts(dist_full, qp, gp, qc, gc)
    first_o = np.array([r[ofs[q]:ofs[q + 1]].min() if ofs[q + 1] > ofs[q] else 0 for q in range(len(qp))])
    for rk in range(world):
        np.testing.assert_allclose(out[rk][0], cmc_o, atol=1e-7)
        assert abs(out[rk][1] - mAP_o) < 1e-12
        np.testing.assert_array_equal(out[rk][3], first_o)
    assert all(out[0][1] == out[rk][1] for rk in range(world))


def test_streamed_evaluation_host_logic():
    """Pure host logic of the streamed evaluation: query-block boundaries (multiples of 1024 rows,
    first 0, last Q, strictly increasing) and the slab size rule (at least six slabs for small
    shards, 8 192 .. 131 072 rows, an explicit value wins)."""
    from demo2_b200.parallel import ShardedEvaluator as SE
    for Q, groups in [(20000, 4), (2600, 4), (2048, 2), (1500, 4), (4200, 3), (1024, 8)]:
        qb = SE._query_bounds(Q, groups)
        assert qb[0] == 0 and qb[-1] == Q and all(b > a for a, b in zip(qb, qb[1:]))
        assert all(b % 1024 == 0 for b in qb[1:-1]) and len(qb) - 1 <= max(1, min(groups, Q // 1024))
    assert SE._query_bounds(20000, 4) == [0, 5120, 10240, 15360, 20000]
    assert SE._slab_rows(None, 669000) == 111616 and SE._slab_rows(None, 84000) == 14080
    assert SE._slab_rows(None, 100) == 8192 and SE._slab_rows(None, 10 ** 7) == 131072
    assert SE._slab_rows(500, 10 ** 6) == 500
