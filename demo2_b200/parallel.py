"""Gallery-sharded evaluation: one process per GPU, gallery rows split contiguously over the
ranks, queries replicated.

Rank counts are sums over gallery items, hence additive over any gallery partition:

    every rank   plan (labels) -> records of its local same-identity pairs      [tcgen05 extract GEMM]
    all ranks    all-gather the records (KB..MB) and merge them per query       [torch.distributed]
    every rank   thresholds of the merged records; counts of its LOCAL gallery  [tcgen05 count GEMM]
    all ranks    all-reduce(sum) of the counts                                  [torch.distributed]
    every rank   finalize -> identical CMC / mAP on all ranks

The collectives move a few MB once per evaluation, so they are issued through
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests); the compute stages are the
C-ABI entry points of libdemo_b200.  The record merge is plain index arithmetic on tensors and
runs on whatever device the engine produces (this is what the world_size-2 gloo tests cover,
with a numpy engine standing in for the CUDA stages).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


_NVTX = bool(os.environ.get("DEMO_NVTX"))


class _Phases:
    """NVTX ranges around the stages of an evaluation (DEMO_NVTX=1; for nsys timelines of the
    multi-GPU step: plan / records / exchange / thresholds / count / allreduce / finalize)."""

    def __init__(self):
        self.open = False

    def __call__(self, name=None):
        if not _NVTX:
            return
        if self.open:
            torch.cuda.nvtx.range_pop()
        if name:
            torch.cuda.nvtx.range_push("demo:" + name)
        self.open = bool(name)


def shard_range(G: int, world: int, rank: int):
    """Contiguous gallery shard [lo, hi) of `rank`."""
    base, rem = divmod(G, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_records(cnt_all: torch.Tensor, recs_all: torch.Tensor, sizes=None):
    """cnt_all [P, Q] records per (rank, sorted query); recs_all [P, 3, Tmax] int32 rows
    (dist bits, gidx, junk) in each rank's local CSR order.  Returns the merged CSR offsets
    [Q+1] (int32), the merged [3, T_total] records (per query: rank 0's records first), T_total
    and the largest per-query count.  Index arithmetic on whole tensors: no per-rank loop and a
    single device->host read (skipped when the caller passes ``sizes`` = (T_total, max_cnt))."""
    P, Q = cnt_all.shape
    dev = cnt_all.device
    cnt_all = cnt_all.to(torch.int64)
    total = cnt_all.sum(0)
    ofs = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
    ofs[1:] = torch.cumsum(total, 0)
    if sizes is None:
        host = torch.stack([ofs[-1], total.max() if Q else ofs[-1]]).cpu()
        T, max_cnt = int(host[0]), int(host[1])
    else:
        T, max_cnt = sizes
    merged = torch.zeros((3, max(T, 1)), dtype=torch.int32, device=recs_all.device)
    if T == 0:
        return ofs.to(torch.int32), merged, T, max_cnt
    t_max = recs_all.shape[2]
    cnt_flat = cnt_all.reshape(-1)                                   # (rank, query) pairs, rank-major
    start_flat = torch.cumsum(cnt_flat, 0) - cnt_flat                # first record of the pair in rank-major order
    pair = torch.repeat_interleave(torch.arange(P * Q, device=dev), cnt_flat, output_size=T)
    within = torch.arange(T, device=dev) - start_flat[pair]
    r, q = pair // Q, pair % Q
    rank_prefix = (torch.cumsum(cnt_all, 0) - cnt_all).reshape(-1)   # records of lower ranks for the same query
    rank_start = start_flat.reshape(P, Q)[:, 0]                      # first record of every rank
    dest = ofs[:-1][q] + rank_prefix[pair] + within
    src = r * t_max + (torch.arange(T, device=dev) - rank_start[r])
    merged[:, dest] = recs_all.permute(1, 0, 2).reshape(3, P * t_max)[:, src]
    return ofs.to(torch.int32), merged, T, max_cnt


# ------------------------------------------------------------------------------------------------
# collectives
# ------------------------------------------------------------------------------------------------
class TorchCollectives:
    """torch.distributed (NCCL on GPUs, gloo in the CPU tests).  With a gloo group and CUDA tensors
    (two processes sharing one GPU in the GPU tests) the payload is staged through the host."""

    name = "torch.distributed"

    def __init__(self, world, rank, group=None):
        import torch.distributed as dist
        self.world, self.rank, self.group, self.dist = world, rank, group, dist
        self.stage = dist.get_backend(group) == "gloo"

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        flat = t.contiguous().view(-1)
        src = flat.cpu() if (self.stage and flat.is_cuda) else flat
        out = torch.empty(self.world * src.numel(), dtype=src.dtype, device=src.device)
        self.dist.all_gather_into_tensor(out, src, group=self.group)
        return out.to(t.device).view((self.world,) + tuple(t.shape))

    def all_reduce_sum(self, t: torch.Tensor):
        if self.stage and t.is_cuda:
            h = t.cpu()
            self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM, group=self.group)
            t.copy_(h)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def check(self):
        pass


class LibCollectives:
    """The library's own NCCL communicator (C ABI demo_comm_*): both collectives of an evaluation
    are enqueued on the compute stream right behind the kernels that produce their inputs.  One
    communicator per process; the 128-byte id travels through the torch.distributed group."""

    name = "libdemo_b200 NCCL communicator"
    _ready = None   # (world, rank) once initialised

    def __init__(self, world, rank, group=None):
        import torch.distributed as dist
        self.world, self.rank = world, rank
        self.lib = _lib.require_device()
        if LibCollectives._ready is None:
            if not self.lib.demo_comm_available():
                raise _lib.DemoError("libnccl not available: " + self.lib.demo_last_error().decode())
            ident = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                check(self.lib.demo_comm_unique_id(C.c_void_p(ident.data_ptr())))
            dev = torch.device("cuda", torch.cuda.current_device())
            on_dev = dist.get_backend(group) != "gloo"
            t = ident.to(dev) if on_dev else ident
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = t.cpu().contiguous()
            check(self.lib.demo_comm_init(rank, world, C.c_void_p(ident.data_ptr())))
            LibCollectives._ready = (world, rank)
        elif LibCollectives._ready != (world, rank):
            raise _lib.DemoError("the library communicator belongs to rank %d of %d" % LibCollectives._ready[::-1])

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        src = t.contiguous()
        out = torch.empty((self.world,) + tuple(src.shape), dtype=src.dtype, device=src.device)
        check(self.lib.demo_comm_all_gather(ptr(src), ptr(out), src.numel() * src.element_size(), stream_ptr()))
        return out

    def all_reduce_sum(self, t: torch.Tensor):
        assert t.dtype in (torch.int32, torch.uint32) and t.is_contiguous()
        check(self.lib.demo_comm_all_reduce_sum_u32(ptr(t), t.numel(), stream_ptr()))

    def check(self):
        check(self.lib.demo_comm_check())

    @staticmethod
    def destroy():
        if LibCollectives._ready is not None:
            check(_lib.load().demo_comm_destroy())
            LibCollectives._ready = None


def make_collectives(world, rank, group=None, prefer_library: bool = True):
    """In-library NCCL when the group runs on NCCL (one GPU per process), torch.distributed
    otherwise (gloo: CPU tests, or two test processes sharing one GPU)."""
    import torch.distributed as dist
    if prefer_library and torch.cuda.is_available() and dist.get_backend(group) == "nccl":
        try:
            return LibCollectives(world, rank, group)
        except Exception as exc:  # keep evaluating through torch.distributed, but say so
            import sys
            print("demo2_b200: library communicator unavailable (%s); using torch.distributed" % exc, file=sys.stderr)
    return TorchCollectives(world, rank, group)


# ------------------------------------------------------------------------------------------------
# compute stages
# ------------------------------------------------------------------------------------------------
NO_PID = -2 ** 31   # pid of the placeholder gallery row of a rank without gallery items
MIN_PIECE_ROWS = 32768   # grouped streamed evaluation: fewest queried gallery rows per query block worth a separate piece


class CudaEngine:
    """The compute stages on one GPU (C ABI of libdemo_b200)."""

    def __init__(self):
        self.lib = _lib.require_device()

    def plan(self, q_pid, g_pid, q_cam, g_cam):
        from .metrics import RankPlan
        return RankPlan(q_pid, g_pid, q_cam, g_cam, defer=True)

    def workspace(self, plan, d, max_cnt):
        from .metrics import _EvalWorkspace
        return _EvalWorkspace(plan.Q, plan.G, d, plan.T, matrix=False, max_cnt=max_cnt)

    def prepare(self, plan, w, x, which, row0, nrows, normalize, host_input=False):
        """pid-sorted rows [row0, row0 + nrows) of the queries (which=0) / gallery (which=1); x is a
        device tensor, or a PINNED host tensor read in place over PCIe (host_input)."""
        flags = (_lib.FLAG_L2NORM if normalize else 0) | (_lib.FLAG_HOST_INPUT if host_input else 0)
        check(self.lib.demo_eval_prepare(ptr(x), x.shape[0], x.shape[1], x.stride(0), flags, which, int(row0),
                                         int(nrows), ptr(plan.buf), plan.nbytes, plan.Q, plan.G, plan.T, ptr(w.buf),
                                         w.nbytes, None, stream_ptr()))

    def extract(self, plan, w, g_index_base, g_index=None, q_row0=0, q_nrows=None):
        q_nrows = plan.Q - q_row0 if q_nrows is None else q_nrows
        check(self.lib.demo_eval_extract(plan.Q, plan.G, w.d, ptr(plan.q_cam), ptr(plan.g_cam), int(g_index_base),
                                         ptr(g_index), ptr(plan.buf), plan.nbytes, plan.T, ptr(w.buf), w.nbytes,
                                         None, None, None, int(q_row0), int(q_nrows), stream_ptr()))
        n = max(plan.T, 1)
        recs = torch.stack([w.view("rec_dist", torch.float32, n).view(torch.int32),
                            w.view("rec_gidx", torch.int32, n), w.view("rec_junk", torch.int32, n)])
        return recs[:, :plan.T]

    def records(self, plan, qf, gf, g_index_base, normalize, g_index=None, max_cnt=0):
        from .metrics import _features
        q, g = _features(qf), _features(gf)
        w = self.workspace(plan, q.shape[1], max(max_cnt, plan.max_cnt))
        self.prepare(plan, w, q, 0, 0, plan.Q, normalize)
        self.prepare(plan, w, g, 1, 0, plan.G, normalize)
        return w, self.extract(plan, w, g_index_base, g_index)

    def thresholds(self, rec_ofs, recs, Q):
        T = recs.shape[1]
        dev = recs.device
        n = max(T, 1)
        thr_cnt = torch.empty(Q, dtype=torch.int32, device=dev)
        thr_val = torch.empty(n, dtype=torch.float32, device=dev)
        thr_gidx = torch.empty(n, dtype=torch.int32, device=dev)
        thr_junk = torch.empty(n, dtype=torch.int32, device=dev)
        recs = recs.contiguous()
        check(self.lib.demo_build_thresholds(ptr(rec_ofs), ptr(recs[0]), ptr(recs[1]), ptr(recs[2]), Q,
                                             ptr(thr_cnt), ptr(thr_val), ptr(thr_gidx), ptr(thr_junk), stream_ptr()))
        return thr_cnt, thr_val, thr_gidx, thr_junk

    def thresholds_alloc(self, T, Q, dev):
        n = max(int(T), 1)
        return (torch.empty(Q, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev))

    def thresholds_into(self, rec_ofs, recs, thr, q_row0, q_nrows):
        """Thresholds of the pid-sorted queries [q_row0, q_row0 + q_nrows) into preallocated arrays
        (rec_ofs holds absolute offsets, so a query block is a shifted view)."""
        thr_cnt, thr_val, thr_gidx, thr_junk = thr
        check(self.lib.demo_build_thresholds(ptr(rec_ofs[q_row0:]), ptr(recs[0]), ptr(recs[1]), ptr(recs[2]),
                                             int(q_nrows), ptr(thr_cnt[q_row0:]), ptr(thr_val), ptr(thr_gidx),
                                             ptr(thr_junk), stream_ptr()))

    def count(self, w, plan, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt, g_row0=0, g_nrows=None,
              reserve_sms=0, q_row0=0, q_nrows=None):
        g_nrows = w.G - g_row0 if g_nrows is None else g_nrows
        q_nrows = w.Q - q_row0 if q_nrows is None else q_nrows
        check(self.lib.demo_eval_count_range(w.Q, w.G, w.d, plan.T, ptr(w.buf), w.nbytes, ptr(thr_ofs), ptr(thr_cnt),
                                             ptr(thr_val), ptr(thr_gidx), ptr(counts), int(max_cnt), 0, int(g_row0),
                                             int(g_nrows), int(q_row0), int(q_nrows), int(reserve_sms), stream_ptr()))

    def finalize(self, thr_ofs, thr_cnt, thr_junk, counts, q_perm, Q, max_rank):
        dev = counts.device
        cmc = torch.empty(max_rank, dtype=torch.float32, device=dev)
        scal = torch.empty(4, dtype=torch.float64, device=dev)  # [mAP, nvalid(int bits), -, -]
        ap = torch.empty(Q, dtype=torch.float64, device=dev)
        first = torch.empty(Q, dtype=torch.int32, device=dev)
        scratch = torch.empty(4096, dtype=torch.float64, device=dev)
        nvalid = scal[1:2].view(torch.int32)
        check(self.lib.demo_cmc_map_finalize(ptr(thr_ofs), ptr(thr_cnt), ptr(thr_junk), ptr(counts), ptr(q_perm), Q,
                                             max_rank, ptr(cmc), ptr(scal), ptr(nvalid), ptr(ap), ptr(first),
                                             ptr(scratch), stream_ptr()))
        return cmc, scal, ap, first

    def launches(self, count_launches=1, slab_blocks=0):
        # plan: iota, gallery key / unkey, ranges, band list | 2x prep, gidx, fill records, extract GEMM |
        # thresholds (short + long lists) | per count launch: count GEMM, tie resolver, conditional tie-fix GEMM (+ per
        # flagged 256-row block: store GEMM + 2 streaming count kernels) | per-query AP, reduce
        # (CUB sort/scan kernels and memsets not counted)
        return 5 + 5 + 2 + 3 * count_launches + 3 * slab_blocks + 2

    @staticmethod
    def event():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e


class ShardedEvaluator:
    """Gallery-sharded evaluation of one rank.  Per evaluation there is ONE host round trip before
    the final read of the metrics: the plan's sizes (and, with several ranks, every rank's sizes,
    all-gathered on the device first) are read together; everything else is enqueued on the
    stream -- records, the fixed-size record exchange, thresholds, the count GEMM, the all-reduce
    of the counts and the finalisation."""

    def __init__(self, world: int = 1, rank: int = 0, group=None, engine=None, collectives=None):
        self.world, self.rank, self.group = world, rank, group
        self.engine = engine if engine is not None else CudaEngine()
        if collectives is None and world > 1:
            if isinstance(self.engine, CudaEngine):
                collectives = make_collectives(world, rank, group)
            else:
                collectives = TorchCollectives(world, rank, group)
        self.coll = collectives

    # ---- label-only part: plan, [sizes of every rank], one host read ---------------------------
    @staticmethod
    def _query_bounds(Q, groups):
        """Boundaries of the pid-sorted query blocks of the grouped streamed evaluation (multiples of
        1024 rows: the slab granularity of the count path)."""
        K = max(1, min(int(groups), Q // 1024))
        qb = [0] + [min(Q, -(-(j * Q) // (K * 1024)) * 1024) for j in range(1, K)] + [Q]
        return sorted(set(qb))

    def _plan_and_sizes(self, q_pid, g_pid_local, q_cam, g_cam_local, groups=None, local_ok=True):
        """groups (several ranks, streamed evaluation): also returns, from the same host read, what the
        grouped flow needs per query block -- local record offsets, the end of the local gallery piece
        the block asks for, the longest per-rank record slice and the global record offset -- and
        whether EVERY rank can take that flow (sizes["groups"]["ok"])."""
        eng = self.engine
        n_local = int(len(g_pid_local))
        if n_local == 0:
            # a rank without gallery items still takes part in the exchange: placeholder row that
            # matches no query, no records, no counting
            dev = g_pid_local.device if isinstance(g_pid_local, torch.Tensor) else None
            g_pid_local = torch.full((1,), NO_PID, dtype=torch.int32, device=dev)
            g_cam_local = torch.zeros(1, dtype=torch.int32, device=dev)
        plan = eng.plan(q_pid, g_pid_local, q_cam, g_cam_local)
        Q = plan.Q
        deferred = plan.T is None
        if deferred:
            info = plan.info
        else:
            info = torch.tensor([plan.T, plan.max_cnt, 0, 0], dtype=torch.int32, device=plan.rec_ofs.device)
        sizes = None
        if self.world > 1:
            cnt_local = (plan.rec_ofs[1:] - plan.rec_ofs[:-1]).to(torch.int32)
            tail = [n_local] + ([1 if (local_ok and n_local > 0) else 0] if groups else [])
            meta = torch.cat([cnt_local, info.to(torch.int32),
                              torch.tensor(tail, dtype=torch.int32, device=cnt_local.device)])
            gathered = self.coll.all_gather(meta)                                   # [P, Q + 5 (+ 1)]
            cnt_all = gathered[:, :Q]
            tot_max = cnt_all.sum(0).max().reshape(1).to(torch.int64) if Q else torch.zeros(1, dtype=torch.int64)
            P, W = self.world, 5 + (1 if groups else 0)
            extra, qb = [], None
            if groups:
                qb = self._query_bounds(Q, groups)
                K = len(qb) - 1
                idx = torch.tensor(qb, dtype=torch.int64, device=gathered.device)
                last = idx[1:] - 1
                rec_ofs = plan.rec_ofs.to(gathered.device)
                ends = plan.g_lo.to(gathered.device)[last] + rec_ofs[last + 1] - rec_ofs[last]
                pre = torch.zeros((P, Q + 1), dtype=torch.int64, device=gathered.device)
                pre[:, 1:] = torch.cumsum(cnt_all.to(torch.int64), 1)
                at = pre[:, idx]                                                    # [P, K + 1]
                extra = [rec_ofs[idx].to(torch.int64), ends.to(torch.int64),
                         (at[:, 1:] - at[:, :-1]).max(0).values, at.sum(0)]
            host = torch.cat([gathered[:, Q:].reshape(-1).to(torch.int64), tot_max.to(gathered.device)] + extra).cpu()
            per_rank = host[:P * W].reshape(P, W)
            if deferred:
                plan.finish(per_rank[self.rank, :4])
            sizes = dict(cnt_all=cnt_all, t_max=int(per_rank[:, 0].max()), T_sum=int(per_rank[:, 0].sum()),
                         G_total=int(per_rank[:, 4].sum()), max_all=int(host[P * W]))
            if groups:
                x = host[P * W + 1:].tolist()
                # worth it only while every rank's pieces stay large: 4 ranks (21k queried rows per piece)
                # measured 57.6 ms grouped against 47.8 ms plain, 8 ranks 46.4 against 40.7; 2 ranks
                # (41k rows per piece) 93 against 99 ms
                big = bool(per_rank[:, 3].min() >= MIN_PIECE_ROWS * K)
                sizes["groups"] = dict(qb=qb, rec_lo=x[:K + 1], ends=x[K + 1:2 * K + 1], gmax=x[2 * K + 1:3 * K + 1],
                                       gofs=x[3 * K + 1:4 * K + 2],
                                       ok=bool(per_rank[:, 5].min() > 0) and K > 1 and big)
        elif deferred:
            plan.finish(info.cpu())
        return plan, n_local, sizes

    def _exchange(self, plan, recs, sizes):
        if self.world > 1:
            padded = torch.zeros((3, max(sizes["t_max"], 1)), dtype=torch.int32, device=recs.device)
            padded[:, :recs.shape[1]] = recs
            recs_all = self.coll.all_gather(padded)
            return merge_records(sizes["cnt_all"], recs_all, sizes=(sizes["T_sum"], sizes["max_all"]))
        return plan.rec_ofs, recs, plan.T, plan.max_cnt

    def _finish(self, plan, thr, counts, thr_ofs, G_total, max_rank):
        from .metrics import EvalResult
        thr_cnt, thr_val, thr_gidx, thr_junk = thr
        if G_total < max_rank:
            max_rank = G_total
        cmc_d, scal_d, ap, first = self.engine.finalize(thr_ofs, thr_cnt, thr_junk, counts, plan.q_perm, plan.Q,
                                                        max_rank)
        # one D2H copy for cmc | mAP | num_valid
        slab = torch.cat([scal_d.view(torch.uint8), cmc_d.view(torch.uint8)]).cpu()
        mAP = np.float64(slab[:8].view(torch.float64)[0].item())
        nvalid = int(slab[8:12].view(torch.int32)[0].item())
        cmc = slab[32:].view(torch.float32).numpy().copy()
        if self.coll is not None:
            self.coll.check()
        return EvalResult(cmc, mAP, nvalid, ap, first,
                          detail={"thr_ofs": thr_ofs, "thr_cnt": thr_cnt, "thr_gidx": thr_gidx, "thr_junk": thr_junk,
                                  "counts": counts, "q_perm": plan.q_perm})

    def evaluate(self, qf, gf_local, q_pid, g_pid_local, q_cam, g_cam_local, g_index_base: int = 0,
                 normalize: bool = False, max_rank: int = 50, timers: dict | None = None, g_index=None):
        """Features on the device (or anything metrics._features accepts).  g_index (optional):
        global gallery index of every local gallery row -- the tie-break key -- instead of
        g_index_base + local row."""
        eng = self.engine
        ev = getattr(eng, "event", None) if timers is not None else None
        mark = (lambda: ev()) if ev else (lambda: None)
        ph = _Phases()

        t0 = mark()
        ph("plan")
        plan, n_local, sizes = self._plan_and_sizes(q_pid, g_pid_local, q_cam, g_cam_local)
        Q = plan.Q
        t1 = mark()
        ph("records")
        max_all = sizes["max_all"] if sizes else plan.max_cnt
        if n_local > 0:
            kw = {}
            if g_index is not None:
                kw["g_index"] = g_index
            if isinstance(eng, CudaEngine):
                kw["max_cnt"] = max_all
            w, recs = eng.records(plan, qf, gf_local, g_index_base, normalize, **kw)
        else:
            w, recs = None, torch.zeros((3, 0), dtype=torch.int32, device=plan.rec_ofs.device)
        t2 = mark()
        ph("exchange")
        thr_ofs, merged, T, max_cnt = self._exchange(plan, recs, sizes)
        t3 = mark()
        ph("thresholds")
        thr = eng.thresholds(thr_ofs, merged, Q)
        counts = torch.zeros(max(T, 1), dtype=torch.int32, device=merged.device)
        t4 = mark()
        ph("count")
        if T > 0 and n_local > 0:
            eng.count(w, plan, thr_ofs, thr[0], thr[1], thr[2], counts, max_cnt)
        t5 = mark()
        ph("allreduce")
        if self.world > 1:
            self.coll.all_reduce_sum(counts)
        t6 = mark()
        ph("finalize")
        res = self._finish(plan, thr, counts, thr_ofs, sizes["G_total"] if sizes else n_local, max_rank)
        t7 = mark()
        ph()
        if timers is not None and ev:
            timers.update({"plan": (t0, t1), "records": (t1, t2), "exchange": (t2, t3), "thresholds": (t3, t4),
                           "count": (t4, t5), "allreduce": (t5, t6), "finalize": (t6, t7)})
            timers["launches"] = eng.launches(1, -(-Q // 256) if max_cnt > 63 else 0)
        return res

    # ---- host-resident inputs: the gallery is pulled in slab by slab while the GEMM ranks --------
    def evaluate_host(self, q_host, g_host_local, q_pid, g_pid_local, q_cam, g_cam_local, g_index_base: int = 0,
                      normalize: bool = False, max_rank: int = 50, timers: dict | None = None,
                      slab_rows: int | None = None, shard_query_upload: bool = True, reserve_sms: int | None = None,
                      query_groups: int | None = None):
        """One evaluation whose features live in PINNED HOST memory (what R1_mAP_eval.update
        accumulates when the model runs elsewhere, utils/metrics.py:244).  No fp32 copy of the
        gallery is made on the device: demo_eval_prepare pulls the rows over PCIe in pid-sorted
        order.  The rows some query asks for come first (demo_eval_plan) -- once they are in, the
        records, thresholds and the count GEMM of that slab run while a side stream pulls in the
        remaining slabs; every slab is counted as soon as it has landed (rank counts are additive
        over gallery partitions).  With several ranks each rank uploads 1/P of the queries and
        the rest arrives over NVLink (all-gather)."""
        eng = self.engine
        assert isinstance(eng, CudaEngine), "evaluate_host needs the CUDA engine"
        if reserve_sms is None:
            reserve_sms = int(os.environ.get("DEMO_RESERVE_SMS", "0"))   # env: experiments
        if query_groups is None:
            query_groups = int(os.environ.get("DEMO_QUERY_GROUPS", "4"))
        if (self.world == 1 and query_groups > 1 and len(g_pid_local) > 0 and q_host.shape[0] >= 2 * 1024
                and isinstance(g_host_local, torch.Tensor) and not g_host_local.is_cuda and g_host_local.is_pinned()
                and g_host_local.dtype == torch.float32 and g_host_local.dim() == 2 and g_host_local.stride(1) == 1
                and g_host_local.shape[1] % 4 == 0 and g_host_local.shape[1] <= 2048):
            return self._evaluate_host_grouped(q_host, g_host_local, q_pid, g_pid_local, q_cam, g_cam_local,
                                               g_index_base, normalize, max_rank, timers, slab_rows, query_groups)
        dev = torch.device("cuda", torch.cuda.current_device())
        ev = eng.event if timers is not None else None
        mark = (lambda: ev()) if ev else (lambda: None)
        pinned = lambda t: isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned() and t.dtype == torch.float32 \
            and t.dim() == 2 and t.stride(1) == 1 and t.shape[1] % 4 == 0 and t.shape[1] <= 2048

        t0 = mark()
        want_groups = query_groups if (self.world > 1 and query_groups > 1 and q_host.shape[0] >= 2 * 1024) else None
        plan, n_local, sizes = self._plan_and_sizes(q_pid, g_pid_local, q_cam, g_cam_local, groups=want_groups,
                                                    local_ok=pinned(g_host_local))
        Q = plan.Q
        max_all = sizes["max_all"] if sizes else plan.max_cnt
        t1 = mark()
        main = torch.cuda.current_stream()
        d = q_host.shape[1]
        w = eng.workspace(plan, d, max_all) if n_local > 0 else None
        # -- queries
        q_dev = None
        if self.world > 1 and shard_query_upload and isinstance(self.coll, LibCollectives) and not q_host.is_cuda:
            per = -(-Q // self.world)
            lo, hi = min(Q, self.rank * per), min(Q, (self.rank + 1) * per)
            part = torch.zeros((per, d), dtype=torch.float32, device=dev)
            part[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
            q_dev = self.coll.all_gather(part).view(self.world * per, d)[:Q]
        if n_local > 0:
            if q_dev is not None:
                eng.prepare(plan, w, q_dev, 0, 0, Q, normalize)
            elif pinned(q_host):
                eng.prepare(plan, w, q_host, 0, 0, Q, normalize, host_input=True)
            else:
                from .metrics import _features
                eng.prepare(plan, w, _features(q_host), 0, 0, Q, normalize)
        if sizes and sizes.get("groups") and sizes["groups"]["ok"]:
            return self._host_grouped_ranks(plan, sizes, w, g_host_local, g_index_base, normalize, max_rank, timers,
                                            slab_rows, (t0, t1))
        # -- gallery slab 0: the queried rows
        g_src, g_host_in = g_host_local, pinned(g_host_local)
        if n_local > 0 and not g_host_in:
            from .metrics import _features
            g_src = _features(g_host_local)
        G = plan.G
        p0 = min(G, -(-max(plan.n_queried, 1) // 256) * 256) if n_local > 0 else 0
        bounds = [0, p0]
        step = self._slab_rows(slab_rows, G - p0)
        while bounds[-1] < G and n_local > 0:
            bounds.append(min(G, bounds[-1] + step))
        up_events = []
        if n_local > 0:
            eng.prepare(plan, w, g_src, 1, 0, p0, normalize, host_input=g_host_in)
            first_in = torch.cuda.Event()
            first_in.record(main)
            side = self._side_stream()
            side.wait_event(first_in)
            with torch.cuda.stream(side):
                for a, b in zip(bounds[1:-1], bounds[2:]):
                    eng.prepare(plan, w, g_src, 1, a, b - a, normalize, host_input=g_host_in)
                    e = torch.cuda.Event()
                    e.record(side)
                    up_events.append(e)
            recs = eng.extract(plan, w, g_index_base)
        else:
            recs = torch.zeros((3, 0), dtype=torch.int32, device=dev)
        t2 = mark()
        thr_ofs, merged, T, max_cnt = self._exchange(plan, recs, sizes)
        t3 = mark()
        thr = eng.thresholds(thr_ofs, merged, Q)
        counts = torch.zeros(max(T, 1), dtype=torch.int32, device=dev)
        t4 = mark()
        n_count = 0
        if n_local > 0:
            for i, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
                if i > 0:
                    main.wait_event(up_events[i - 1])
                if T > 0 and b > a:
                    # reserve_sms > 0: while later slabs are still being pulled in, the persistent GEMM
                    # grid leaves that many SMs to the kernel that pulls them.  Measured on 20k x 1M: not
                    # needed once the pulling grid is small (64 blocks): 189.5 ms with 0, 191.9 with 8.
                    busy = g_host_in and i + 2 < len(bounds)
                    eng.count(w, plan, thr_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=a, g_nrows=b - a,
                              reserve_sms=reserve_sms if busy else 0)
                    n_count += 1
        t5 = mark()
        if self.world > 1:
            self.coll.all_reduce_sum(counts)
        t6 = mark()
        res = self._finish(plan, thr, counts, thr_ofs, sizes["G_total"] if sizes else n_local, max_rank)
        t7 = mark()
        if timers is not None and ev:
            timers.update({"plan": (t0, t1), "upload_queried+records": (t1, t2), "exchange": (t2, t3),
                           "thresholds": (t3, t4), "count(+upload of the other slabs)": (t4, t5),
                           "allreduce": (t5, t6), "finalize": (t6, t7)})
            timers["launches"] = eng.launches(max(n_count, 1), (-(-Q // 256) if max_cnt > 63 else 0) * max(n_count, 1)) \
                + max(len(bounds) - 3, 0)
            timers["slabs"] = len(bounds) - 1
            timers["queried_rows"] = plan.n_queried
        return res

    def _evaluate_host_grouped(self, q_host, g_host, q_pid, g_pid, q_cam, g_cam, g_index_base, normalize, max_rank,
                               timers, slab_rows, groups):
        """evaluate_host on one GPU with the ranking started before the queried gallery rows are
        all in: the pid-sorted queries are cut into `groups` blocks; the gallery rows block j asks
        for are a contiguous piece of the sorted gallery, pulled over PCIe in that order by the side
        stream (one continuous transfer: pieces 0 .. groups-1, then the rest in slabs).  As soon
        as piece j has landed the main stream extracts the records of block j, builds its
        thresholds and counts the two new rectangles (blocks 0..j x piece j, block j x pieces
        0..j-1) -- rank counts are additive over any tiling of Q x G -- so the tensor cores start
        after 1 / groups of the first phase instead of after all of it."""
        eng = self.engine
        dev = torch.device("cuda", torch.cuda.current_device())
        ev = eng.event if timers is not None else None
        mark = (lambda: ev()) if ev else (lambda: None)
        t0 = mark()
        plan = eng.plan(q_pid, g_pid, q_cam, g_cam)
        Q, G, d = plan.Q, plan.G, q_host.shape[1]
        qb = self._query_bounds(Q, groups)
        K = len(qb) - 1
        # gallery rows the queries before each boundary ask for: end of the band of the last such query
        last = torch.tensor([b - 1 for b in qb[1:]], dtype=torch.int64, device=plan.g_lo.device)
        ends = plan.g_lo[last] + plan.rec_ofs[last + 1] - plan.rec_ofs[last]
        host = torch.cat([plan.info, ends.to(torch.int32)]).cpu()          # the one host round trip
        plan.finish(host[:4])
        T, max_cnt = plan.T, plan.max_cnt
        p0 = min(G, -(-max(plan.n_queried, 1) // 256) * 256)
        ends_host = host[4:4 + K - 1].tolist()
        stride = 1                               # small pieces are not worth their launches: merge query blocks
        while -(-K // stride) > 1 and plan.n_queried < MIN_PIECE_ROWS * -(-K // stride):
            stride *= 2
        if stride > 1:
            keep = list(range(stride, K, stride))            # interior boundaries that stay
            qb = [0] + [qb[i] for i in keep] + [Q]
            ends_host = [ends_host[i - 1] for i in keep]
            K = len(qb) - 1
        gb = [0] + [min(p0, int(e)) for e in ends_host] + [p0]
        for j in range(1, len(gb)):
            gb[j] = max(gb[j], gb[j - 1])
        bounds = [p0]
        step = self._slab_rows(slab_rows, G - p0)
        while bounds[-1] < G:
            bounds.append(min(G, bounds[-1] + step))
        t1 = mark()
        main = torch.cuda.current_stream()
        w = eng.workspace(plan, d, max_cnt)
        pinned_q = (isinstance(q_host, torch.Tensor) and not q_host.is_cuda and q_host.is_pinned()
                    and q_host.dtype == torch.float32 and q_host.dim() == 2 and q_host.stride(1) == 1)
        if pinned_q:
            eng.prepare(plan, w, q_host, 0, 0, Q, normalize, host_input=True)
        else:
            from .metrics import _features
            eng.prepare(plan, w, _features(q_host), 0, 0, Q, normalize)
        start = torch.cuda.Event()
        start.record(main)
        side = self._side_stream()
        side.wait_event(start)
        piece_in, slab_in = [], []
        with torch.cuda.stream(side):
            for j in range(K):
                eng.prepare(plan, w, g_host, 1, gb[j], gb[j + 1] - gb[j], normalize, host_input=True)
                e = torch.cuda.Event()
                e.record(side)
                piece_in.append(e)
            for a, b in zip(bounds[:-1], bounds[1:]):
                eng.prepare(plan, w, g_host, 1, a, b - a, normalize, host_input=True)
                e = torch.cuda.Event()
                e.record(side)
                slab_in.append(e)
        thr = eng.thresholds_alloc(T, Q, dev)
        counts = torch.zeros(max(T, 1), dtype=torch.int32, device=dev)
        recs = None
        n_count = 0
        for j in range(K):
            main.wait_event(piece_in[j])
            recs = eng.extract(plan, w, g_index_base, q_row0=qb[j], q_nrows=qb[j + 1] - qb[j])
            eng.thresholds_into(plan.rec_ofs, recs, thr, qb[j], qb[j + 1] - qb[j])   # also zeroes thr_cnt when T == 0
            if T > 0:
                if gb[j + 1] > gb[j]:      # every block known so far x the new piece
                    eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=gb[j],
                              g_nrows=gb[j + 1] - gb[j], q_row0=0, q_nrows=qb[j + 1])
                    n_count += 1
                if j > 0 and gb[j] > 0:    # the new block x the pieces that were already there
                    eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=0, g_nrows=gb[j],
                              q_row0=qb[j], q_nrows=qb[j + 1] - qb[j])
                    n_count += 1
        t2 = mark()
        for i, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            main.wait_event(slab_in[i])
            if T > 0 and b > a:
                eng.count(w, plan, plan.rec_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=a, g_nrows=b - a)
                n_count += 1
        t3 = mark()
        res = self._finish(plan, thr, counts, plan.rec_ofs, G, max_rank)
        t4 = mark()
        if timers is not None and ev:
            timers.update({"plan": (t0, t1), "queried pieces: upload + records + thresholds + count": (t1, t2),
                           "count(+upload of the other slabs)": (t2, t3), "finalize": (t3, t4)})
            timers["launches"] = eng.launches(max(n_count, 1), (-(-Q // 256) if max_cnt > 63 else 0) * max(n_count, 1)) \
                + K + len(bounds) - 1 + 2 * (K - 1)
            timers["slabs"] = K + len(bounds) - 1
            timers["queried_rows"] = plan.n_queried
            timers["query_groups"] = K
        return res

    def _host_grouped_ranks(self, plan, sizes, w, g_host, g_index_base, normalize, max_rank, timers, slab_rows, t01):
        """Several ranks, every one with a pinned host shard: the grouped flow of
        _evaluate_host_grouped with one record exchange per query block.  The records of a block are
        a contiguous slice of every rank's local record array (padded to the longest slice, all-
        gathered, merged into the block's piece of the global CSR), so thresholds and the first count
        rectangles of block j exist while the rows of block j + 1 are still coming in."""
        eng = self.engine
        dev = torch.device("cuda", torch.cuda.current_device())
        ev = eng.event if timers is not None else None
        mark = (lambda: ev()) if ev else (lambda: None)
        g = sizes["groups"]
        qb, K = g["qb"], len(g["qb"]) - 1
        Q, G = plan.Q, plan.G
        T, max_cnt = sizes["T_sum"], sizes["max_all"]
        p0 = min(G, -(-max(plan.n_queried, 1) // 256) * 256)
        gb = [0] + [min(p0, int(e)) for e in g["ends"][:K - 1]] + [p0]
        for j in range(1, len(gb)):
            gb[j] = max(gb[j], gb[j - 1])
        bounds = [p0]
        step = self._slab_rows(slab_rows, G - p0)
        while bounds[-1] < G:
            bounds.append(min(G, bounds[-1] + step))
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)                       # the queries are prepared (enqueued by the caller)
        side = self._side_stream()
        side.wait_event(start)
        piece_in, slab_in = [], []
        with torch.cuda.stream(side):
            for j in range(K):
                eng.prepare(plan, w, g_host, 1, gb[j], gb[j + 1] - gb[j], normalize, host_input=True)
                e = torch.cuda.Event()
                e.record(side)
                piece_in.append(e)
            for a, b in zip(bounds[:-1], bounds[1:]):
                eng.prepare(plan, w, g_host, 1, a, b - a, normalize, host_input=True)
                e = torch.cuda.Event()
                e.record(side)
                slab_in.append(e)
        cnt_all = sizes["cnt_all"]
        total = cnt_all.to(torch.int64).sum(0)
        thr_ofs = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
        thr_ofs[1:] = torch.cumsum(total, 0)
        thr_ofs = thr_ofs.to(torch.int32)
        merged = torch.zeros((3, max(T, 1)), dtype=torch.int32, device=dev)
        thr = eng.thresholds_alloc(T, Q, dev)
        counts = torch.zeros(max(T, 1), dtype=torch.int32, device=dev)
        n_count = 0
        for j in range(K):
            main.wait_event(piece_in[j])
            recs = eng.extract(plan, w, g_index_base, q_row0=qb[j], q_nrows=qb[j + 1] - qb[j])
            lo, hi = int(g["rec_lo"][j]), int(g["rec_lo"][j + 1])
            padded = torch.zeros((3, max(int(g["gmax"][j]), 1)), dtype=torch.int32, device=dev)
            padded[:, :hi - lo] = recs[:, lo:hi]
            recs_all = self.coll.all_gather(padded)
            a, b = int(g["gofs"][j]), int(g["gofs"][j + 1])
            if b > a:
                _, piece, _, _ = merge_records(cnt_all[:, qb[j]:qb[j + 1]], recs_all, sizes=(b - a, max_cnt))
                merged[:, a:b] = piece
                eng.thresholds_into(thr_ofs, merged, thr, qb[j], qb[j + 1] - qb[j])
            else:
                thr[0][qb[j]:qb[j + 1]] = 0
            if T > 0:
                if gb[j + 1] > gb[j]:
                    eng.count(w, plan, thr_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=gb[j],
                              g_nrows=gb[j + 1] - gb[j], q_row0=0, q_nrows=qb[j + 1])
                    n_count += 1
                if j > 0 and gb[j] > 0:
                    eng.count(w, plan, thr_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=0, g_nrows=gb[j],
                              q_row0=qb[j], q_nrows=qb[j + 1] - qb[j])
                    n_count += 1
        t2 = mark()
        for i, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            main.wait_event(slab_in[i])
            if T > 0 and b > a:
                eng.count(w, plan, thr_ofs, thr[0], thr[1], thr[2], counts, max_cnt, g_row0=a, g_nrows=b - a)
                n_count += 1
        t3 = mark()
        self.coll.all_reduce_sum(counts)
        t4 = mark()
        res = self._finish(plan, thr, counts, thr_ofs, sizes["G_total"], max_rank)
        t5 = mark()
        if timers is not None and ev:
            timers.update({"plan": t01, "queried pieces: upload + records + exchange + thresholds + count": (t01[1], t2),
                           "count(+upload of the other slabs)": (t2, t3), "allreduce": (t3, t4), "finalize": (t4, t5)})
            timers["launches"] = eng.launches(max(n_count, 1), (-(-Q // 256) if max_cnt > 63 else 0) * max(n_count, 1)) \
                + K + len(bounds) - 1 + 2 * (K - 1)
            timers["slabs"] = K + len(bounds) - 1
            timers["queried_rows"] = plan.n_queried
            timers["query_groups"] = K
        return res

    @staticmethod
    def _slab_rows(slab_rows, rest):
        """Rows per slab of the gallery part that is pulled in behind the queried rows.  A slab is
        counted only once it has landed completely, so with a small shard (8 ranks: 84k rows, where
        the step is bound by the 8 concurrent PCIe streams, ~22 GB/s each) ONE 131 072-row slab left
        the tensor cores waiting for the whole transfer: at least six slabs, 8 192 .. 131 072 rows."""
        if slab_rows is not None:
            return max(1, int(slab_rows))
        return max(8192, min(131072, -(-rest // (6 * 256)) * 256))

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(priority=0)
        return self._side


def evaluate_gallery_chunks(qf, gf, q_pid, g_pid, q_cam, g_cam, n_chunks: int, normalize: bool = False,
                            max_rank: int = 50, engine=None):
    """The sharded algorithm with all "ranks" executed one after another on ONE device: the
    gallery is cut into `n_chunks` contiguous chunks, every chunk contributes its records, the
    merged thresholds are counted against every chunk and the counts simply add up.  Result is
    bit-identical to the unsharded evaluation (rank counts are additive over gallery partitions);
    used to test the multi-GPU data flow on a single GPU and to evaluate galleries chunk-wise."""
    from .metrics import EvalResult
    eng = engine if engine is not None else CudaEngine()
    G = len(g_pid)
    plans, works, recs_l, cnts = [], [], [], []
    for c in range(n_chunks):
        lo, hi = shard_range(G, n_chunks, c)
        plan = eng.plan(q_pid, g_pid[lo:hi], q_cam, g_cam[lo:hi])
        if plan.T is None:
            plan.finish(plan.info.cpu())
        w, recs = eng.records(plan, qf, gf[lo:hi], lo, normalize)
        plans.append(plan)
        works.append(w)
        recs_l.append(recs)
        cnts.append((plan.rec_ofs[1:] - plan.rec_ofs[:-1]).to(torch.int32))
    t_max = max(max(int(r.shape[1]) for r in recs_l), 1)
    padded = torch.zeros((n_chunks, 3, t_max), dtype=torch.int32, device=recs_l[0].device)
    for c, r in enumerate(recs_l):
        padded[c, :, :r.shape[1]] = r
    thr_ofs, merged, T, max_cnt = merge_records(torch.stack(cnts), padded)
    Q = plans[0].Q
    thr_cnt, thr_val, thr_gidx, thr_junk = eng.thresholds(thr_ofs, merged, Q)
    counts = torch.zeros(max(T, 1), dtype=torch.int32, device=merged.device)
    if T > 0:
        for plan, w in zip(plans, works):
            eng.count(w, plan, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt)
    cmc_d, scal_d, ap, first = eng.finalize(thr_ofs, thr_cnt, thr_junk, counts, plans[0].q_perm, Q,
                                            min(max_rank, G))
    scal = scal_d.cpu()
    return EvalResult(cmc_d.cpu().numpy(), np.float64(scal[0].item()), int(scal[1:2].view(torch.int32)[0].item()),
                      ap, first)


# ------------------------------------------------------------------------------------------------
# Row-sharded k-reciprocal re-ranking (second row)
# ------------------------------------------------------------------------------------------------
class CudaRerankEngine:
    """The four re-ranking stages of one rank on its GPU (C ABI demo_rerank_shard_*)."""

    def __init__(self):
        self.lib = _lib.require_device()

    def dims(self, N, k1, k2):
        K, cap, capq = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.demo_rerank_dims(N, k1, k2, C.byref(K), C.byref(cap), C.byref(capq)))
        return K.value, cap.value, capq.value

    def begin(self, feat, Q, k1, k2, row0, nrows, rows_cap, normalize):
        from .metrics import _features, _ws
        self.feat = _features(feat)
        self.N, self.d = self.feat.shape
        self.Q, self.k1, self.k2 = Q, k1, k2
        self.row0, self.nrows, self.rows_cap = row0, nrows, rows_cap
        self.flags = _lib.FLAG_L2NORM if normalize else 0
        self.nbytes = self.lib.demo_rerank_shard_workspace_bytes(self.N, Q, self.d, k1, k2, rows_cap)
        self.ws = _ws(self.nbytes)
        return self.feat.device

    def topk(self, rank_rows):
        check(self.lib.demo_rerank_shard_topk(ptr(self.feat), self.N, self.Q, self.d, self.feat.stride(0), self.flags,
                                              self.k1, self.k2, self.row0, self.nrows, self.rows_cap, ptr(rank_rows),
                                              None, ptr(self.ws), self.nbytes, stream_ptr()))

    def krecip(self, rank_all, v_idx, v_val, v_cnt):
        check(self.lib.demo_rerank_shard_krecip(self.N, self.Q, self.d, self.k1, self.k2, self.row0, self.nrows,
                                                self.rows_cap, ptr(rank_all), ptr(v_idx), ptr(v_val), ptr(v_cnt),
                                                ptr(self.ws), self.nbytes, stream_ptr()))

    def expand(self, rank_all, v_idx, v_val, v_cnt, q_idx, q_val, q_cnt):
        check(self.lib.demo_rerank_shard_expand(self.N, self.k1, self.k2, self.row0, self.nrows, ptr(rank_all),
                                                ptr(v_idx), ptr(v_val), ptr(v_cnt), ptr(q_idx), ptr(q_val),
                                                ptr(q_cnt), stream_ptr()))

    def jaccard(self, f_idx, f_val, f_cnt, lambda_value, out_rows):
        check(self.lib.demo_rerank_shard_jaccard(self.N, self.Q, self.d, self.k1, self.k2, float(lambda_value),
                                                 self.row0, self.nrows, self.rows_cap, ptr(f_idx), ptr(f_val),
                                                 ptr(f_cnt), ptr(out_rows), out_rows.stride(0) if out_rows.numel() else
                                                 self.N - self.Q, ptr(self.ws), self.nbytes, stream_ptr()))


class ShardedReranker:
    """re_ranking(probFea, galFea, k1, k2, lambda) with the rows of the N x N problem split
    contiguously over the ranks (features replicated):

        every rank   distances of its rows against all N rows + their nearest neighbours   [tcgen05 GEMM, top-k]
        all ranks    all-gather the neighbour lists                  (N x K int32)
        every rank   k-reciprocal sets + softmax weights of its rows -> sparse V rows
        all ranks    all-gather the sparse V rows                    (N x cap: int32 index, fp16 weight)
        every rank   local query expansion of its rows
        all ranks    all-gather the expanded rows                    (N x capq)
        every rank   inverted index (whole V, redundantly) + Jaccard / blend of its query rows
        all ranks    all-gather the [Q, G] result rows

    Every stage is row-local given the gathered lists, so the result is bit-identical to the
    single-GPU re_ranking.  ``ranks`` > 1 with world == 1 runs the ranks one after another on one
    device (the same data flow without a process group; used by the single-GPU tests)."""

    def __init__(self, world: int = 1, rank: int = 0, group=None, engine_factory=None):
        self.world, self.rank, self.group = world, rank, group
        self.engine_factory = engine_factory if engine_factory is not None else CudaRerankEngine

    def _all_gather_rows(self, full: torch.Tensor, rpr: int):
        """In-place all-gather of equal row slices: rank r owns rows [r*rpr, (r+1)*rpr)."""
        import torch.distributed as dist
        mine = full[self.rank * rpr:(self.rank + 1) * rpr]
        if dist.get_backend(self.group) == "gloo" and full.is_cuda:   # two test processes sharing one GPU
            host = torch.empty(full.shape, dtype=full.dtype)
            dist.all_gather_into_tensor(host, mine.cpu(), group=self.group)
            full.copy_(host)
            return
        dist.all_gather_into_tensor(full, mine.clone(), group=self.group)

    def re_ranking(self, probFea, galFea, k1: int, k2: int, lambda_value: float, normalize: bool = False,
                   emulate_ranks: int | None = None):
        P = emulate_ranks if emulate_ranks else self.world
        ranks = range(P) if emulate_ranks else [self.rank]
        engines = {r: self.engine_factory() for r in ranks}
        q = probFea if isinstance(probFea, torch.Tensor) else torch.as_tensor(np.asarray(probFea))
        g = galFea if isinstance(galFea, torch.Tensor) else torch.as_tensor(np.asarray(galFea))
        Q, G = q.shape[0], g.shape[0]
        N = Q + G
        feat = torch.cat([q.float(), g.float().to(q.device)], dim=0)
        rpr = -(-N // P)                       # rows per rank (the last ranks may own fewer / none)
        n_pad = rpr * P
        dev = None
        for r in ranks:
            row0 = min(r * rpr, N)
            dev = engines[r].begin(feat, Q, int(k1), int(k2), row0, min(rpr, N - row0), rpr, normalize)
        K, cap, capq = engines[ranks[0]].dims(N, int(k1), int(k2))
        rank_all = torch.zeros((n_pad, K), dtype=torch.int32, device=dev)
        v_idx = torch.zeros((n_pad, cap), dtype=torch.int32, device=dev)
        v_val = torch.zeros((n_pad, cap), dtype=torch.float16, device=dev)
        v_cnt = torch.zeros(n_pad, dtype=torch.int32, device=dev)
        gather = (lambda t: self._all_gather_rows(t, rpr)) if (self.world > 1 and not emulate_ranks) else (lambda t: None)

        for r in ranks:
            engines[r].topk(rank_all[min(r * rpr, N):])
        gather(rank_all)
        for r in ranks:
            engines[r].krecip(rank_all, v_idx, v_val, v_cnt)
        for t in (v_idx, v_val, v_cnt):
            gather(t)
        f_idx, f_val, f_cnt = v_idx, v_val, v_cnt
        if int(k2) != 1:
            q_idx = torch.zeros((n_pad, capq), dtype=torch.int32, device=dev)
            q_val = torch.zeros((n_pad, capq), dtype=torch.float16, device=dev)
            q_cnt = torch.zeros(n_pad, dtype=torch.int32, device=dev)
            for r in ranks:
                engines[r].expand(rank_all, v_idx, v_val, v_cnt, q_idx, q_val, q_cnt)
            for t in (q_idx, q_val, q_cnt):
                gather(t)
            f_idx, f_val, f_cnt = q_idx, q_val, q_cnt
        # result rows: rank r owns the queries [r*rpr, (r+1)*rpr) & [0, Q)
        qs = min(rpr, Q)                       # largest number of query rows one rank can own
        out = torch.empty((Q, G), dtype=torch.float32, device=dev)
        local = {}
        for r in ranks:
            local[r] = torch.zeros((qs, G), dtype=torch.float32, device=dev)
            engines[r].jaccard(f_idx, f_val, f_cnt, lambda_value, local[r])
        if self.world > 1 and not emulate_ranks:
            import torch.distributed as dist
            buf = torch.empty((P * qs, G), dtype=torch.float32, device=dev)
            if dist.get_backend(self.group) == "gloo" and buf.is_cuda:
                host = torch.empty(buf.shape, dtype=buf.dtype)
                dist.all_gather_into_tensor(host, local[self.rank].cpu(), group=self.group)
                buf.copy_(host)
            else:
                dist.all_gather_into_tensor(buf, local[self.rank], group=self.group)
            pieces = {r: buf[r * qs:(r + 1) * qs] for r in range(P)}
        else:
            pieces = local
        for r in range(P):
            lo, hi = min(r * rpr, Q), min((r + 1) * rpr, Q)
            if hi > lo:
                out[lo:hi] = pieces[r][:hi - lo]
        return out


# ------------------------------------------------------------------------------------------------
# Evaluation under DDP: every rank keeps the features IT extracted
# ------------------------------------------------------------------------------------------------
class DistributedR1mAP:
    """R1_mAP_eval for a validation set that was split over the ranks (a DistributedSampler on the
    val loader) instead of being evaluated on rank 0 alone (engine/processor.py:146-148).

    ``update((feat, pid, camid, index))`` takes what each rank extracted plus the dataset index of
    every sample; samples with ``index < num_query`` are queries (the reference's split,
    utils/metrics.py:347-353).  ``compute()`` all-gathers the (few) query features and labels,
    keeps every rank's gallery features where they are -- they become that rank's gallery shard --
    and runs the gallery-sharded rank-count evaluation.

    * A sample that shows up more than once (DistributedSampler pads the tail of the index list by
      repeating its head when len(dataset) % world != 0) is scored once: the copy on the lowest
      rank wins, first occurrence within a rank.
    * Ties are broken by the DATASET order of the gallery (index - num_query is the tie-break key),
      exactly as a single-process evaluation of the concatenated features would.
    * A rank that ends up without gallery items still takes part in the exchange.
    Returns (cmc, mAP), identical on all ranks."""

    def __init__(self, num_query, world: int = 1, rank: int = 0, group=None, max_rank=50, feat_norm=True,
                 engine=None):
        self.num_query, self.max_rank, self.feat_norm = num_query, max_rank, feat_norm
        self.world, self.rank, self.group = world, rank, group
        self.evaluator = ShardedEvaluator(world=world, rank=rank, group=group, engine=engine)
        self.reset()

    def reset(self):
        self.feats, self.pids, self.camids, self.index = [], [], [], []

    def update(self, output):
        feat, pid, camid, index = output
        self.feats.append(feat.detach())
        self.pids.extend(np.asarray(pid).tolist())
        self.camids.extend(np.asarray(camid.cpu() if isinstance(camid, torch.Tensor) else camid).tolist())
        self.index.extend(np.asarray(index.cpu() if isinstance(index, torch.Tensor) else index).tolist())

    def _gather_var(self, t: torch.Tensor):
        """all-gather of tensors whose first dimension differs per rank (padded to the maximum);
        staged through the host on a gloo group."""
        import torch.distributed as dist
        dev = t.device
        if dist.get_backend(self.group) == "gloo" and t.is_cuda:
            t = t.cpu()
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        pad = torch.zeros((max(max(sizes), 1),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0).to(dev), sizes

    def compute(self):
        feats = torch.cat(self.feats, dim=0).float()
        dev = feats.device
        pids = torch.as_tensor(self.pids, dtype=torch.int64, device=dev)
        cams = torch.as_tensor(self.camids, dtype=torch.int64, device=dev)
        idx = torch.as_tensor(self.index, dtype=torch.int64, device=dev)
        # first occurrence of every dataset index on this rank, in dataset order
        idx_s, order = torch.sort(idx, stable=True)
        first = torch.ones_like(idx_s, dtype=torch.bool)
        first[1:] = idx_s[1:] != idx_s[:-1]
        sel = order[first]
        idx, pids, cams, feats = idx[sel], pids[sel], cams[sel], feats[sel]
        if self.world > 1:
            # the copy on the lowest rank wins
            all_idx, sizes = self._gather_var(idx)
            below = all_idx[:sum(sizes[:self.rank])]
            mine = ~torch.isin(idx, below) if below.numel() else torch.ones_like(idx, dtype=torch.bool)
            idx, pids, cams, feats = idx[mine], pids[mine], cams[mine], feats[mine]
        is_q = idx < self.num_query
        qf, q_lab = feats[is_q], torch.stack([idx[is_q], pids[is_q], cams[is_q]], dim=1)
        gf, g_pid, g_cam = feats[~is_q], pids[~is_q], cams[~is_q]
        g_index = (idx[~is_q] - self.num_query).to(torch.int32)
        if self.world > 1:
            qf, _ = self._gather_var(qf)
            q_lab, _ = self._gather_var(q_lab)
        order = torch.argsort(q_lab[:, 0])                 # dataset order of the queries on every rank
        qf, q_lab = qf[order], q_lab[order]
        res = self.evaluator.evaluate(qf, gf, q_lab[:, 1].to(torch.int32), g_pid.to(torch.int32),
                                      q_lab[:, 2].to(torch.int32), g_cam.to(torch.int32), g_index_base=0,
                                      normalize=bool(self.feat_norm), max_rank=self.max_rank, g_index=g_index)
        assert res.num_valid > 0, "Error: all query identities do not appear in gallery"
        self.last_result = res
        return res.cmc, res.mAP
