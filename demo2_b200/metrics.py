"""Drop-in for the reference's ``utils/metrics.py`` hot path on B200.

Same names, arguments, return types and error behaviour as the reference
(``utils/metrics.py:110-169, 221-248, 341-369, 395-401``); the arithmetic runs in
``libdemo_b200.so`` (hand-written sm_100a CUDA behind a C ABI).  PyTorch is used for
device memory and streams only.

    euclidean_distance(qf, gf)                      -> float32 ndarray [Q, G]
    cosine_similarity(qf, gf)                       -> float32 ndarray [Q, G]   (north_star addition)
    eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50) -> (cmc, mAP)
    R1_mAP_eval(num_query, max_rank=50, feat_norm=True, reranking=False)
        .reset() / .update((feat, pid, camid, img_paths)) / .compute() -> 7-tuple

plus the device-level entry points used by bench.py and the multi-GPU path:
``sqdist_device``, ``evaluate_features``, ``RankPlan``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

# Matrices above this many elements are not materialised by R1_mAP_eval.compute
# (20k x 1M would be 80 GB); callers in the reference ignore the returned distmat
# (engine/processor.py:239, 275).
MAX_MATERIALIZE = 1 << 27


# ------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------
def _dev():
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _features(x) -> torch.Tensor:
    """Any array-like -> contiguous-row fp32 CUDA tensor (H2D copy if needed)."""
    dev = _dev()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x, dtype=np.float32))
    x = x.detach()
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() != 2:
        raise ValueError("features must be 2-D [n, d], got shape %s" % (tuple(x.shape),))
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


def _labels(x) -> torch.Tensor:
    """ids -> int32 CUDA tensor (the reference carries int64 numpy arrays; values are small)."""
    dev = _dev()
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.dtype != torch.int32:
            if t.numel() and (int(t.max()) > 2 ** 31 - 1 or int(t.min()) < -2 ** 31):
                raise ValueError("ids do not fit int32")
            t = t.to(torch.int32)
        return t.to(dev, non_blocking=True).contiguous()
    a = np.asarray(x)
    if a.dtype.kind not in "iu":
        a = a.astype(np.int64)
    if a.size and (a.max() > 2 ** 31 - 1 or a.min() < -2 ** 31):
        raise ValueError("ids do not fit int32")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev, non_blocking=True)


def _ws(nbytes: int) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=_dev())


def to_numpy(t: torch.Tensor) -> np.ndarray:
    """Device matrix -> numpy (what the reference's functions return) through PINNED host memory:
    a pageable ``.cpu()`` of the 58.8 MB RGBNT100 matrix takes an order of magnitude longer than
    the kernels that produced it.  The array owns its pinned block (torch's caching host allocator
    recycles it once the array is dropped)."""
    if not t.is_cuda:
        return t.numpy()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


# ------------------------------------------------------------------------------
# distance matrix
# ------------------------------------------------------------------------------
# Longest rows the tensor-core distance is validated for (csrc/prep.cuh kMaxCompensatedDim): beyond
# it demo_sqdist_f32 uses the fp32 FMA kernel and the fused evaluation the materialised matrix.
MAX_TENSOR_DIM = 2048


def sqdist_device(qf, gf, mode: int = _lib.DIST_SQ, normalize: bool = False, simt: bool = False,
                  want_rowmax: bool = False, want_normalized: bool = False, out: torch.Tensor | None = None):
    """Distance matrix on the device.  Returns ``out`` ([Q, G] fp32 CUDA tensor) or a tuple
    ``(out, rowmax | None, qn | None, gn | None)`` when extras are requested."""
    lib = _lib.require_device()
    q, g = _features(qf), _features(gf)
    Q, d = q.shape
    G, d2 = g.shape
    if d != d2:
        raise ValueError("feature dims differ: %d vs %d" % (d, d2))
    flags = mode | (_lib.FLAG_L2NORM if normalize else 0) | (_lib.FLAG_SIMT if simt else 0)
    if out is None:
        out = torch.empty((Q, G), dtype=torch.float32, device=q.device)
    rowmax = torch.empty(Q, dtype=torch.float32, device=q.device) if want_rowmax else None
    qn = torch.empty((Q, d), dtype=torch.float32, device=q.device) if want_normalized else None
    gn = torch.empty((G, d), dtype=torch.float32, device=q.device) if want_normalized else None
    if Q and G:
        nbytes = lib.demo_sqdist_workspace_bytes(Q, G, d, flags)
        ws = _ws(nbytes)
        check(lib.demo_sqdist_f32(ptr(q), ptr(g), Q, G, d, q.stride(0), g.stride(0), ptr(out), out.stride(0),
                                  flags, ptr(rowmax), ptr(qn), ptr(gn), ptr(ws), nbytes, stream_ptr()))
    if want_rowmax or want_normalized:
        return out, rowmax, qn, gn
    return out


def euclidean_distance(qf, gf):
    """utils/metrics.py:395-401 -- squared L2 (no clamp, no sqrt), returned as numpy."""
    return to_numpy(sqdist_device(qf, gf, _lib.DIST_SQ))


def cosine_similarity(qf, gf):
    """q g^T / (|q| |g|^T) as numpy.  Not in the reference's utils/metrics.py;
    defined as 1 - 2 * cosine_dist (layers/triplet_loss.py:34-48)."""
    return to_numpy(sqdist_device(qf, gf, _lib.DIST_COS_SIM))


# ------------------------------------------------------------------------------
# rank-count evaluation
# ------------------------------------------------------------------------------
class RankPlan:
    """Label-only part of an evaluation (device resident, reusable across evaluations with the
    same ids): pid-sorted permutations (gallery rows whose pid some query asks for first), the
    record CSR and the extract work list.

    ``defer=True`` only ENQUEUES the plan: ``info`` (device int32[4] = T, max same-pid count, band
    units, #queried gallery rows) can then be read together with other data in one host round
    trip and handed to ``finish``."""

    def __init__(self, q_pids, g_pids, q_camids, g_camids, defer: bool = False):
        lib = _lib.require_device()
        self.q_pid, self.g_pid = _labels(q_pids), _labels(g_pids)
        self.q_cam, self.g_cam = _labels(q_camids), _labels(g_camids)
        self.Q, self.G = int(self.q_pid.numel()), int(self.g_pid.numel())
        if self.q_cam.numel() != self.Q or self.g_cam.numel() != self.G:
            raise ValueError("pid / camid length mismatch")
        self.nbytes = lib.demo_plan_bytes(self.Q, self.G)
        self.buf = _ws(self.nbytes)
        check(lib.demo_eval_plan(ptr(self.q_pid), ptr(self.g_pid), self.Q, self.G, ptr(self.buf), self.nbytes,
                                 None, stream_ptr()))
        ptrs = [C.c_void_p() for _ in range(5)]
        check(lib.demo_plan_pointers(ptr(self.buf), self.nbytes, self.Q, self.G, *[C.byref(p) for p in ptrs[:4]]))
        check(lib.demo_plan_info(ptr(self.buf), self.nbytes, self.Q, self.G, C.byref(ptrs[4])))
        base = self.buf.data_ptr()

        def view(p, n):
            off = p.value - base
            return self.buf[off:off + 4 * n].view(torch.int32)

        self.q_perm = view(ptrs[0], self.Q)
        self.g_perm = view(ptrs[1], self.G)
        self.rec_ofs = view(ptrs[2], self.Q + 1)
        self.g_lo = view(ptrs[3], self.Q)
        self.info = view(ptrs[4], 4)
        self.T = self.max_cnt = self.band_units = self.n_queried = None
        if not defer:
            self.finish(self.info.cpu())

    def finish(self, info_host):
        self.T, self.max_cnt, self.band_units, self.n_queried = (int(v) for v in info_host[:4])
        return self


@dataclass
class EvalResult:
    cmc: np.ndarray          # float32 [max_rank]
    mAP: np.float64
    num_valid: int
    ap: torch.Tensor         # float64 [Q] on device, -1 for skipped queries
    first: torch.Tensor      # int32 [Q] on device, rank of first correct match, 0 = skipped
    qn: torch.Tensor | None = None
    gn: torch.Tensor | None = None
    detail: dict | None = None  # device views behind positive_ranks(): thr_ofs, thr_cnt, thr_gidx, thr_junk, counts, q_perm

    def positive_ranks(self):
        """Per valid positive, in canonical order (query ascending, gallery index ascending):
        ``(pos_ofs int64[Q+1], gidx int64[T'], r int64[T'], c int64[T'])`` with r_p = 1 + #{valid
        gallery items before p} and c_p = 1 + #{positives before p}.  Host
        arrays; this is the index-level result the rank-count kernels produce (tests compare it
        with the reference's argsort positions)."""
        return positive_ranks_from_detail(self.detail)


def positive_ranks_from_detail(detail):
    if detail is None:
        raise ValueError("this result carries no per-positive detail")
    ofs = detail["thr_ofs"].cpu().numpy().astype(np.int64)
    cnt = detail["thr_cnt"].cpu().numpy().astype(np.int64)
    q_perm = detail["q_perm"].cpu().numpy().astype(np.int64)
    gidx = detail["thr_gidx"].cpu().numpy().astype(np.int64)
    junk = detail["thr_junk"].cpu().numpy().astype(np.int64)
    counts = detail["counts"].cpu().numpy().astype(np.int64)
    Q = len(cnt)
    per_q = np.zeros(Q, np.int64)
    per_q[q_perm] = cnt
    pos_ofs = np.concatenate([[0], np.cumsum(per_q)])
    T = int(pos_ofs[-1])
    out_g, out_r, out_c = np.zeros(T, np.int64), np.zeros(T, np.int64), np.zeros(T, np.int64)
    for i in range(Q):                     # i = position in pid-sorted query order
        n = int(cnt[i])
        if n == 0:
            continue
        s0, q = int(ofs[i]), int(q_perm[i])
        g = gidx[s0:s0 + n]
        r = 1 + counts[s0:s0 + n] - junk[s0:s0 + n]
        order = np.argsort(g, kind="stable")
        d0 = int(pos_ofs[q])
        out_g[d0:d0 + n], out_r[d0:d0 + n], out_c[d0:d0 + n] = g[order], r[order], (np.arange(n) + 1)[order]
    return pos_ofs, out_g, out_r, out_c


class _EvalWorkspace:
    """Caller-owned workspace + typed views of its result slots."""

    def __init__(self, Q, G, d, T, matrix: bool, max_cnt: int = 0):
        lib = _lib.load()
        self.Q, self.G, self.d, self.T = Q, G, (8 if matrix else d), T
        # max_cnt > 63: room for the distance slab of the query blocks with long threshold lists
        self.nbytes = lib.demo_eval_workspace_bytes_ex(Q, G, self.d, T, 0 if matrix else int(max_cnt))
        self.buf = _ws(self.nbytes)
        ptrs = [C.c_void_p() for _ in range(13)]
        check(lib.demo_eval_ws_pointers(ptr(self.buf), self.nbytes, Q, G, self.d, T, *[C.byref(p) for p in ptrs]))
        base = self.buf.data_ptr()
        names = ["cmc", "map", "nvalid", "ap", "first", "counts", "thr_cnt", "thr_val", "thr_gidx", "thr_junk",
                 "rec_dist", "rec_gidx", "rec_junk"]
        self.off = {n: p.value - base for n, p in zip(names, ptrs)}

    def view(self, name, dtype, count):
        o = self.off[name]
        return self.buf[o:o + count * torch.empty((), dtype=dtype).element_size()].view(dtype)

    def detail(self, plan):
        n = max(self.T, 1)
        return {"thr_ofs": plan.rec_ofs, "thr_cnt": self.view("thr_cnt", torch.int32, self.Q),
                "thr_gidx": self.view("thr_gidx", torch.int32, n), "thr_junk": self.view("thr_junk", torch.int32, n),
                "counts": self.view("counts", torch.int32, n), "q_perm": plan.q_perm}

    def read_metrics(self, max_rank):
        # cmc[4096] | map | nvalid live at the tail of the workspace: one D2H copy
        lo, hi = self.off["cmc"], self.off["nvalid"] + 16
        host = self.buf[lo:hi].cpu().numpy()
        cmc = host[:4 * max_rank].view(np.float32).copy()
        mAP = host[self.off["map"] - lo:self.off["map"] - lo + 8].view(np.float64)[0]
        nvalid = int(host[self.off["nvalid"] - lo:self.off["nvalid"] - lo + 4].view(np.int32)[0])
        return cmc, np.float64(mAP), nvalid


def _effective_max_rank(max_rank: int, num_g: int) -> int:
    if num_g < max_rank:  # utils/metrics.py:118-120
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
        return num_g
    return max_rank


def evaluate_features(qf, gf, q_pids=None, g_pids=None, q_camids=None, g_camids=None, max_rank: int = 50,
                      normalize: bool = False, plan: RankPlan | None = None,
                      want_normalized: bool = False) -> EvalResult:
    """Distance + ranking + CMC/mAP from features without materialising Q x G (fused tcgen05
    GEMM epilogues).  Equivalent to eval_func(euclidean_distance(qf, gf), ...)."""
    lib = _lib.require_device()
    q, g = _features(qf), _features(gf)
    if plan is None:
        plan = RankPlan(q_pids, g_pids, q_camids, g_camids)
    Q, d = q.shape
    G = g.shape[0]
    if (Q, G) != (plan.Q, plan.G) or g.shape[1] != d:
        raise ValueError("feature / label shapes disagree")
    if d > MAX_TENSOR_DIM:
        # outside the validated range of the tensor-core distance: fp32 FMA matrix + streaming count
        if Q * G > 2 ** 32:
            raise ValueError("feature dim %d > %d needs the materialised matrix path; %d x %d is too large for it"
                             % (d, MAX_TENSOR_DIM, Q, G))
        dist, _, qn, gn = sqdist_device(q, g, normalize=normalize, want_normalized=True)
        res = evaluate_matrix(dist, plan=plan, max_rank=max_rank)
        if want_normalized:
            res.qn, res.gn = qn, gn
        return res
    max_rank = _effective_max_rank(max_rank, G)
    w = _EvalWorkspace(Q, G, d, plan.T, matrix=False, max_cnt=plan.max_cnt)
    qn = torch.empty((Q, d), dtype=torch.float32, device=q.device) if want_normalized else None
    gn = torch.empty((G, d), dtype=torch.float32, device=q.device) if want_normalized else None
    flags = _lib.FLAG_L2NORM if normalize else 0
    check(lib.demo_eval_features(ptr(q), ptr(g), Q, G, d, q.stride(0), g.stride(0), flags, ptr(plan.q_cam),
                                 ptr(plan.g_cam), ptr(plan.buf), plan.nbytes, plan.T, plan.max_cnt, max_rank,
                                 ptr(w.buf), w.nbytes, None, None, None, None, None, ptr(qn), ptr(gn),
                                 stream_ptr()))
    cmc, mAP, nvalid = w.read_metrics(max_rank)
    return EvalResult(cmc, mAP, nvalid, w.view("ap", torch.float64, Q), w.view("first", torch.int32, Q), qn, gn,
                      detail=w.detail(plan))


def evaluate_matrix(distmat, q_pids=None, g_pids=None, q_camids=None, g_camids=None, max_rank: int = 50,
                    plan: RankPlan | None = None) -> EvalResult:
    """eval_func on a (device or host) distance matrix: one streaming pass over the matrix."""
    lib = _lib.require_device()
    dm = _features(distmat)
    if plan is None:
        plan = RankPlan(q_pids, g_pids, q_camids, g_camids)
    Q, G = dm.shape
    if (Q, G) != (plan.Q, plan.G):
        raise ValueError("distmat shape %s does not match the labels (%d, %d)" % ((Q, G), plan.Q, plan.G))
    max_rank = _effective_max_rank(max_rank, G)
    w = _EvalWorkspace(Q, G, 8, plan.T, matrix=True)
    check(lib.demo_eval_matrix(ptr(dm), Q, G, dm.stride(0), ptr(plan.q_cam), ptr(plan.g_cam), ptr(plan.buf),
                               plan.nbytes, plan.T, plan.max_cnt, max_rank, ptr(w.buf), w.nbytes, None, None,
                               None, None, None, stream_ptr()))
    cmc, mAP, nvalid = w.read_metrics(max_rank)
    return EvalResult(cmc, mAP, nvalid, w.view("ap", torch.float64, Q), w.view("first", torch.int32, Q),
                      detail=w.detail(plan))


def evaluate_auto(qf, gf, q_pids=None, g_pids=None, q_camids=None, g_camids=None, max_rank: int = 50,
                  normalize: bool = False, plan: RankPlan | None = None) -> EvalResult:
    """What R1_mAP_eval.compute does without re-ranking: materialise the matrix when it is small
    (one GEMM + one streaming count pass; best when queries have hundreds of positives), fused
    rank-count GEMM epilogue otherwise (Q x G never written)."""
    q, g = _features(qf), _features(gf)
    if plan is None:
        plan = RankPlan(q_pids, g_pids, q_camids, g_camids)
    if q.shape[0] * g.shape[0] <= MAX_MATERIALIZE:
        return evaluate_matrix(sqdist_device(q, g, _lib.DIST_SQ, normalize=normalize), plan=plan, max_rank=max_rank)
    return evaluate_features(q, g, plan=plan, normalize=normalize, max_rank=max_rank)


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50):
    """utils/metrics.py:110-169 -- Market-1501 CMC / mAP.  Ties are broken by ascending gallery
    index (the reference's np.argsort leaves tie order unspecified)."""
    r = evaluate_matrix(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
    assert r.num_valid > 0, "Error: all query identities do not appear in gallery"  # :163
    return r.cmc, r.mAP


# Name of the rank-list file eval_func_msrv writes into the working directory, as the reference
# does unconditionally (utils/metrics.py:38-39, 70-77).  Set to None to skip the file.
RANK_LIST_FILE = "re.txt"


def filtered_rank_lists(distmat, remove_counts, max_rank: int):
    """Device top-k feeding the rank-list file / ranked-result visualisation: for every query the
    first ``max_rank + max(remove_counts)`` gallery indices by (distance, index) -- enough to
    still hold ``max_rank`` items after the caller drops the discarded ones.  Returns an int32
    host array [Q, k] (np.argsort(distmat, axis=1)[:, :k] at utils/metrics.py:21 / :279)."""
    from .reranking import topk_rows
    dist = distmat if isinstance(distmat, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(distmat, np.float32))
    dist = dist.to(_dev())
    G = dist.shape[1]
    k = int(min(G, max_rank + (int(np.max(remove_counts)) if len(remove_counts) else 0)))
    if k <= 256:
        return topk_rows(dist, k).cpu().numpy()
    # more discarded items per query than the top-k kernel holds: full device sort of the rows
    # (stable, same (distance, index) order); only reached with > 200 same-identity items
    return torch.sort(dist, dim=1, stable=True).indices[:, :k].to(torch.int32).cpu().numpy()


def eval_func_msrv(distmat, q_pids, g_pids, q_camids, g_camids, q_sceneids, g_sceneids, max_rank=50):
    """utils/metrics.py:12-107 -- MSVR310 protocol: gallery items with the query's pid AND scene
    id are discarded (:67); cameras only appear in the rank-list file.  CMC / mAP come from the
    same rank-count kernels as eval_func with the scene ids in the role of the camera ids; the
    rank-list file (re.txt) is written from a device top-k."""
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    q_sceneids, g_sceneids = np.asarray(q_sceneids), np.asarray(g_sceneids)
    dist = distmat if isinstance(distmat, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(distmat, np.float32))
    dist = dist.to(_dev())
    num_g = dist.shape[1]
    r = evaluate_matrix(dist, q_pids, g_pids, q_sceneids, g_sceneids, max_rank)
    if num_g < max_rank:
        max_rank = num_g
    if RANK_LIST_FILE:
        # discarded items per query = gallery items with the same (pid, scene)
        key_g = g_pids.astype(np.int64) * (1 << 31) + g_sceneids.astype(np.int64)
        key_q = q_pids.astype(np.int64) * (1 << 31) + q_sceneids.astype(np.int64)
        uniq, cnt = np.unique(key_g, return_counts=True)
        pos = np.searchsorted(uniq, key_q)
        pos_c = np.minimum(pos, len(uniq) - 1)
        removed = np.where(uniq[pos_c] == key_q, cnt[pos_c], 0)
        top = filtered_rank_lists(dist, removed, max_rank)
        with open(RANK_LIST_FILE, "w") as f:
            f.write("rank list file\n")
            for qi in range(len(q_pids)):
                order = top[qi]
                keep = ~((g_pids[order] == q_pids[qi]) & (g_sceneids[order] == q_sceneids[qi]))
                sel = order[keep][:max_rank]
                f.write("{}_s{}_v{}:\n".format(q_pids[qi], q_sceneids[qi], q_camids[qi]))
                f.write("".join("{}_s{}_v{}  ".format(a, c, b) for a, b, c in
                                zip(g_pids[sel], g_camids[sel], g_sceneids[sel])) + "\n")
    assert r.num_valid > 0, "Error: all query identities do not appear in gallery"  # :101
    return r.cmc, r.mAP


# ------------------------------------------------------------------------------
# evaluators
# ------------------------------------------------------------------------------
class R1_mAP():
    """utils/metrics.py:172-218 -- the MSVR310 evaluator.  ``update`` takes the 5-tuple
    (feat, pid, camid, sceneid, img_path); features are normalised iff ``feat_norm == 'yes'``
    (:196); ``compute`` returns (cmc, mAP, distmat, pids, camids, qf, gf)."""

    def __init__(self, num_query, max_rank=50, feat_norm='yes'):
        super(R1_mAP, self).__init__()
        self.num_query = num_query
        self.max_rank = max_rank
        self.feat_norm = feat_norm

    def reset(self):
        self.feats = []
        self.pids = []
        self.camids = []
        self.sceneids = []
        self.img_path = []

    def update(self, output):
        feat, pid, camid, sceneid, img_path = output
        self.feats.append(feat.detach())
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid.cpu() if isinstance(camid, torch.Tensor) else camid))
        self.sceneids.extend(np.asarray(sceneid.cpu() if isinstance(sceneid, torch.Tensor) else sceneid))
        self.img_path.extend(img_path)

    def compute(self):
        dev = _dev()
        feats = torch.cat([f.to(dev, non_blocking=True) for f in self.feats], dim=0).float()
        norm = self.feat_norm == 'yes'
        if norm:
            print("The test feature is normalized")
        nq = self.num_query
        dist_dev, _, qf, gf = sqdist_device(feats[:nq], feats[nq:], _lib.DIST_SQ, normalize=norm, want_normalized=True)
        if not norm:
            qf, gf = feats[:nq], feats[nq:]
        cmc, mAP = eval_func_msrv(dist_dev, np.asarray(self.pids[:nq]), np.asarray(self.pids[nq:]),
                                  np.asarray(self.camids[:nq]), np.asarray(self.camids[nq:]),
                                  np.asarray(self.sceneids[:nq]), np.asarray(self.sceneids[nq:]))
        return cmc, mAP, to_numpy(dist_dev), self.pids, self.camids, qf, gf


class R1_mAP_eval():
    """utils/metrics.py:221-369.  Quirks kept: ``feat_norm`` is truthiness-tested (:343),
    ``compute`` ignores ``self.max_rank`` (:364) and re-ranks with k1=50, k2=15, lambda=0.3 (:359).
    Difference: ``update`` leaves CUDA features on the device (the reference's ``feat.cpu()``
    at :244 is a per-batch sync)."""

    def __init__(self, num_query, max_rank=50, feat_norm=True, reranking=False):
        super(R1_mAP_eval, self).__init__()
        self.num_query = num_query
        self.max_rank = max_rank
        self.feat_norm = feat_norm
        self.reranking = reranking
        self.reset()

    def reset(self):
        self.feats = []
        self.pids = []
        self.camids = []
        self.img_paths = []
        self.img_prefixes = {
            'RGB': '../RGBNT201/test/RGB/',
            'NIR': '../RGBNT201/test/NI/',
            'TIR': '../RGBNT201/test/TI/'
        }

    def update(self, output):  # called once for each batch
        feat, pid, camid, img_paths = output
        self.feats.append(feat.detach())
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid.cpu() if isinstance(camid, torch.Tensor) else camid))
        self.img_paths.extend(img_paths)

    def set_image_prefixes(self, rgb_prefix, nir_prefix, tir_prefix):
        self.img_prefixes['RGB'] = rgb_prefix
        self.img_prefixes['NIR'] = nir_prefix
        self.img_prefixes['TIR'] = tir_prefix

    def ranked_results(self, distmat, topk=10, num2vis=100):
        """The selection step of visualize_ranked_results (utils/metrics.py:273-297) without the
        plotting: for each of the first ``num2vis`` queries the ``topk`` nearest gallery indices
        among the gallery items seen by a DIFFERENT camera (:279-280) and their pids (:297).  The
        same-camera columns are masked on the device and the row top-k kernel replaces the full
        np.argsort; ties go to the lower gallery index."""
        from .reranking import topk_rows
        nq = self.num_query
        n = min(num2vis, nq)
        dist = distmat if isinstance(distmat, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(distmat, np.float32))
        dist = dist[:n].to(_dev()).float()
        cams = torch.as_tensor(np.asarray(self.camids, dtype=np.int64), device=dist.device)
        same = cams[nq:].unsqueeze(0) == cams[:n].unsqueeze(1)
        masked = torch.where(same, torch.full_like(dist, float("inf")), dist)
        k = min(int(topk), dist.shape[1])
        idx = topk_rows(masked, k).cpu().numpy()
        n_ok = (~same).sum(1).cpu().numpy()          # fewer than topk different-camera items: shorter list
        pids = np.asarray(self.pids)
        lists = [idx[i, :min(k, int(n_ok[i]))].tolist() for i in range(n)]
        return lists, [[pids[j + nq] for j in row] for row in lists]

    def compute(self):  # called after each epoch
        dev = _dev()
        feats = torch.cat([f.to(dev, non_blocking=True) for f in self.feats], dim=0).float()
        norm = bool(self.feat_norm)
        if norm:
            print("The test feature is normalized")
        nq = self.num_query
        q_pids = np.asarray(self.pids[:nq])
        q_camids = np.asarray(self.camids[:nq])
        g_pids = np.asarray(self.pids[nq:])
        g_camids = np.asarray(self.camids[nq:])
        qraw, graw = feats[:nq], feats[nq:]
        Q, G = qraw.shape[0], graw.shape[0]
        plan = RankPlan(q_pids, g_pids, q_camids, g_camids)

        if self.reranking:
            print('=> Enter reranking')
            from .reranking import re_ranking_device
            dist_dev, qf, gf = re_ranking_device(qraw, graw, k1=50, k2=15, lambda_value=0.3, normalize=norm,
                                                 want_normalized=True)
            res = evaluate_matrix(dist_dev, plan=plan)
            distmat = to_numpy(dist_dev)
        else:
            print('=> Computing DistMat with euclidean_distance')
            if Q * G <= MAX_MATERIALIZE:
                dist_dev, _, qf, gf = sqdist_device(qraw, graw, _lib.DIST_SQ, normalize=norm, want_normalized=True)
                res = evaluate_matrix(dist_dev, plan=plan)
                distmat = to_numpy(dist_dev)
            else:
                res = evaluate_features(qraw, graw, plan=plan, normalize=norm, want_normalized=True)
                qf, gf, distmat = res.qn, res.gn, None
        assert res.num_valid > 0, "Error: all query identities do not appear in gallery"
        self.last_result = res
        return res.cmc, res.mAP, distmat, self.pids, self.camids, qf, gf
