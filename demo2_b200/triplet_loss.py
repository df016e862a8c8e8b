"""Drop-in for the reference's ``layers/triplet_loss.py`` on B200.

    normalize(x, axis=-1)                                         (:5-13)
    euclidean_dist(x, y) / cosine_dist(x, y)                      (:16-48)   differentiable
    hard_example_mining(dist_mat, labels, return_inds=False)      (:51-104)
    TripletLoss / MultiModalTripletLoss(margin=None, hard_factor=0.0)
        .__call__(global_feat, labels, normalize_feature=False) -> (loss, dist_ap, dist_an)   (:107-167)

``TripletLoss`` takes a fused path.  Training-size batches (N <= 256, the reference's PK batches
are 64 or 128) run ONE kernel for the whole forward -- fp32 Gram matrix, sqrt / clamp, hard
mining, ranking loss and its mean, for all modalities at once (``triplet_loss_multi``) -- and one
kernel for the backward; there is no host synchronisation (the PK-batch check of the reference's
``view(N, -1)`` is evaluated on the device and inspected at the next call).  Larger batches use
the tcgen05 distance GEMM with the mining fused into its epilogue (the N x N matrix is never
written) and a sparse backward.  The free functions keep the reference's signatures for callers
that build the matrix themselves.  All arithmetic is fp32 (the reference runs the matmul in fp16
under autocast on GPU).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream_ptr
from .metrics import _ws, sqdist_device


def normalize(x, axis=-1):
    """:5-13  x / (||x||_2 + 1e-12) along `axis` (plain torch: elementwise, differentiable)."""
    x = 1. * x / (torch.norm(x, 2, axis, keepdim=True).expand_as(x) + 1e-12)
    return x


def _as_cuda_f32(x: torch.Tensor) -> torch.Tensor:
    _lib.require_device()
    if not x.is_cuda:
        x = x.cuda()
    return x.float()


def _labels_i32(labels: torch.Tensor, dev) -> torch.Tensor:
    return labels.to(device=dev, dtype=torch.int32).contiguous()


class _EuclideanDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        d = sqdist_device(x, y, _lib.DIST_SQRT)
        ctx.save_for_backward(x, y, d)
        return d

    @staticmethod
    def backward(ctx, g):
        x, y, d = ctx.saved_tensors
        # d = sqrt(clamp(s, 1e-12)):  dd/ds = 1/(2d) where the clamp is inactive, else 0;
        # ds/dx_i = 2 (x_i - y_j)  ->  grad_x = rowsum(W) * x - W y,  W = g / d
        w = torch.where(d * d > 1e-12, g / d, torch.zeros_like(g))
        gx = w.sum(1, keepdim=True) * x - w @ y
        gy = w.sum(0).unsqueeze(1) * y - w.t() @ x
        return gx, gy


def euclidean_dist(x, y):
    """:16-31  sqrt(clamp(|x|^2 + |y|^2^T - 2 x y^T, 1e-12)) -> [m, n] (on the device)."""
    x, y = _as_cuda_f32(x), _as_cuda_f32(y)
    if x.requires_grad or y.requires_grad:
        return _EuclideanDist.apply(x, y)
    return sqdist_device(x, y, _lib.DIST_SQRT)


def cosine_dist(x, y):
    """:34-48  (1 - x y^T / (|x| |y|^T)) / 2.  (Unused by the reference's callers; the backward
    is expressed through the similarity with plain torch ops.)"""
    x, y = _as_cuda_f32(x), _as_cuda_f32(y)
    if x.requires_grad or y.requires_grad:
        xn = x / x.norm(dim=1, keepdim=True)
        yn = y / y.norm(dim=1, keepdim=True)
        return (1. - xn @ yn.t()) / 2
    return sqdist_device(x, y, _lib.DIST_COS_DIST)


def _check_equal_positives(npos: torch.Tensor):
    # the reference's dist_mat[is_pos].view(N, -1) (:79) fails unless every anchor has the same
    # number of positives
    if bool((npos != npos[0]).any()):
        raise RuntimeError("hard_example_mining: anchors have different numbers of positives "
                           "(the reference's view(N, -1) at layers/triplet_loss.py:79 requires a PK batch)")


def hard_example_mining(dist_mat, labels, return_inds=False):
    """:51-104  hardest positive (max, anchor included) and hardest negative (min) per anchor.
    Differentiable with respect to dist_mat (gather of the selected entries)."""
    lib = _lib.require_device()
    assert len(dist_mat.size()) == 2
    assert dist_mat.size(0) == dist_mat.size(1)
    dm = _as_cuda_f32(dist_mat)
    N = dm.size(0)
    dmc = dm.detach()
    if dmc.stride(1) != 1:
        dmc = dmc.contiguous()
    lab = _labels_i32(labels, dm.device)
    ap = torch.empty(N, dtype=torch.float32, device=dm.device)
    an = torch.empty(N, dtype=torch.float32, device=dm.device)
    p_inds = torch.empty(N, dtype=torch.int64, device=dm.device)
    n_inds = torch.empty(N, dtype=torch.int64, device=dm.device)
    npos = torch.empty(N, dtype=torch.int32, device=dm.device)
    check(lib.demo_hard_example_mining(ptr(dmc), N, dmc.stride(0), ptr(lab), ptr(ap), ptr(an), ptr(p_inds),
                                       ptr(n_inds), ptr(npos), stream_ptr()))
    _check_equal_positives(npos)
    if bool((n_inds < 0).any()):
        raise RuntimeError("hard_example_mining: an anchor has no negative sample in the batch")
    if dm.requires_grad:
        ap = dm.gather(1, p_inds.unsqueeze(1)).squeeze(1)
        an = dm.gather(1, n_inds.unsqueeze(1)).squeeze(1)
    if return_inds:
        return ap, an, p_inds, n_inds
    return ap, an


class _FusedHardTriplet(torch.autograd.Function):
    """(dist_ap, dist_an, p_inds, n_inds) = mine(euclidean_dist(x, x), labels) in one kernel."""

    @staticmethod
    def forward(ctx, x, labels, check_pk):
        lib = _lib.require_device()
        x = x.contiguous() if x.stride(1) != 1 else x
        N, d = x.shape
        # up to 12 800 anchors: the library refuses larger batches here, in forward (its backward
        # kernel keeps 16 bytes per anchor in shared memory)
        lab = _labels_i32(labels, x.device)
        ap = torch.empty(N, dtype=torch.float32, device=x.device)
        an = torch.empty(N, dtype=torch.float32, device=x.device)
        p_inds = torch.empty(N, dtype=torch.int64, device=x.device)
        n_inds = torch.empty(N, dtype=torch.int64, device=x.device)
        npos = torch.empty(N, dtype=torch.int32, device=x.device) if check_pk else None
        nbytes = lib.demo_triplet_workspace_bytes(N, d)
        ws = _ws(nbytes)
        check(lib.demo_triplet_hard_fwd(ptr(x), N, d, x.stride(0), ptr(lab), ptr(ap), ptr(an), ptr(p_inds),
                                        ptr(n_inds), ptr(npos), ptr(ws), nbytes, stream_ptr()))
        if check_pk:
            _check_equal_positives(npos)
        ctx.save_for_backward(x, ap, an, p_inds, n_inds)
        ctx.mark_non_differentiable(p_inds, n_inds)
        return ap, an, p_inds, n_inds

    @staticmethod
    def backward(ctx, g_ap, g_an, _gp, _gn):
        lib = _lib.require_device()
        x, ap, an, p_inds, n_inds = ctx.saved_tensors
        N, d = x.shape
        g_ap = torch.zeros_like(ap) if g_ap is None else g_ap.contiguous().float()
        g_an = torch.zeros_like(an) if g_an is None else g_an.contiguous().float()
        grad = torch.empty((N, d), dtype=torch.float32, device=x.device)
        check(lib.demo_triplet_hard_bwd(ptr(x), N, d, x.stride(0), ptr(p_inds), ptr(n_inds), ptr(ap), ptr(an),
                                        ptr(g_ap), ptr(g_an), ptr(grad), grad.stride(0), stream_ptr()))
        return grad, None, None


def fused_hard_mining(global_feat, labels, check_pk: bool = True):
    """euclidean_dist(x, x) + hard_example_mining(..., return_inds=True) without materialising
    the matrix.  Returns (dist_ap, dist_an, p_inds, n_inds); differentiable in global_feat."""
    x = _as_cuda_f32(global_feat)
    return _FusedHardTriplet.apply(x, labels, check_pk)


# ------------------------------------------------------------------------------------------------
# one-launch path for training-size batches
# ------------------------------------------------------------------------------------------------
_SMALL_WS = {}        # device index -> zero-initialised workspace (tickets return to 0 after every launch)
_PENDING = []         # (pinned status copy, event) of launches not inspected yet
_STATUS_RING = {}     # device index -> (device int32 [64, 8], pinned host mirror, next slot)
_MAX_BATCH = None


def _small_limit() -> int:
    global _MAX_BATCH
    if _MAX_BATCH is None:
        _MAX_BATCH = int(_lib.load().demo_triplet_loss_max_batch())
    return _MAX_BATCH


def _status_slot(dev):
    """A (device, pinned host) pair of int32[8] status words from a small ring: no allocation per call."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ring = _STATUS_RING.get(key)
    if ring is None:
        ring = [torch.zeros((64, 8), dtype=torch.int32, device=dev), torch.zeros((64, 8), dtype=torch.int32).pin_memory(), 0]
        _STATUS_RING[key] = ring
    i = ring[2]
    ring[2] = (i + 1) % 64
    if len(_PENDING) >= 60:          # the ring is about to wrap: drain before a slot is reused
        check_pending_status(block=True)
    return ring[0][i], ring[1][i]


def _small_workspace(dev) -> torch.Tensor:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _SMALL_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.load().demo_triplet_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
        _SMALL_WS[key] = ws
    return ws


def _raise_for_status(bits: int, when: str):
    if bits & 1:
        raise RuntimeError("hard_example_mining: anchors have different numbers of positives (the reference's "
                           "view(N, -1) at layers/triplet_loss.py:79 requires a PK batch)" + when)
    if bits & 2:
        raise RuntimeError("hard_example_mining: an anchor has no negative sample in the batch" + when)


def check_pending_status(block: bool = False):
    """Inspects the device-side batch checks of earlier fused launches (unequal numbers of positives,
    anchors without a negative) without stalling the stream: entries whose copy has completed are
    read, the others stay queued unless ``block``."""
    while _PENDING:
        host, ev = _PENDING[0]
        if not block and not ev.query():
            return
        ev.synchronize()
        _PENDING.pop(0)
        _raise_for_status(int(host.max().item()) if host.numel() else 0, " [reported by an earlier TripletLoss call]")


class _TripletLossFused(torch.autograd.Function):
    """(loss[B], dist_ap[B, N], dist_an[B, N]) for B feature matrices sharing the labels: one
    forward kernel, one backward kernel."""

    @staticmethod
    def forward(ctx, labels, margin, hard_factor, check, *xs):
        lib = _lib.require_device()
        xs = tuple(x if (x.stride(1) == 1 and x.stride(0) == xs[0].stride(0)) else x.contiguous() for x in xs)
        if any(x.stride(0) != xs[0].stride(0) for x in xs):
            xs = tuple(x.contiguous() for x in xs)
        B = len(xs)
        N, d = xs[0].shape
        dev = xs[0].device
        lab = labels.to(dev)
        if lab.dtype not in (torch.int32, torch.int64):
            lab = lab.to(torch.int64)
        lab = lab.contiguous()
        fbuf = torch.empty(B * (1 + 2 * N), dtype=torch.float32, device=dev)     # loss | dist_ap | dist_an
        loss, ap, an = fbuf[:B], fbuf[B:B + B * N].view(B, N), fbuf[B + B * N:].view(B, N)
        ibuf = torch.empty((2, B, N), dtype=torch.int64, device=dev)
        p_inds, n_inds = ibuf[0], ibuf[1]
        status, status_host = _status_slot(dev) if check else (None, None)
        ws = _small_workspace(dev)
        ptrs = (C.c_void_p * B)(*[x.data_ptr() for x in xs])
        m = -1.0 if margin is None else float(margin)
        _lib.check(lib.demo_triplet_loss_fwd(ptrs, B, N, d, xs[0].stride(0), ptr(lab), 1 if lab.dtype == torch.int64 else 0,
                                             m, float(hard_factor), ptr(loss), ptr(ap), ptr(an), ptr(p_inds), ptr(n_inds),
                                             ptr(status), ptr(ws), ws.numel(), stream_ptr()))
        if check == "now":
            _raise_for_status(int(status[:B].max().item()), "")
        elif check:
            status_host.copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            _PENDING.append((status_host[:B], ev))
        ctx.save_for_backward(ap, an, p_inds, n_inds, *xs)
        ctx.margin, ctx.hard_factor = m, float(hard_factor)
        ctx.mark_non_differentiable(p_inds, n_inds)
        return loss, ap, an, p_inds, n_inds

    @staticmethod
    def backward(ctx, g_loss, g_ap, g_an, _gp, _gn):
        lib = _lib.require_device()
        ap, an, p_inds, n_inds, *xs = ctx.saved_tensors
        B = len(xs)
        N, d = xs[0].shape
        dev = xs[0].device
        g_loss = None if g_loss is None else g_loss.contiguous().float()
        g_ap = None if g_ap is None else g_ap.contiguous().float()
        g_an = None if g_an is None else g_an.contiguous().float()
        grad = torch.empty((B, N, d), dtype=torch.float32, device=dev)
        ptrs = (C.c_void_p * B)(*[x.data_ptr() for x in xs])
        _lib.check(lib.demo_triplet_loss_bwd(ptrs, B, N, d, xs[0].stride(0), ctx.margin, ctx.hard_factor, ptr(ap), ptr(an),
                                             ptr(p_inds), ptr(n_inds), ptr(g_loss), ptr(g_ap), ptr(g_an), ptr(grad), d,
                                             stream_ptr()))
        return (None, None, None, None) + tuple(grad[b] for b in range(B))


def triplet_loss_multi(feats, labels, margin=None, hard_factor=0.0, check="deferred", return_inds: bool = False):
    """TripletLoss for several same-shaped feature matrices that share the labels -- the per-modality
    calls of layers/make_loss.py:47-52 (RGB / NIR / TIR heads of one batch) -- in ONE kernel.
    Returns (loss [B], dist_ap [B, N], dist_an [B, N]) (+ p_inds, n_inds); row b is exactly what
    ``TripletLoss(margin, hard_factor)(feats[b], labels)`` returns.  N <= 256.

    ``check``: "deferred" (default) evaluates the reference's implicit batch requirements (equal
    number of positives per anchor, a negative for every anchor) on the device and raises at the
    next call, with no synchronisation; "now" synchronises and raises immediately; False skips."""
    xs = [_as_cuda_f32(x) for x in feats]
    N = xs[0].shape[0]
    if N > _small_limit():
        raise ValueError("triplet_loss_multi handles up to %d anchors per batch (got %d)" % (_small_limit(), N))
    if any(tuple(x.shape) != tuple(xs[0].shape) for x in xs):
        raise ValueError("all feature matrices must have the same shape")
    check_pending_status()
    loss, ap, an, p_inds, n_inds = _TripletLossFused.apply(labels, margin, hard_factor, check, *xs)
    if return_inds:
        return loss, ap, an, p_inds, n_inds
    return loss, ap, an


class TripletLoss(object):
    """
    Triplet loss using HARDER example mining (layers/triplet_loss.py:107-135).  Batches of up to
    256 anchors: one fused kernel (distance + mining + loss); larger: tcgen05 distance GEMM with
    the mining in its epilogue.  ``check_pk``: "deferred" (default; no host synchronisation, a
    malformed batch raises at the next call), True (synchronise and raise immediately, as the
    reference's view(N, -1) would), False.
    """

    def __init__(self, margin=None, hard_factor=0.0, check_pk="deferred"):
        self.margin = margin
        self.hard_factor = hard_factor
        self.check_pk = check_pk
        if margin is not None:
            self.ranking_loss = nn.MarginRankingLoss(margin=margin)
        else:
            self.ranking_loss = nn.SoftMarginLoss()

    def forward_multi(self, feats, labels, normalize_feature=False):
        """All modalities of one batch in one launch: (loss [B], dist_ap [B, N], dist_an [B, N])."""
        if normalize_feature:
            feats = [normalize(f, axis=-1) for f in feats]
        check = "now" if self.check_pk is True else self.check_pk
        return triplet_loss_multi(feats, labels, self.margin, self.hard_factor, check)

    def __call__(self, global_feat, labels, normalize_feature=False):
        if normalize_feature:
            global_feat = normalize(global_feat, axis=-1)
        if global_feat.dim() == 2 and global_feat.shape[0] <= _small_limit():
            check = "now" if self.check_pk is True else self.check_pk
            loss, dist_ap, dist_an = triplet_loss_multi([global_feat], labels, self.margin, self.hard_factor, check)
            return loss[0], dist_ap[0], dist_an[0]
        dist_ap, dist_an, _, n_inds = fused_hard_mining(global_feat, labels, bool(self.check_pk))
        if self.check_pk and bool((n_inds < 0).any()):
            raise RuntimeError("hard_example_mining: an anchor has no negative sample in the batch")

        dist_ap = dist_ap * (1.0 + self.hard_factor)
        dist_an = dist_an * (1.0 - self.hard_factor)

        y = torch.ones_like(dist_an)
        if self.margin is not None:
            loss = self.ranking_loss(dist_an, dist_ap, y)
        else:
            loss = self.ranking_loss(dist_an - dist_ap, y)
        return loss, dist_ap, dist_an


class MultiModalTripletLoss(TripletLoss):
    """Verbatim duplicate of TripletLoss in the reference (layers/triplet_loss.py:139-167)."""
