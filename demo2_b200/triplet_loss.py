"""Drop-in for the reference's ``layers/triplet_loss.py`` on B200.

    normalize(x, axis=-1)                                         (:5-13)
    euclidean_dist(x, y) / cosine_dist(x, y)                      (:16-48)   differentiable
    hard_example_mining(dist_mat, labels, return_inds=False)      (:51-104)
    TripletLoss / MultiModalTripletLoss(margin=None, hard_factor=0.0)
        .__call__(global_feat, labels, normalize_feature=False) -> (loss, dist_ap, dist_an)   (:107-167)

``TripletLoss`` takes the fused path: distance GEMM + hard mining in one tcgen05 kernel
(the N x N matrix is never written) and a sparse backward (only the 2N selected pairs carry
gradient).  The free functions keep the reference's signatures for callers that build the
matrix themselves.  All arithmetic is fp32 (the reference runs the matmul in fp16 under
autocast on GPU).
"""
from __future__ import annotations

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


class TripletLoss(object):
    """
    Triplet loss using HARDER example mining (layers/triplet_loss.py:107-135), fused
    distance + mining forward and sparse backward.
    """

    def __init__(self, margin=None, hard_factor=0.0, check_pk=True):
        self.margin = margin
        self.hard_factor = hard_factor
        self.check_pk = check_pk
        if margin is not None:
            self.ranking_loss = nn.MarginRankingLoss(margin=margin)
        else:
            self.ranking_loss = nn.SoftMarginLoss()

    def __call__(self, global_feat, labels, normalize_feature=False):
        if normalize_feature:
            global_feat = normalize(global_feat, axis=-1)
        dist_ap, dist_an, _, _ = fused_hard_mining(global_feat, labels, self.check_pk)

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
