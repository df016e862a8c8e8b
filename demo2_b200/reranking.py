"""Drop-in for the reference's ``utils/reranking.py`` (k-reciprocal re-ranking, Zhong et al.
CVPR'17) on B200.

    re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False)
        -> float32 ndarray [Q, G]                       (utils/reranking.py:29, the fork's form)
    re_ranking(q_g, q_q, g_g, k1=20, k2=6, lambda_value=0.3)
        -> float32 ndarray [Q, G]                       (distance-matrix form named by north_star;
           q_g / q_q / g_g are the SQUARED-distance blocks of one all-pairs matrix)

The float16 semantics of the reference (V, V_qe, temp_min, jaccard_dist) are reproduced on the
device; see csrc/rerank.cu.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .metrics import _dev, _features, _ws, to_numpy


def _is_matrix_like(x) -> bool:
    return isinstance(x, (np.ndarray, torch.Tensor)) and getattr(x, "ndim", 0) == 2


def re_ranking_device(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False,
                      normalize: bool = False, want_normalized: bool = False):
    """Feature form on the device.  Returns the [Q, G] CUDA tensor (and the normalised
    query / gallery features when ``want_normalized``)."""
    lib = _lib.require_device()
    q, g = _features(probFea), _features(galFea)
    Q, d = q.shape
    G = g.shape[0]
    N = Q + G
    feat = torch.cat([q, g], dim=0)
    loc = None
    if local_distmat is not None:
        loc = _features(local_distmat)
        if tuple(loc.shape) != (N, N):
            raise ValueError("local_distmat must be [%d, %d]" % (N, N))
    out = torch.empty((Q, G), dtype=torch.float32, device=feat.device)
    featn = torch.empty((N, d), dtype=torch.float32, device=feat.device) if want_normalized else None
    nbytes = lib.demo_rerank_workspace_bytes(N, Q, d, int(k1), int(k2))
    ws = _ws(nbytes)
    flags = _lib.FLAG_L2NORM if normalize else 0
    check(lib.demo_rerank(ptr(feat), N, Q, d, feat.stride(0), flags, int(k1), int(k2), float(lambda_value),
                          ptr(loc), loc.stride(0) if loc is not None else 0, 1 if only_local else 0, ptr(out),
                          out.stride(0), ptr(featn), ptr(ws), nbytes, stream_ptr()))
    if want_normalized:
        return out, featn[:Q], featn[Q:]
    return out


def re_ranking_matrix_device(q_g, q_q, g_g, k1=20, k2=6, lambda_value=0.3):
    """Distance-matrix form on the device: blocks of the all-pairs squared-distance matrix
    [[q_q, q_g], [q_g^T, g_g]]."""
    lib = _lib.require_device()
    qg, qq, gg = _features(q_g), _features(q_q), _features(g_g)
    Q, G = qg.shape
    if tuple(qq.shape) != (Q, Q) or tuple(gg.shape) != (G, G):
        raise ValueError("expected q_g [Q,G], q_q [Q,Q], g_g [G,G]")
    N = Q + G
    X = torch.cat([torch.cat([qq, qg], dim=1), torch.cat([qg.t(), gg], dim=1)], dim=0).contiguous()
    out = torch.empty((Q, G), dtype=torch.float32, device=X.device)
    nbytes = lib.demo_rerank_workspace_bytes(N, Q, 8, int(k1), int(k2))
    ws = _ws(nbytes)
    check(lib.demo_rerank_matrix(ptr(X), X.stride(0), N, Q, int(k1), int(k2), float(lambda_value), ptr(out),
                                 out.stride(0), ptr(ws), nbytes, stream_ptr()))
    return out


def re_ranking(*args, **kwargs):
    """See module docstring: dispatches on the call form (2 feature matrices vs 3 distance blocks)."""
    if len(args) >= 3 and _is_matrix_like(args[2]):
        return to_numpy(re_ranking_matrix_device(*args, **kwargs))
    names = ["probFea", "galFea", "k1", "k2", "lambda_value", "local_distmat", "only_local"]
    params = dict(zip(names, args))
    params.update(kwargs)
    if "q_g" in params:
        return to_numpy(re_ranking_matrix_device(**params))
    return to_numpy(re_ranking_device(params["probFea"], params["galFea"], params["k1"], params["k2"],
                                      params["lambda_value"], params.get("local_distmat"),
                                      params.get("only_local", False)))


def topk_rows(mat, k: int, want_values: bool = False):
    """k smallest entries per row ascending by (value, column): the stable-argsort prefix
    np.argsort(mat, axis=1, kind='stable')[:, :k] (utils/reranking.py:48, utils/metrics.py:279)."""
    lib = _lib.require_device()
    m = _features(mat)
    rows, cols = m.shape
    idx = torch.empty((rows, k), dtype=torch.int32, device=m.device)
    val = torch.empty((rows, k), dtype=torch.float32, device=m.device) if want_values else None
    check(lib.demo_topk_rows(ptr(m), rows, cols, m.stride(0), int(k), ptr(idx), ptr(val), stream_ptr()))
    return (idx, val) if want_values else idx
