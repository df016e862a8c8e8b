"""CPU oracle for the DeMo retrieval hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm.  It is the
checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The shipped path (``demo2_b200``) never routes through this module and fails
loudly when its CUDA library is missing.

Parity pinning: the reference has NO tests or golden vectors for this path
(SURVEY.md section 4 / 8c), so this oracle is pinned against outputs of the
reference's own functions, imported in the build container by
``tests/golden/make_golden.py`` (matplotlib/seaborn stubbed); the resulting
vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py``.

Reference lines followed (paths relative to the reference repo):
  utils/metrics.py:12-107    eval_func_msrv (MSVR310 protocol + re.txt rank-list file)
  utils/metrics.py:110-169   eval_func
  utils/metrics.py:172-218   R1_mAP (MSVR310 evaluator)
  utils/metrics.py:341-369   R1_mAP_eval.compute (normalise -> split -> dist -> eval)
  utils/metrics.py:395-401   euclidean_distance
  utils/reranking.py:29-100  re_ranking
  layers/triplet_loss.py:5-135  normalize / euclidean_dist / cosine_dist /
                                hard_example_mining / TripletLoss

Tie rule.  The reference ranks with ``np.argsort`` (unstable); its order inside
a group of exactly equal distances is unspecified.  The oracle (and the CUDA
path) break ties by ascending gallery index (``kind='stable'``), see
SURVEY.md appendix A9.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F16 = np.float16


# --------------------------------------------------------------------------
# distances
# --------------------------------------------------------------------------
def l2_normalize(x, eps: float = 1e-12):
    """torch.nn.functional.normalize(x, dim=1, p=2): x / max(||x||, eps)
    (utils/metrics.py:345)."""
    x = np.ascontiguousarray(x, dtype=F32)
    n = np.sqrt(np.sum(x * x, axis=1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


def euclidean_distance(qf, gf):
    """Squared L2, no clamp, no sqrt: (|q|^2 (+) |g|^2^T) + (-2) q g^T in fp32
    (utils/metrics.py:395-401; addmm_ with beta=1, alpha=-2)."""
    qf = np.ascontiguousarray(qf, dtype=F32)
    gf = np.ascontiguousarray(gf, dtype=F32)
    qq = np.sum(qf * qf, axis=1, keepdims=True, dtype=F32)
    gg = np.sum(gf * gf, axis=1, keepdims=True, dtype=F32).T
    base = (qq + gg).astype(F32)
    dot = (qf @ gf.T).astype(F32)
    return (base + F32(-2.0) * dot).astype(F32)


def cosine_similarity(qf, gf):
    """q g^T / (|q| |g|^T).  Absent from the reference;
    defined as the similarity counterpart of cosine_dist: 1 - 2 * cosine_dist."""
    qf = np.ascontiguousarray(qf, dtype=F32)
    gf = np.ascontiguousarray(gf, dtype=F32)
    qn = np.sqrt(np.sum(qf * qf, axis=1, keepdims=True, dtype=F32))
    gn = np.sqrt(np.sum(gf * gf, axis=1, keepdims=True, dtype=F32)).T
    return ((qf @ gf.T).astype(F32) / (qn * gn)).astype(F32)


# --------------------------------------------------------------------------
# CMC / mAP
# --------------------------------------------------------------------------
def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank: int = 50, sort_kind: str = "stable",
              timing: dict | None = None):
    """Market-1501 protocol (utils/metrics.py:110-169) with the stable tie rule.
    Returns (cmc float32[max_rank], mAP float64).  sort_kind=None reproduces the reference's
    default (unstable) np.argsort -- used only to TIME the reference algorithm in bench.py, which
    also asks for the phase split (``timing``: seconds spent in the argsort / the per-query loop)."""
    import time as _time
    _t0 = _time.perf_counter()
    distmat = np.asarray(distmat)
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    num_q, num_g = distmat.shape
    if num_g < max_rank:  # :118-120
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    order_all = np.argsort(distmat, axis=1, kind=sort_kind)  # :121
    _t1 = _time.perf_counter()
    cmc_rows, aps = [], []
    for qi in range(num_q):
        order = order_all[qi]
        same_pid = g_pids[order] == q_pids[qi]
        junk = same_pid & (g_camids[order] == q_camids[qi])  # :136
        hits = same_pid[~junk].astype(np.int32)  # :141
        if not hits.any():  # :142-144
            continue
        cum = hits.cumsum()
        first = cum.copy()
        first[first > 1] = 1  # :146-147
        cmc_rows.append(first[:max_rank])
        prec = cum / (np.arange(1, cum.shape[0] + 1) * 1.0)  # :156-158
        aps.append((prec * hits).sum() / hits.sum())  # :159-160
    if timing is not None:
        timing["argsort_s"] = _t1 - _t0
        timing["loop_s"] = _time.perf_counter() - _t1
    assert len(aps) > 0, "Error: all query identities do not appear in gallery"  # :163
    cmc = np.asarray(cmc_rows).astype(F32).sum(0) / float(len(aps))  # :165-166
    return cmc, np.mean(aps)  # :167


def eval_func_msrv(distmat, q_pids, g_pids, q_camids, g_camids, q_sceneids, g_sceneids, max_rank: int = 50,
                   sort_kind: str = "stable", rank_file=None):
    """MSVR310 protocol (utils/metrics.py:12-107): a gallery item is discarded for a query when it
    has the same pid AND the same scene id (:67); cameras only appear in the rank-list file.
    The reference always (re)writes 're.txt' in the working directory (:38-39, :70-77): one header
    line, then per query '{pid}_s{scene}_v{cam}:' and the first max_rank kept gallery items as
    '{pid}_s{scene}_v{cam}  '.  rank_file=None skips the file; the text is returned as third value."""
    distmat = np.asarray(distmat)
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    q_sceneids, g_sceneids = np.asarray(q_sceneids), np.asarray(g_sceneids)
    num_q, num_g = distmat.shape
    if num_g < max_rank:  # :18-20
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    order_all = np.argsort(distmat, axis=1, kind=sort_kind)  # :21
    lines = ["rank list file\n"]  # :38-39
    cmc_rows, aps = [], []
    for qi in range(num_q):
        order = order_all[qi]
        same_pid = g_pids[order] == q_pids[qi]
        remove = same_pid & (g_sceneids[order] == q_sceneids[qi])  # :67
        keep = ~remove
        lines.append("{}_s{}_v{}:\n".format(q_pids[qi], q_sceneids[qi], q_camids[qi]))  # :71
        lines.append("".join("{}_s{}_v{}  ".format(a, c, b) for a, b, c in
                             zip(g_pids[order][keep][:max_rank], g_camids[order][keep][:max_rank],
                                 g_sceneids[order][keep][:max_rank])) + "\n")  # :72-77
        hits = same_pid[keep].astype(np.int32)  # :81
        if not hits.any():  # :82-84
            continue
        cum = hits.cumsum()
        first = cum.copy()
        first[first > 1] = 1
        cmc_rows.append(first[:max_rank])
        prec = cum / (np.arange(1, cum.shape[0] + 1) * 1.0)
        aps.append((prec * hits).sum() / hits.sum())
    assert len(aps) > 0, "Error: all query identities do not appear in gallery"  # :101
    text = "".join(lines)
    if rank_file:
        with open(rank_file, "w") as f:
            f.write(text)
    cmc = np.asarray(cmc_rows).astype(F32).sum(0) / float(len(aps))
    return cmc, np.mean(aps), text


def r1_map_msrv(feats, pids, camids, sceneids, num_query: int, feat_norm="yes", rank_file=None):
    """R1_mAP.compute (utils/metrics.py:193-218): normalise iff feat_norm == 'yes', split, squared
    distance, eval_func_msrv."""
    feats = np.asarray(feats, F32)
    if feat_norm == "yes":
        feats = l2_normalize(feats)
    qf, gf = feats[:num_query], feats[num_query:]
    pids, camids, sceneids = np.asarray(pids), np.asarray(camids), np.asarray(sceneids)
    distmat = euclidean_distance(qf, gf)
    cmc, mAP, text = eval_func_msrv(distmat, pids[:num_query], pids[num_query:], camids[:num_query],
                                    camids[num_query:], sceneids[:num_query], sceneids[num_query:],
                                    rank_file=rank_file)
    return cmc, mAP, distmat, qf, gf, text


def rank_counts(distmat, q_pids, g_pids, q_camids, g_camids):
    """Rank-count form of eval_func.

    For every query returns the positives sorted by (distance, gallery index) with
      r_p = 1 + #{valid g strictly before p},  c_p = 1 + #{positive g before p}
    where valid(g) = not(same pid and same cam) and "before" is lexicographic on
    (distance, index).  r_p - 1 and c_p - 1 are sums over gallery items, hence
    additive over any gallery partition.
    Returns (pos_ofs int64[Q+1], pos_idx int64[T], r int64[T], c int64[T])."""
    distmat = np.asarray(distmat)
    q_pids, g_pids = np.asarray(q_pids), np.asarray(g_pids)
    q_camids, g_camids = np.asarray(q_camids), np.asarray(g_camids)
    num_q = distmat.shape[0]
    ofs = np.zeros(num_q + 1, np.int64)
    idx_l, r_l, c_l = [], [], []
    for qi in range(num_q):
        order = np.argsort(distmat[qi], kind="stable")
        same_pid = g_pids[order] == q_pids[qi]
        junk = same_pid & (g_camids[order] == q_camids[qi])
        valid_rank = np.cumsum(~junk)  # 1-based rank among valid items
        pos = same_pid & ~junk
        where = np.nonzero(pos)[0]
        idx_l.append(order[where].astype(np.int64))
        r_l.append(valid_rank[where].astype(np.int64))
        c_l.append(np.arange(1, len(where) + 1, dtype=np.int64))
        ofs[qi + 1] = ofs[qi] + len(where)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return ofs, cat(idx_l), cat(r_l), cat(c_l)


def cmc_map_from_counts(pos_ofs, r, c, max_rank: int = 50, num_gallery=None):
    """Finalise CMC/mAP from rank counts exactly as eval_func would
    (float32 CMC accumulation :165-166, float64 mAP :167)."""
    if num_gallery is not None and num_gallery < max_rank:
        max_rank = num_gallery
    cmc = np.zeros(max_rank, F32)
    aps = []
    for qi in range(len(pos_ofs) - 1):
        s, e = pos_ofs[qi], pos_ofs[qi + 1]
        if e == s:
            continue
        rr, cc = r[s:e].astype(np.float64), c[s:e].astype(np.float64)
        aps.append((cc / rr).sum() / float(e - s))
        first = int(r[s:e].min())
        if first <= max_rank:
            cmc[first - 1:] += F32(1.0)
    assert len(aps) > 0, "Error: all query identities do not appear in gallery"
    return (cmc / float(len(aps))).astype(F32), np.mean(aps)


# --------------------------------------------------------------------------
# k-reciprocal re-ranking
# --------------------------------------------------------------------------
def all_pairs_sqdist(probFea, galFea):
    """utils/reranking.py:36-41: squared distances over cat(query, gallery)."""
    feat = np.concatenate([np.asarray(probFea, F32), np.asarray(galFea, F32)], axis=0)
    return euclidean_distance(feat, feat)


def stable_topk(mat, k: int):
    """First k columns of np.argsort(mat, axis=1, kind='stable') without the full sort."""
    n_rows, n_cols = mat.shape
    k = min(k, n_cols)
    if n_cols <= 4 * k or n_cols < 256:
        return np.argsort(mat, axis=1, kind="stable")[:, :k].astype(np.int32)
    out = np.empty((n_rows, k), np.int32)
    kth = np.partition(mat, k - 1, axis=1)[:, k - 1]
    for i in range(n_rows):
        row = mat[i]
        less = np.nonzero(row < kth[i])[0]
        eq = np.nonzero(row == kth[i])[0][: k - len(less)]
        cand = np.concatenate([less, eq])
        cand = cand[np.lexsort((cand, row[cand]))]
        out[i] = cand
    return out


def _np_sum_f32(w):
    return np.sum(w)  # numpy pairwise float32 sum


def k_reciprocal_rows(od, rank, k1: int, row0: int = 0):
    """utils/reranking.py:51-71.  Returns per-row (sorted unique index array,
    float16 weights) -- the non-zeros of V before query expansion.  ``od`` may hold only the
    rows [row0, row0 + len(od)) of the normalised matrix (row-sharded tests); ``rank`` is global."""
    n = od.shape[0]
    kh = int(np.around(k1 / 2)) + 1  # half-to-even (appendix A7)
    rows = []
    for li in range(n):
        i = li + row0
        fwd = rank[i, : k1 + 1]
        bwd = rank[fwd, : k1 + 1]
        kri = fwd[np.nonzero(bwd == i)[0]]
        expn = kri
        for cand in kri:
            cf = rank[cand, :kh]
            cb = rank[cf, :kh]
            ckri = cf[np.nonzero(cb == cand)[0]]
            if len(np.intersect1d(ckri, kri)) > 2 / 3 * len(ckri):
                expn = np.append(expn, ckri)
        expn = np.unique(expn)
        w = np.exp(-od[li, expn])  # float32
        rows.append((expn.astype(np.int64), (w / _np_sum_f32(w)).astype(F16)))
    return rows


def re_ranking_from_allpairs(allpairs, query_num: int, k1: int, k2: int, lambda_value: float):
    """Core of utils/reranking.py:45-100 given the all-pairs squared-distance
    matrix (after the optional +local_distmat).  Sparse bookkeeping, identical
    arithmetic (float16 stores / adds as in the reference)."""
    D = np.asarray(allpairs, dtype=F32)
    n = D.shape[0]
    od = np.transpose(D / np.max(D, axis=0)).astype(F32)  # :46
    rank = stable_topk(od, max(k1 + 1, k2))  # :48 (only [:k1+1] and [:k2] are read)
    rows = k_reciprocal_rows(od, rank, k1)  # :51-71
    rows = [(idx[val != 0], val[val != 0]) for idx, val in rows]  # V != 0 tests (:82, :88)

    if k2 != 1:  # :73-78  mean over the k2 nearest rows: fp32 sequential sum, /k2, -> fp16
        exp_rows = []
        for i in range(n):
            acc = {}
            for nb in rank[i, :k2]:
                idx, val = rows[nb]
                for c, v in zip(idx.tolist(), val.astype(F32).tolist()):
                    acc[c] = F32(acc.get(c, F32(0.0)) + F32(v))
            cols = np.fromiter(sorted(acc), dtype=np.int64, count=len(acc))
            vals = np.array([acc[c] for c in cols.tolist()], dtype=F32)
            vals = (vals / F32(k2)).astype(F16)
            keep = vals != 0
            exp_rows.append((cols[keep], vals[keep]))
        rows = exp_rows

    # :80-82 inverted index: for each column j the rows t with V[t, j] != 0 (ascending t)
    inv_rows = [[] for _ in range(n)]
    inv_vals = [[] for _ in range(n)]
    for t in range(n):
        idx, val = rows[t]
        for c, v in zip(idx.tolist(), val.tolist()):
            inv_rows[c].append(t)
            inv_vals[c].append(v)
    inv_rows = [np.asarray(a, np.int64) for a in inv_rows]
    inv_vals = [np.asarray(a, F16) for a in inv_vals]

    jaccard = np.zeros((query_num, n), dtype=F16)
    for i in range(query_num):  # :86-93
        temp_min = np.zeros(n, dtype=F16)
        idx, val = rows[i]
        for j, vij in zip(idx.tolist(), val):
            t = inv_rows[j]
            temp_min[t] = temp_min[t] + np.minimum(vij, inv_vals[j])
        jaccard[i] = 1 - temp_min / (2 - temp_min)
    final = jaccard * (1 - lambda_value) + od[:query_num] * lambda_value  # :95
    return np.ascontiguousarray(final[:query_num, query_num:]).astype(F32)  # :99


def re_ranking(probFea, galFea, k1: int, k2: int, lambda_value: float,
               local_distmat=None, only_local: bool = False):
    """utils/reranking.py:29-100 (feature form)."""
    query_num = np.asarray(probFea).shape[0]
    if only_local:
        D = np.asarray(local_distmat, F32)
    else:
        D = all_pairs_sqdist(probFea, galFea)
        if local_distmat is not None:
            D = D + np.asarray(local_distmat, F32)
    return re_ranking_from_allpairs(D, query_num, k1, k2, lambda_value)


def re_ranking_dense(allpairs, query_num: int, k1: int, k2: int, lambda_value: float):
    """Dense, line-by-line form of utils/reranking.py:45-100 (N x N float16 V).
    Only for small N; used to cross-check the sparse bookkeeping above."""
    D = np.asarray(allpairs, dtype=F32)
    n = D.shape[0]
    od = np.transpose(D / np.max(D, axis=0))
    V = np.zeros_like(od).astype(F16)
    rank = np.argsort(od, kind="stable").astype(np.int32)
    for i, (idx, w) in enumerate(k_reciprocal_rows(od, rank, k1)):
        V[i, idx] = w
    od = od[:query_num]
    if k2 != 1:
        V_qe = np.zeros_like(V, dtype=F16)
        for i in range(n):
            V_qe[i, :] = np.mean(V[rank[i, :k2], :], axis=0)
        V = V_qe
    inv = [np.where(V[:, j] != 0)[0] for j in range(n)]
    jac = np.zeros_like(od, dtype=F16)
    for i in range(query_num):
        temp_min = np.zeros(shape=[1, n], dtype=F16)
        nz = np.where(V[i, :] != 0)[0]
        for j in nz:
            temp_min[0, inv[j]] = temp_min[0, inv[j]] + np.minimum(V[i, j], V[inv[j], j])
        jac[i] = 1 - temp_min / (2 - temp_min)
    final = jac * (1 - lambda_value) + od * lambda_value
    return final[:query_num, query_num:]


# --------------------------------------------------------------------------
# evaluator (utils/metrics.py:221-248, 341-369)
# --------------------------------------------------------------------------
def r1_map_eval(feats, pids, camids, num_query: int, feat_norm=True, reranking=False,
                rerank_params=(50, 15, 0.3)):
    """compute(): normalise, split at num_query, distance or re-ranking, eval_func.
    The reference hard-codes k1=50, k2=15, lambda=0.3 (:359)."""
    feats = np.asarray(feats, F32)
    if feat_norm:
        feats = l2_normalize(feats)
    qf, gf = feats[:num_query], feats[num_query:]
    pids, camids = np.asarray(pids), np.asarray(camids)
    if reranking:
        k1, k2, lam = rerank_params
        distmat = re_ranking(qf, gf, k1, k2, lam)
    else:
        distmat = euclidean_distance(qf, gf)
    cmc, mAP = eval_func(distmat, pids[:num_query], pids[num_query:],
                         camids[:num_query], camids[num_query:])
    return cmc, mAP, distmat, qf, gf


# --------------------------------------------------------------------------
# triplet loss (layers/triplet_loss.py)
# --------------------------------------------------------------------------
def triplet_normalize(x):
    """:5-13  x / (||x|| + 1e-12)."""
    x = np.asarray(x, F32)
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=F32))
    return (x / (n + F32(1e-12))).astype(F32)


def euclidean_dist(x, y):
    """:16-31  sqrt(clamp(|x|^2 + |y|^2^T - 2 x y^T, 1e-12))."""
    x, y = np.asarray(x, F32), np.asarray(y, F32)
    xx = np.sum(x * x, axis=1, keepdims=True, dtype=F32)
    yy = np.sum(y * y, axis=1, keepdims=True, dtype=F32).T
    d = (xx + yy) - F32(2.0) * (x @ y.T).astype(F32)
    return np.sqrt(np.maximum(d, F32(1e-12))).astype(F32)


def cosine_dist(x, y):
    """:34-48  (1 - x y^T / (|x| |y|^T)) / 2."""
    x, y = np.asarray(x, F32), np.asarray(y, F32)
    xn = np.sqrt(np.sum(x * x, axis=1, keepdims=True, dtype=F32))
    yn = np.sqrt(np.sum(y * y, axis=1, keepdims=True, dtype=F32)).T
    return ((F32(1.0) - (x @ y.T).astype(F32) / (xn * yn)) / F32(2.0)).astype(F32)


def hard_example_mining(dist_mat, labels, return_inds: bool = False):
    """:51-104  hardest positive (max, self included) / hardest negative (min) per
    anchor.  Requires the same number of positives for every anchor (the
    reference's view(N, -1) raises otherwise).  Ties -> lowest index."""
    dist_mat = np.asarray(dist_mat)
    labels = np.asarray(labels)
    n = dist_mat.shape[0]
    assert dist_mat.ndim == 2 and dist_mat.shape[1] == n
    is_pos = labels[None, :] == labels[:, None]
    npos = is_pos.sum(1)
    if not (npos == npos[0]).all():
        raise RuntimeError("hard_example_mining: anchors have different numbers of positives")
    ap_src = np.where(is_pos, dist_mat, -np.inf)
    an_src = np.where(~is_pos, dist_mat, np.inf)
    p_inds = ap_src.argmax(1)
    n_inds = an_src.argmin(1)
    dist_ap = dist_mat[np.arange(n), p_inds]
    dist_an = dist_mat[np.arange(n), n_inds]
    if return_inds:
        return dist_ap, dist_an, p_inds.astype(np.int64), n_inds.astype(np.int64)
    return dist_ap, dist_an


def triplet_loss(global_feat, labels, margin=None, hard_factor: float = 0.0,
                 normalize_feature: bool = False):
    """:121-135  returns (loss, dist_ap, dist_an).  SoftMarginLoss (margin None)
    = mean(log(1 + exp(-(an - ap)))); MarginRankingLoss = mean(max(0, ap - an + m))."""
    x = np.asarray(global_feat, F32)
    if normalize_feature:
        x = triplet_normalize(x)
    d = euclidean_dist(x, x)
    ap, an = hard_example_mining(d, labels)
    ap = (ap * F32(1.0 + hard_factor)).astype(F32)
    an = (an * F32(1.0 - hard_factor)).astype(F32)
    if margin is not None:
        loss = np.maximum(F32(0.0), ap - an + F32(margin)).mean(dtype=np.float64)
    else:
        loss = np.log1p(np.exp(-(an - ap).astype(np.float64))).mean()
    return F32(loss), ap, an


def triplet_loss_grad(global_feat, labels, margin=None, hard_factor: float = 0.0):
    """d loss / d x for the un-normalised path (float64 maths), used to check the
    CUDA backward: only the (anchor, hardest-positive) and (anchor,
    hardest-negative) pairs carry gradient; d dist(a,b)/d x_a = (x_a - x_b)/dist."""
    x = np.asarray(global_feat, np.float64)
    n = x.shape[0]
    d = euclidean_dist(x.astype(F32), x.astype(F32)).astype(np.float64)
    _, _, pi, ni = hard_example_mining(d, labels, return_inds=True)
    ap = d[np.arange(n), pi] * (1.0 + hard_factor)
    an = d[np.arange(n), ni] * (1.0 - hard_factor)
    if margin is not None:
        act = (ap - an + margin) > 0
        g_ap = act / n * (1.0 + hard_factor)
        g_an = -act / n * (1.0 - hard_factor)
    else:
        s = 1.0 / (1.0 + np.exp(an - ap))  # sigmoid(-(an-ap))
        g_ap = s / n * (1.0 + hard_factor)
        g_an = -s / n * (1.0 - hard_factor)
    grad = np.zeros_like(x)
    for a in range(n):
        for j, gcoef in ((pi[a], g_ap[a]), (ni[a], g_an[a])):
            dist = d[a, j]
            if dist * dist <= 1e-12:  # clamp region: zero gradient
                continue
            v = (x[a] - x[j]) / dist * gcoef
            grad[a] += v
            grad[j] -= v
    return grad


# --------------------------------------------------------------------------
# distance-matrix batch losses (layers/cluster_loss.py, layers/range_loss.py)
# --------------------------------------------------------------------------
def batch_identities(targets, ordered: bool, ids_per_batch: int, imgs_per_id: int):
    """cluster_loss.py:44-59 / range_loss.py:103-118: every imgs_per_id-th label of a P x K
    ordered batch, else the sorted unique labels."""
    targets = np.asarray(targets)
    if ordered and targets.shape[0] == ids_per_batch * imgs_per_id:
        return targets[0:targets.shape[0]:imgs_per_id]
    return np.unique(targets)


def cluster_loss(features, targets, margin=10, ordered=True, ids_per_batch=16, imgs_per_id=4):
    """cluster_loss.py:33-86, identity by identity as the reference loops.
    Returns (loss, intra_max_distance [P], inter_min_distance [P])."""
    x, t = np.asarray(features, F32), np.asarray(targets)
    labels = batch_identities(t, ordered, ids_per_batch, imgs_per_id)
    P = labels.shape[0]
    centers = np.zeros((P, x.shape[1]), F32)
    intra = np.zeros(P, F32)
    inter = np.zeros(P, F32)
    for i in range(P):
        same = x[t == labels[i]]
        centers[i] = same.mean(axis=0, dtype=F32)                           # :72-73
        intra[i] = euclidean_dist(centers[i:i + 1], same).max()            # :74-76
    for i in range(P):
        others = np.arange(P) != i
        inter[i] = euclidean_dist(centers[i:i + 1], centers[others]).min()  # :79-82
    loss = np.maximum(intra - inter + F32(margin), F32(0)).mean(dtype=F32)  # :85
    return F32(loss), intra, inter


def range_loss(features, targets, k=2, margin=0.1, alpha=0.5, beta=0.5, ordered=True, ids_per_batch=32,
               imgs_per_id=4):
    """range_loss.py:38-201 with the reference's own selection rules: the k largest intra-class
    distances are every second element of the tail of the sorted flattened matrix (:62), the
    smallest centre distance is element [n] of the sorted centre matrix (:89).
    Returns (range_loss, intra_class_loss, inter_class_loss)."""
    x, t = np.asarray(features, F32), np.asarray(targets)
    labels = batch_identities(t, ordered, ids_per_batch, imgs_per_id)
    P = labels.shape[0]
    centers = np.stack([x[t == labels[i]].mean(axis=0, dtype=F32) for i in range(P)])   # :120-130
    cc = np.sort(euclidean_dist(centers, centers).reshape(-1), kind="stable")
    inter = np.maximum(F32(margin) - cc[P], F32(0))                          # :89, :147
    intra = np.zeros(P, F32)
    for i in range(P):
        same = x[t == labels[i]]
        flat = np.sort(euclidean_dist(same, same).reshape(-1), kind="stable")
        top_k = flat[-k * 2::2]                                              # :62
        intra[i] = F32(k) / np.sum(F32(1.0) / top_k, dtype=F32)              # :183-184
    intra_loss = intra.sum(dtype=F32)
    return F32(F32(alpha) * intra_loss + F32(beta) * inter), F32(intra_loss), F32(inter)
