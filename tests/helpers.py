"""Shared test helpers: seeded cases (inputs are regenerated, never stored)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from demo2_b200 import synth  # noqa: E402
from oracle import reid_oracle as oracle  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def make_case(shape: str, seed: int, sigma: float, gallery_is_query: bool = False):
    """Returns (qf, gf, q_pids, g_pids, q_camids, g_camids) with qf/gf L2-normalised by
    the oracle's numpy normalisation (bit-reproducible across hosts)."""
    s = synth.make_named(shape, sigma=sigma, seed=seed)
    qf = oracle.l2_normalize(s.qf.numpy())
    if gallery_is_query:
        return qf, qf.copy(), s.q_pids, s.q_pids.copy(), s.q_camids, s.q_camids.copy()
    gf = oracle.l2_normalize(s.gf.numpy())
    return qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids


def make_scene_ids(num_q: int, num_g: int, seed: int, nscene: int = 6):
    """Scene ids for the MSVR310 protocol (utils/metrics.py:67), seeded separately so the other
    cases stay unchanged."""
    rng = np.random.default_rng(100000 + seed)
    return rng.integers(0, nscene, num_q), rng.integers(0, nscene, num_g)


def sample_index(n_rows: int, n_cols: int, count: int = 4096):
    """Deterministic scattered sample of a matrix (flat indices)."""
    total = n_rows * n_cols
    step = max(1, total // count)
    return np.arange(0, total, step, dtype=np.int64)[:count]


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


# ClusterLoss / RangeLoss fixtures.  P identities x K images, features = spread * centre
# + noise; `ordered` False shuffles the batch and drops a few samples (ragged identities).
BATCH_LOSS_CASES = {
    "pk8x16": dict(P=8, K=16, d=768, seed=11, spread=0.6, ordered=True, k=2, cluster_margin=10, range_margin=0.1),
    "pk16x4": dict(P=16, K=4, d=2048, seed=12, spread=0.5, ordered=True, k=2, cluster_margin=10, range_margin=60.0),
    "ragged": dict(P=6, K=9, d=512, seed=13, spread=0.7, ordered=False, k=3, cluster_margin=5, range_margin=25.0),
}


def batch_loss_case(name: str):
    """(features fp32 [B, d], targets int64 [B]) of a BATCH_LOSS_CASES entry, as CPU torch tensors."""
    import torch
    spec = BATCH_LOSS_CASES[name]
    g = torch.Generator().manual_seed(spec["seed"])
    P, K, d = spec["P"], spec["K"], spec["d"]
    centers = torch.randn(P, d, generator=g) * spec["spread"]
    ids = torch.randperm(100, generator=g)[:P]          # arbitrary, unsorted identity numbers
    targets = ids.repeat_interleave(K)
    feats = centers.repeat_interleave(K, dim=0) + torch.randn(P * K, d, generator=g)
    if not spec["ordered"]:
        keep = torch.randperm(P * K, generator=g)[:P * K - 5]
        feats, targets = feats[keep].contiguous(), targets[keep].contiguous()
    return feats, targets


# ---- gap-masked rank parity (north_star: rank indices exact wherever the reference's distance gap
# exceeds the tolerance) -------------------------------------------------------------------------
GAP_RTOL, GAP_ATOL = 1e-5, 4e-6     # 1e-5 relative on the positive's distance + the absolute floor of two distances


def canonical_ranks_from_oracle(dist, qp, gp, qc, gc):
    """(pos_ofs, gidx, r) of oracle.rank_counts re-ordered canonically (per query by gallery index)."""
    ofs, idx, r, _ = oracle.rank_counts(dist, qp, gp, qc, gc)
    g2, r2 = idx.copy(), r.copy()
    for q in range(len(ofs) - 1):
        s, e = ofs[q], ofs[q + 1]
        o = np.argsort(idx[s:e], kind="stable")
        g2[s:e], r2[s:e] = idx[s:e][o], r[s:e][o]
    return ofs, g2, r2


def compare_ranks_outside_near_ties(pos_ofs, ranks, golden, report_name=None):
    """Compares per-positive ranks (canonical order) with the reference's (tests/golden/posrank_*):
    EXACT equality for every positive whose reference gap exceeds the tolerance, exact AP / first
    rank for the queries without any near-tied positive.  Returns a dict of statistics (masked
    fractions, mismatches inside the mask) for the parity report."""
    g_ofs, g_rank, gap, gd = golden["pos_ofs"], golden["rank"].astype(np.int64), golden["gap"], golden["dist"]
    np.testing.assert_array_equal(pos_ofs, g_ofs)
    clear = gap > GAP_RTOL * np.abs(gd) + GAP_ATOL
    ranks = np.asarray(ranks, np.int64)
    bad = np.nonzero(clear & (ranks != g_rank))[0]
    assert bad.size == 0, "rank differs on %d positives with a clear gap (first: #%d ours %d ref %d gap %.3g)" % (
        bad.size, bad[0], ranks[bad[0]], g_rank[bad[0]], gap[bad[0]])
    Q = len(g_ofs) - 1
    q_clear = np.array([bool(clear[g_ofs[q]:g_ofs[q + 1]].all()) and g_ofs[q + 1] > g_ofs[q] for q in range(Q)])
    first_clear = np.zeros(Q, bool)      # the reference's best positive has a clear gap
    ap_diff, first_bad = 0.0, 0

    def ap_of(r):
        rs = np.sort(r)
        return float(((np.arange(len(rs)) + 1) / rs).sum() / len(rs))

    for q in range(Q):
        s, e = g_ofs[q], g_ofs[q + 1]
        if e == s:
            continue
        best = s + int(np.argmin(g_rank[s:e]))
        first_clear[q] = clear[best]
        if first_clear[q]:
            first_bad += int(ranks[s:e].min() != g_rank[s:e].min())
        if q_clear[q]:
            ap_diff = max(ap_diff, abs(ap_of(ranks[s:e]) - ap_of(g_rank[s:e])))
    assert first_bad == 0, "first-match rank differs on %d queries whose best positive has a clear gap" % first_bad
    assert ap_diff <= 1e-12
    stats = {"positives": int(len(g_rank)), "masked_positive_frac": float(1.0 - clear.mean()),
             "mismatch_inside_mask": int(((ranks != g_rank) & ~clear).sum()),
             "queries_all_clear_frac": float(q_clear.mean()), "queries_first_clear_frac": float(first_clear.mean())}
    if report_name:
        write_parity_report(report_name, stats)
    return stats


def write_parity_report(name, stats):
    """Measured worst cases of the parity tests, appended to gpurun_out/parity_report.jsonl (scratch;
    the summary is copied into profiles/ by hand)."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, **stats}) + "\n")
    except OSError:
        pass
