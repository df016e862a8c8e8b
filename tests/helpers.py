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
