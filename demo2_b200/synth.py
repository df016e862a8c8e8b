"""Seeded synthetic ReID features.

The same generator feeds the CPU oracle, the golden-vector script and the CUDA
path, so every comparison is on identical tensors.  Shapes follow the
reference's own logs (RGBNT201: 836/836/30 ids/2 cams, RGBNT100:
1715/8575/50/8; d = 3 x 512 from modeling/make_model.py:470,736).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

# name -> (Q, G, d, nid, ncam)
SHAPES = {
    "rgbnt201": (836, 836, 1536, 30, 2),
    "rgbnt100": (1715, 8575, 1536, 50, 8),
    "msvr310": (591, 1055, 1536, 52, 8),
    "large": (20000, 1000000, 1536, 50000, 8),
}


@dataclass
class ReIDSet:
    qf: torch.Tensor  # [Q, d] fp32, NOT normalised
    gf: torch.Tensor  # [G, d] fp32, NOT normalised
    q_pids: np.ndarray  # int64 [Q]
    g_pids: np.ndarray
    q_camids: np.ndarray
    g_camids: np.ndarray

    @property
    def num_query(self) -> int:
        return int(self.qf.shape[0])


def make_reid_set(Q: int, G: int, d: int, nid: int, ncam: int, sigma: float = 4.0,
                  seed: int = 0, chunk: int = 65536) -> ReIDSet:
    """Clustered features: centre[pid] + sigma * noise.  Draw order is part of the
    contract: centres, q_pid, g_pid, q_cam, g_cam, q noise, g noise."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    centers = torch.randn(nid, d, generator=g)
    q_pids = rng.integers(0, nid, Q)
    g_pids = rng.integers(0, nid, G)
    q_camids = rng.integers(0, ncam, Q)
    g_camids = rng.integers(0, ncam, G)

    def draw(pids: np.ndarray) -> torch.Tensor:
        n = len(pids)
        out = torch.empty(n, d)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            idx = torch.from_numpy(pids[s:e])
            out[s:e] = torch.randn(e - s, d, generator=g).mul_(sigma).add_(centers[idx])
        return out

    # NOTE: for n <= chunk this is exactly centres[pid] + sigma * randn(n, d, g).
    qf = draw(q_pids)
    gf = draw(g_pids)
    return ReIDSet(qf, gf, q_pids, g_pids, q_camids, g_camids)


def make_named(name: str, sigma: float = 4.0, seed: int = 0) -> ReIDSet:
    Q, G, d, nid, ncam = SHAPES[name]
    return make_reid_set(Q, G, d, nid, ncam, sigma=sigma, seed=seed)


def make_triplet_batch(n_ids: int = 8, n_inst: int = 16, d: int = 768, seed: int = 0,
                       modalities: int = 3):
    """BASELINE config #5: PK batch (8 ids x 16 instances), un-normalised fp32, one
    tensor per modality (configs/RGBNT100/DeMo.yml:23,36; make_model.py:743-746)."""
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(n_ids * n_inst, d, generator=g) for _ in range(modalities)]
    labels = torch.arange(n_ids).repeat_interleave(n_inst)
    return xs, labels
