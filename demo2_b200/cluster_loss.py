"""Drop-in for ``ClusterLoss`` of the reference's ``layers/cluster_loss.py:8-101``.

The reference loops over the identities of the batch and calls ``_euclidean_dist`` (:17-31) twice per
identity (centre -> its samples, centre -> the other centres: 2P small GEMMs and 2P host syncs for
``.max()`` / ``.min()`` into pre-allocated vectors).  Here the class centres are one segment mean,
and ONE launch of the library's distance entry point (``demo_sqdist_f32`` in sqrt mode, through the
differentiable ``euclidean_dist``) produces the [P, B + P] matrix centre x (samples | centres); the
intra-class maximum and the inter-class minimum are masked row reductions of that matrix.

``ClusterLoss_local`` (:104-248, a dynamic-programming alignment over local stripes) is not on the
distance-matrix path of SURVEY.md section 8 and is not provided.
"""
from __future__ import annotations

import torch
from torch import nn

from .triplet_loss import euclidean_dist


def batch_identities(targets: torch.Tensor, ordered: bool, ids_per_batch: int, imgs_per_id: int) -> torch.Tensor:
    """The identities of a batch in the order the reference enumerates them (:44-59): every
    ``imgs_per_id``-th label of a P x K ordered batch, otherwise the sorted unique labels."""
    if ordered and targets.size(0) == ids_per_batch * imgs_per_id:
        return targets[0:targets.size(0):imgs_per_id]
    return targets.unique()


def class_centers(features: torch.Tensor, targets: torch.Tensor, unique_labels: torch.Tensor):
    """Mean feature of every identity (:70-73) as one [P, B] membership product; returns the
    centres and the boolean membership matrix."""
    member = unique_labels.unsqueeze(1).eq(targets.unsqueeze(0))               # [P, B]
    w = member.to(features.dtype)
    centers = (w @ features) / w.sum(1, keepdim=True)
    return centers, member


class ClusterLoss(nn.Module):
    """mean_i relu(max_{x in class i} |c_i - x| - min_{j != i} |c_i - c_j| + margin)   (:85).

    ``forward(features [B, d], targets [B]) -> (loss, intra_max_distance [P], inter_min_distance [P])``.
    ``use_gpu`` is accepted for signature compatibility; the distances are always computed on the B200.
    """

    def __init__(self, margin=10, use_gpu=True, ordered=True, ids_per_batch=16, imgs_per_id=4):
        super(ClusterLoss, self).__init__()
        self.use_gpu = use_gpu
        self.margin = margin
        self.ordered = ordered
        self.ids_per_batch = ids_per_batch
        self.imgs_per_id = imgs_per_id

    def _cluster_loss(self, features, targets, ordered=True, ids_per_batch=16, imgs_per_id=4):
        features = features.cuda().float()
        targets = targets.to(features.device)
        unique_labels = batch_identities(targets, ordered, ids_per_batch, imgs_per_id)
        P, B = unique_labels.size(0), features.size(0)
        if P < 2:  # the reference takes .min() of an empty [1, 0] matrix here (:81)
            raise RuntimeError("ClusterLoss needs at least two identities in the batch")
        centers, member = class_centers(features, targets, unique_labels)
        dist = euclidean_dist(centers, torch.cat([features, centers], dim=0))   # [P, B + P], one launch
        intra = dist[:, :B].masked_fill(~member, float("-inf"))
        intra_max_distance = intra.max(dim=1)[0]                                # :74-76
        # a label that occurs twice in `unique_labels` (un-ordered batch declared ordered) is a
        # different centre with the same value in the reference as well: only the diagonal is excluded
        eye = torch.eye(P, dtype=torch.bool, device=dist.device)
        inter = dist[:, B:].masked_fill(eye, float("inf"))
        inter_min_distance = inter.min(dim=1)[0]                                # :79-82
        cluster_loss = torch.mean(torch.relu(intra_max_distance - inter_min_distance + self.margin))
        return cluster_loss, intra_max_distance, inter_min_distance

    def forward(self, features, targets):
        assert features.size(0) == targets.size(0), "features.size(0) is not equal to targets.size(0)"
        return self._cluster_loss(features, targets, self.ordered, self.ids_per_batch, self.imgs_per_id)
