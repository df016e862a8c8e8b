"""Drop-in for ``RangeLoss`` of the reference's ``layers/range_loss.py:7-214``.

    range_loss = alpha * intra_class_loss + beta * inter_class_loss
    intra_class_loss = sum over identities of the harmonic mean of the k largest intra-class distances
    inter_class_loss = relu(margin - smallest distance between two class centres)

The reference calls ``_pairwise_distance`` (:24-36) once per identity plus once for the centres and
sorts every flattened matrix on the way (``[-2k::2]`` of the sorted values: each unordered pair
occurs twice in a symmetric matrix, :62; element ``[n]`` of the sorted centre matrix: the n
self-distances come first, :89).  Here ONE launch of the library's distance entry point
(``demo_sqdist_f32`` in sqrt mode, through the differentiable ``euclidean_dist``) gives the
(B + P)^2 matrix over (samples | centres); the k largest distances of every identity are a masked
top-k over its unordered pairs (a < b) and the centre term is the minimum off-diagonal entry --
no per-identity loop, no host synchronisation.

Defined where the reference is degenerate: an identity with fewer than k unordered pairs makes the
reference's slice pick self-distances (the clamp floor sqrt(1e-12) = 1e-6 up to cancellation noise);
here the missing entries are exactly 1e-6.
"""
from __future__ import annotations

import torch
from torch import nn

from .cluster_loss import batch_identities, class_centers
from .triplet_loss import euclidean_dist

_SELF_DISTANCE = 1e-6   # sqrt(clamp(0, min=1e-12)), range_loss.py:35


class RangeLoss(nn.Module):
    """``forward(features [B, d], targets [B]) -> (range_loss, intra_class_loss, inter_class_loss)``.
    ``use_gpu`` is accepted for signature compatibility; the distances are always computed on the B200."""

    def __init__(self, k=2, margin=0.1, alpha=0.5, beta=0.5, use_gpu=True, ordered=True, ids_per_batch=32,
                 imgs_per_id=4):
        super(RangeLoss, self).__init__()
        self.use_gpu = use_gpu
        self.margin = margin
        self.k = k
        self.alpha = alpha
        self.beta = beta
        self.ordered = ordered
        self.ids_per_batch = ids_per_batch
        self.imgs_per_id = imgs_per_id

    def _range_loss(self, features, targets, ordered, ids_per_batch, imgs_per_id):
        features = features.cuda().float()
        targets = targets.to(features.device)
        unique_labels = batch_identities(targets, ordered, ids_per_batch, imgs_per_id)
        P, B = unique_labels.size(0), features.size(0)
        centers, member = class_centers(features, targets, unique_labels)       # :92-130
        z = torch.cat([features, centers], dim=0)
        dist = euclidean_dist(z, z)                                             # [(B + P), (B + P)], one launch

        # inter-class: smallest centre-to-centre distance (:65-90, :132-147)
        cc = dist[B:, B:]
        if P < 2:  # the reference indexes element [n] of a sorted 1-element matrix here (:89)
            raise IndexError("RangeLoss needs at least two identities in the batch")
        off_diag = ~torch.eye(P, dtype=torch.bool, device=dist.device)
        min_inter_class_center_distance = cc.masked_fill(~off_diag, float("inf")).min()
        inter_class_loss = torch.relu(self.margin - min_inter_class_center_distance)

        # intra-class: k largest distances among the unordered pairs of every identity (:38-63, :149-186)
        upper = torch.ones(B, B, dtype=torch.bool, device=dist.device).triu(1)
        pairs = member.unsqueeze(2) & member.unsqueeze(1) & upper.unsqueeze(0)  # [P, B, B]
        ss = dist[:B, :B].unsqueeze(0).expand(P, B, B).masked_fill(~pairs, float("-inf")).reshape(P, B * B)
        top_k = ss.topk(min(self.k, B * B), dim=1)[0]
        top_k = torch.where(torch.isinf(top_k), torch.full_like(top_k, _SELF_DISTANCE), top_k)
        if top_k.size(1) < self.k:
            pad = top_k.new_full((P, self.k - top_k.size(1)), _SELF_DISTANCE)
            top_k = torch.cat([top_k, pad], dim=1)
        intra_distance = self.k / torch.sum(1.0 / top_k, dim=1)                 # :183-184
        intra_class_loss = torch.sum(intra_distance)

        range_loss = self.alpha * intra_class_loss + self.beta * inter_class_loss
        return range_loss, intra_class_loss, inter_class_loss

    def forward(self, features, targets):
        assert features.size(0) == targets.size(0), "features.size(0) is not equal to targets.size(0)"
        return self._range_loss(features, targets, self.ordered, self.ids_per_batch, self.imgs_per_id)
