"""GPU parity for the ClusterLoss / RangeLoss drop-ins (layers/cluster_loss.py:8-101,
layers/range_loss.py:7-214): values and gradients minted from the reference
(tests/golden/make_golden.py::batch_loss_cases) and the CPU oracle on the same inputs.
fp32 throughout; tolerance 1e-5 relative on the outputs (BASELINE.json), 2e-5 on gradients
(a gradient entry is a difference of O(1) features divided by a distance)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import BATCH_LOSS_CASES, batch_loss_case, load_golden, oracle

pytestmark = pytest.mark.gpu


def _kw(spec):
    return dict(ordered=spec["ordered"], ids_per_batch=spec["P"], imgs_per_id=spec["K"])


@pytest.mark.parametrize("name", ["pk8x16", "pk16x4", "ragged"])
def test_cluster_loss_matches_reference(name):
    from demo2_b200.cluster_loss import ClusterLoss
    spec, g = BATCH_LOSS_CASES[name], load_golden("batch_loss_" + name)
    feats, targets = batch_loss_case(name)
    x = feats.cuda().requires_grad_(True)
    loss, intra, inter = ClusterLoss(margin=spec["cluster_margin"], **_kw(spec))(x, targets.cuda())
    loss.backward()
    np.testing.assert_allclose(intra.detach().cpu().numpy(), g["cl_intra"], rtol=1e-5)
    np.testing.assert_allclose(inter.detach().cpu().numpy(), g["cl_inter"], rtol=1e-5)
    np.testing.assert_allclose(float(loss.detach()), float(g["cl_loss"]), rtol=1e-5)
    gx = x.grad.cpu().numpy()
    np.testing.assert_allclose(gx.ravel()[::41], g["cl_gx"], rtol=2e-5, atol=1e-7)
    assert abs(float(np.abs(gx.astype(np.float64)).sum()) - float(g["cl_gx_abs"])) < 1e-5 * float(g["cl_gx_abs"])
    o_loss, o_intra, o_inter = oracle.cluster_loss(feats.numpy(), targets.numpy(), margin=spec["cluster_margin"],
                                                   **_kw(spec))
    np.testing.assert_allclose(intra.detach().cpu().numpy(), o_intra, rtol=1e-5)
    np.testing.assert_allclose(inter.detach().cpu().numpy(), o_inter, rtol=1e-5)


@pytest.mark.parametrize("name", ["pk8x16", "pk16x4", "ragged"])
def test_range_loss_matches_reference(name):
    from demo2_b200.range_loss import RangeLoss
    spec, g = BATCH_LOSS_CASES[name], load_golden("batch_loss_" + name)
    feats, targets = batch_loss_case(name)
    x = feats.cuda().requires_grad_(True)
    rl, intra, inter = RangeLoss(k=spec["k"], margin=spec["range_margin"], **_kw(spec))(x, targets.cuda())
    rl.backward()
    np.testing.assert_allclose(float(intra.detach()), float(g["rl_intra"]), rtol=1e-5)
    np.testing.assert_allclose(float(inter.detach()), float(g["rl_inter"]), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(rl.detach()), float(g["rl_loss"]), rtol=1e-5)
    gx = x.grad.cpu().numpy()
    np.testing.assert_allclose(gx.ravel()[::41], g["rl_gx"], rtol=2e-5, atol=1e-7)
    assert abs(float(np.abs(gx.astype(np.float64)).sum()) - float(g["rl_gx_abs"])) < 1e-5 * float(g["rl_gx_abs"])
    o_rl, o_intra, o_inter = oracle.range_loss(feats.numpy(), targets.numpy(), k=spec["k"],
                                               margin=spec["range_margin"], **_kw(spec))
    np.testing.assert_allclose(float(rl.detach()), float(o_rl), rtol=1e-5)


def test_batch_losses_cpu_inputs_and_errors():
    """CPU tensors are moved to the device (the reference's use_gpu=True does the same, range_loss.py:210-212);
    a single identity has no inter-class term (the reference raises on the empty matrix)."""
    from demo2_b200.cluster_loss import ClusterLoss
    from demo2_b200.range_loss import RangeLoss
    feats, targets = batch_loss_case("pk8x16")
    a = ClusterLoss(ids_per_batch=8, imgs_per_id=16)(feats, targets)[0]
    b = ClusterLoss(ids_per_batch=8, imgs_per_id=16)(feats.cuda(), targets.cuda())[0]
    assert a.is_cuda and float(a.detach()) == float(b.detach())
    one = torch.zeros(16, dtype=torch.long)
    with pytest.raises(RuntimeError):
        ClusterLoss(ordered=False)(feats[:16], one)
    with pytest.raises(IndexError):
        RangeLoss(ordered=False)(feats[:16], one)
    with pytest.raises(AssertionError):
        RangeLoss()(feats, targets[:-1])


def test_range_loss_few_pairs_uses_clamp_floor():
    """k larger than the number of unordered pairs of an identity: the reference's slice reaches the
    self-distances (1e-6 = sqrt of the 1e-12 clamp); the drop-in substitutes exactly that value."""
    from demo2_b200.range_loss import RangeLoss
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(4, 64, generator=g)
    targets = torch.tensor([0, 0, 1, 1])
    _, intra, _ = RangeLoss(k=2, ordered=False)(feats, targets)
    d01 = float((feats[0] - feats[1]).norm())
    d23 = float((feats[2] - feats[3]).norm())
    want = sum(2.0 / (1.0 / d + 1.0 / 1e-6) for d in (d01, d23))
    np.testing.assert_allclose(float(intra), want, rtol=1e-4)
