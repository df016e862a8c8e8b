"""Host logic of the ClusterLoss / RangeLoss drop-ins on CPU: the modules run unchanged with the
library distance call replaced by the same formula in plain torch (the CUDA path itself is covered
by tests/test_gpu_batch_losses.py), against the vectors minted from the reference."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import BATCH_LOSS_CASES, batch_loss_case, load_golden


def _torch_euclidean_dist(x, y):
    xx = (x * x).sum(1, keepdim=True)
    yy = (y * y).sum(1, keepdim=True).t()
    return (xx + yy - 2 * x @ y.t()).clamp(min=1e-12).sqrt()


@pytest.fixture
def cpu_losses(monkeypatch):
    import demo2_b200.cluster_loss as CL
    import demo2_b200.range_loss as RL
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(CL, "euclidean_dist", _torch_euclidean_dist)
    monkeypatch.setattr(RL, "euclidean_dist", _torch_euclidean_dist)
    return CL, RL


@pytest.mark.parametrize("name", ["pk8x16", "pk16x4", "ragged"])
def test_batch_loss_host_logic_matches_reference(cpu_losses, name):
    CL, RL = cpu_losses
    spec, g = BATCH_LOSS_CASES[name], load_golden("batch_loss_" + name)
    feats, targets = batch_loss_case(name)
    kw = dict(ordered=spec["ordered"], ids_per_batch=spec["P"], imgs_per_id=spec["K"])
    x = feats.clone().requires_grad_(True)
    loss, intra, inter = CL.ClusterLoss(margin=spec["cluster_margin"], **kw)(x, targets)
    loss.backward()
    np.testing.assert_allclose(intra.detach().numpy(), g["cl_intra"], rtol=1e-5)
    np.testing.assert_allclose(inter.detach().numpy(), g["cl_inter"], rtol=1e-5)
    np.testing.assert_allclose(float(loss.detach()), float(g["cl_loss"]), rtol=1e-5)
    np.testing.assert_allclose(x.grad.numpy().ravel()[::41], g["cl_gx"], rtol=2e-5, atol=1e-7)
    x = feats.clone().requires_grad_(True)
    rl, r_intra, r_inter = RL.RangeLoss(k=spec["k"], margin=spec["range_margin"], **kw)(x, targets)
    rl.backward()
    np.testing.assert_allclose(float(r_intra.detach()), float(g["rl_intra"]), rtol=1e-5)
    np.testing.assert_allclose(float(r_inter.detach()), float(g["rl_inter"]), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(rl.detach()), float(g["rl_loss"]), rtol=1e-5)
    np.testing.assert_allclose(x.grad.numpy().ravel()[::41], g["rl_gx"], rtol=2e-5, atol=1e-7)


def test_batch_identities_order(cpu_losses):
    """P x K ordered batch: identities in order of appearance (cluster_loss.py:46-47); otherwise sorted unique."""
    CL, _ = cpu_losses
    t = torch.tensor([7, 7, 3, 3, 9, 9])
    assert CL.batch_identities(t, True, 3, 2).tolist() == [7, 3, 9]
    assert CL.batch_identities(t, True, 4, 2).tolist() == [3, 7, 9]
    assert CL.batch_identities(t, False, 3, 2).tolist() == [3, 7, 9]
