"""GPU parity for the CenterLoss drop-in (layers/center_loss.py:7-47) against values minted from the
reference (tests/golden/make_golden.py::center_loss_case) and against plain torch autograd."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _inputs():
    g = torch.Generator().manual_seed(5)
    C, D, B = 50, 1536, 64
    centers = torch.randn(C, D, generator=g)
    labels = torch.arange(B) % C
    x = torch.randn(B, D, generator=g) * 0.5 + centers[labels]
    return centers, labels, x


@pytest.mark.parametrize("use_gpu", [True, False])
def test_center_loss_matches_reference(use_gpu):
    from demo2_b200.center_loss import CenterLoss
    gold = load_golden("center_loss_b64_c50")
    centers, labels, x = _inputs()
    cl = CenterLoss(num_classes=50, feat_dim=1536, use_gpu=use_gpu)
    with torch.no_grad():
        cl.centers.copy_(centers)
    xin = (x.cuda() if use_gpu else x.clone()).requires_grad_(True)
    loss = cl(xin, labels.cuda() if use_gpu else labels)
    loss.backward()
    assert abs(float(loss.detach()) - float(gold["loss"])) < 1e-5 * float(gold["loss"])       # fp32, 1e-5 relative
    np.testing.assert_allclose(xin.grad.cpu().numpy().ravel()[::37], gold["gx_sample"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cl.centers.grad.cpu().numpy().ravel()[::37], gold["gc_sample"], rtol=1e-5, atol=1e-7)
    assert abs(float(xin.grad.double().sum()) - float(gold["gx_sum"])) < 1e-4


def test_center_loss_clamp_and_duplicates():
    """A sample sitting exactly on its centre: the clamp floor 1e-12 applies and no gradient flows."""
    from demo2_b200.center_loss import CenterLoss
    cl = CenterLoss(num_classes=4, feat_dim=64, use_gpu=True)
    x = cl.centers.detach()[[1, 3, 1]].clone()
    x[2] += 0.5
    x.requires_grad_(True)
    labels = torch.tensor([1, 3, 1]).cuda()
    loss = cl(x, labels)
    loss.backward()
    ref = ((x.detach()[2] - cl.centers.detach()[1]) ** 2).sum() / 3
    assert abs(float(loss.detach()) - float(ref)) < 1e-4
    assert float(x.grad[0].abs().max()) < 1e-3 and float(x.grad[2].abs().max()) > 0.1
