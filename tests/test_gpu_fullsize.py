"""Full-size (BASELINE.json configs[3]-scale) GPU checks through size-independent properties:
shard additivity of the rank counts, agreement of sampled queries with the CPU oracle run on
our own distances, and the all-device multi-"rank" data flow executed on one GPU."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import make_case, oracle

pytestmark = pytest.mark.gpu


def _synth(Q, G, d, nid, ncam, seed=0):
    rng = np.random.default_rng(seed)
    qp, gp = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    qc, gc = rng.integers(0, ncam, Q), rng.integers(0, ncam, G)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    centers = torch.randn(nid, d, device="cuda", generator=gen)
    qf = centers[torch.from_numpy(qp).cuda()] + 4 * torch.randn(Q, d, device="cuda", generator=gen)
    gf = torch.empty(G, d, device="cuda")
    gpd = torch.from_numpy(gp).cuda()
    for s in range(0, G, 131072):
        e = min(G, s + 131072)
        gf[s:e] = centers[gpd[s:e]] + 4 * torch.randn(e - s, d, device="cuda", generator=gen)
    return qf, gf, qp, gp, qc, gc


def _oracle_rows(M, qf, gf, rows, qp, gp, qc, gc, normalize):
    """Oracle per-query AP / first rank for a few query rows, on OUR distances of those rows."""
    dist = M.sqdist_device(qf[rows], gf, normalize=normalize).cpu().numpy()
    ofs, idx, r, c = oracle.rank_counts(dist, qp[rows], gp, qc[rows], gc)
    ap, first = [], []
    for q in range(len(rows)):
        s, e = ofs[q], ofs[q + 1]
        ap.append((c[s:e] / r[s:e]).sum() / (e - s) if e > s else -1.0)
        first.append(r[s:e].min() if e > s else 0)
    return np.array(ap), np.array(first)


def test_chunked_equals_unsharded_small():
    from demo2_b200 import metrics, parallel
    qf, gf, qp, gp, qc, gc = make_case("rgbnt201", 2, 4.0)
    gf = gf.copy()
    gf[500:520] = gf[10:30]     # exact ties across chunk boundaries
    ref = metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
    for n in (2, 3, 7):
        res = parallel.evaluate_gallery_chunks(qf, gf, qp, gp, qc, gc, n)
        np.testing.assert_array_equal(res.first.cpu().numpy(), ref.first.cpu().numpy())
        np.testing.assert_array_equal(res.ap.cpu().numpy(), ref.ap.cpu().numpy())
        assert res.mAP == ref.mAP and res.num_valid == ref.num_valid
        np.testing.assert_array_equal(res.cmc, ref.cmc)


def test_full_size_additivity_and_sampled_oracle():
    """20 000 x 1 000 000 x 1536 (the headline workload): one-shot fused evaluation == the
    2-"rank" data flow (records all-gathered, counts added), bit for bit; 12 sampled queries
    agree exactly with the oracle ranking of our own distances."""
    from demo2_b200 import metrics, parallel
    Q, G, d = 20000, 1000000, 1536
    qf, gf, qp, gp, qc, gc = _synth(Q, G, d, 50000, 8)
    one = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, normalize=True)
    assert one.num_valid > 0.99 * Q
    assert abs(float(one.cmc[-1]) - float((one.first.cpu().numpy()[one.first.cpu().numpy() > 0] <= 50).mean())) < 1e-6
    rows = np.random.default_rng(1).choice(Q, 12, replace=False)
    ap_o, first_o = _oracle_rows(metrics, qf, gf, rows, qp, gp, qc, gc, normalize=True)
    np.testing.assert_array_equal(one.first.cpu().numpy()[rows], first_o)
    np.testing.assert_allclose(one.ap.cpu().numpy()[rows], ap_o, atol=1e-12)
    ap1, first1 = one.ap.cpu().numpy(), one.first.cpu().numpy()
    del one
    torch.cuda.empty_cache()
    two = parallel.evaluate_gallery_chunks(qf, gf, qp, gp, qc, gc, 2, normalize=True)
    np.testing.assert_array_equal(two.first.cpu().numpy(), first1)
    np.testing.assert_array_equal(two.ap.cpu().numpy(), ap1)


def test_rgbnt100_rerank_against_oracle_stages():
    """RGBNT100 scale (N = 10 290): re-ranking stages vs the oracle fed with our all-pairs matrix."""
    from demo2_b200 import metrics, reranking
    qf, gf, qp, gp, qc, gc = make_case("rgbnt100", 0, 5.0)
    feat = np.concatenate([qf, gf])
    D = metrics.sqdist_device(feat, feat).cpu().numpy()
    ours = reranking.re_ranking(qf, gf, 20, 6, 0.3)
    expect = oracle.re_ranking_from_allpairs(D.T, len(qf), 20, 6, 0.3)
    assert (ours == expect).mean() > 0.9995
    assert np.abs(ours - expect).max() <= 4 * 2.0 ** -11
    cmc, mAP = metrics.eval_func(ours, qp, gp, qc, gc)
    cmc_o, mAP_o = oracle.eval_func(expect, qp, gp, qc, gc)
    assert abs(mAP - mAP_o) < 1e-5
