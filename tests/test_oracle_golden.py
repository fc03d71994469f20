"""Pins the CPU oracle against fixtures generated from the unmodified reference.

tests/golden/ref_*.npz are produced by tests/golden/make_golden.py, which imports the
reference (float64 scores, CPU autograd).  Tolerances: the oracle does the pair math
in double like "ref64", but the reference's score-independent weights are float32
(max-DCG summed in float32 in ATen's order), so losses/gradients of the NDCG losses
agree only to a few float32 ulps of the weights: 2e-6 relative.  Everything else
agrees to 1e-9.  Rankings are integer: bit-exact on the valid prefix.
"""
import glob
import os

import numpy as np
import pytest
from pytest import approx

import oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))
ADDITIVE = ("hinge", "dcg_hinge", "logistic")
LAMBDA = ("arp1", "arp2", "ndcg1", "ndcg2")


def _load(path):
    return dict(np.load(path))


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
def test_oracle_loss_vs_reference(path, mode):
    g = _load(path)
    sigma = float(g["sigma"])
    if mode in ADDITIVE:
        loss, grad = oracle.pairwise_additive(mode, g["scores"], g["relevance"], g["n"], sigma=sigma)
    else:
        loss, grad = oracle.lambda_loss(mode, g["scores"], g["relevance"], g["n"], sigma=sigma)
    rel = 2e-6 if mode.startswith("ndcg") else 1e-9
    ref_loss, ref_grad = g[f"{mode}_loss64"], g[f"{mode}_grad64"]
    assert loss == approx(ref_loss, rel=rel, abs=1e-9)
    gmax = np.abs(ref_grad).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(grad - ref_grad) <= rel * gmax + 1e-12)


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_oracle_hinge_f32_mode_grad_is_exact(path):
    """ref32 hinge: integer-valued gradient, bit-exact against the reference's float32 run."""
    g = _load(path)
    loss, grad = oracle.pairwise_additive("hinge", g["scores"], g["relevance"], g["n"], f32=True)
    assert np.array_equal(grad, g["hinge_grad32"].astype(np.float64))
    assert loss == approx(g["hinge_loss32"], rel=1e-5, abs=1e-6)


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_oracle_ranking_vs_reference(path):
    g = _load(path)
    rk = oracle.rank_by_score(g["scores"], g["n"])
    L = g["scores"].shape[1]
    for b, nb in enumerate(g["n"]):
        assert np.array_equal(rk[b, :nb], g["ranking"][b, :nb])
        assert sorted(rk[b, nb:]) == list(range(nb, L))
        assert sorted(g["ranking"][b, nb:]) == list(range(nb, L))


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
@pytest.mark.parametrize("exp", [True, False])
def test_oracle_metrics_vs_reference(path, exp):
    g = _load(path)
    tag = "exp" if exp else "lin"
    s, y, n = g["scores"], g["relevance"], g["n"]
    assert oracle.dcg(s, y, n, exp=exp) == approx(g[f"dcg_all_{tag}"], rel=2e-6, abs=1e-6)
    assert oracle.ndcg(s, y, n, exp=exp) == approx(g[f"ndcg_all_{tag}"], rel=2e-6, abs=1e-6)
    for k in (1, 3, 10, 1000):
        assert oracle.dcg(s, y, n, k=k, exp=exp) == approx(g[f"dcg_k{k}_{tag}"], rel=2e-6, abs=1e-6)
        assert oracle.ndcg(s, y, n, k=k, exp=exp) == approx(g[f"ndcg_k{k}_{tag}"], rel=2e-6, abs=1e-6)
    assert oracle.arp(s, y, n) == approx(g["arp"], rel=2e-6, abs=1e-6)


def test_listnet_oracle_is_self_consistent():
    """ListNet is unpinned (absent from the reference): check the closed-form gradient
    of the oracle against central differences of its own loss, and basic identities."""
    rng = np.random.default_rng(7)
    B, L = 5, 19
    s = rng.standard_normal((B, L)).astype(np.float32)
    y = rng.integers(0, 5, (B, L)).astype(np.int64)
    n = np.asarray([19, 10, 1, 0, 7], dtype=np.int64)
    loss, grad = oracle.listnet(s, y, n)
    assert loss[3] == 0.0 and np.all(grad[3] == 0) and loss[2] == approx(0.0, abs=1e-12)
    assert np.all(grad[np.arange(L)[None, :] >= n[:, None]] == 0)
    assert grad.sum(axis=1) == approx(np.zeros(B), abs=1e-12)
    eps = 1e-3
    for b, j in ((0, 3), (1, 9), (4, 0)):
        sp, sm = s.copy(), s.copy()
        sp[b, j] += eps
        sm[b, j] -= eps
        num = (oracle.listnet(sp, y, n)[0][b] - oracle.listnet(sm, y, n)[0][b]) / (
            float(sp[b, j]) - float(sm[b, j]))
        assert grad[b, j] == approx(num, rel=1e-4, abs=1e-6)
