"""Size-independent properties of the CPU oracle (the checker the GPU parity tests rely on): what the
domain guarantees for every loss of the path must hold for the restatement too.

  * translation invariance: every pairwise / lambda loss only sees score differences, so the
    per-query gradient sums to zero;
  * permutation equivariance: shuffling the documents of a query shuffles the gradient and leaves
    the loss unchanged (tie-free scores);
  * padding: scores and relevance beyond n[b] never matter;
  * metrics: ndcg in [0, 1], equal to 1 when the scores order the documents by relevance; arp of a
    single relevant document is its rank.
"""
import numpy as np
import pytest

import oracle

ADDITIVE = ("hinge", "dcg_hinge", "logistic")
LAMBDA = ("arp1", "arp2", "ndcg1", "ndcg2")


def _loss(mode, s, y, n):
    if mode in ADDITIVE:
        return oracle.pairwise_additive(mode, s, y, n)
    return oracle.lambda_loss(mode, s, y, n)


def _batch(seed, B=5, L=23):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((B, L)).astype(np.float32)
    y = rng.integers(0, 5, size=(B, L))
    n = rng.integers(2, L + 1, size=B)
    n[0] = L
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    return s, y, n


@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
def test_gradient_sums_to_zero_and_padding_is_ignored(mode):
    s, y, n = _batch(1)
    loss, grad = _loss(mode, s, y, n)
    scale = np.abs(grad).max(axis=1) + 1e-12
    assert np.all(np.abs(grad.sum(axis=1)) <= 1e-9 * scale * s.shape[1])
    pad = np.arange(s.shape[1])[None, :] >= n[:, None]
    assert np.all(grad[pad] == 0.0)
    s2, y2 = s.copy(), y.copy()
    s2[pad] = 1e6
    y2[pad] = 4
    loss2, grad2 = _loss(mode, s2, y2, n)
    assert np.array_equal(loss, loss2) and np.array_equal(grad, grad2)


@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
def test_permutation_equivariance(mode):
    s, y, n = _batch(2)
    loss, grad = _loss(mode, s, y, n)
    rng = np.random.default_rng(3)
    s2, y2 = s.copy(), y.copy()
    perms = []
    for b in range(s.shape[0]):
        p = rng.permutation(n[b])
        perms.append(p)
        s2[b, : n[b]] = s[b, p]
        y2[b, : n[b]] = y[b, p]
    loss2, grad2 = _loss(mode, s2, y2, n)
    assert np.allclose(loss, loss2, rtol=1e-9, atol=1e-12)
    for b, p in enumerate(perms):
        assert np.allclose(grad2[b, : n[b]], grad[b, p], rtol=1e-8, atol=1e-12)


def test_metric_properties():
    s, y, n = _batch(4)
    v = oracle.ndcg(s, y, n, k=10)
    assert np.all((v >= 0) & (v <= 1 + 1e-12))
    perfect = y.astype(np.float32) + 0.001 * np.arange(y.shape[1], 0, -1, dtype=np.float32)[None, :]
    assert np.allclose(oracle.ndcg(perfect, y, n, k=10), np.where(y.max(axis=1) > 0, 1.0, 0.0))
    one = np.zeros_like(y)
    pos = np.array([int(k) // 2 for k in n])
    one[np.arange(len(n)), pos] = 1
    ranks = np.array([1 + (s[b, : n[b]] > s[b, pos[b]]).sum() for b in range(len(n))], dtype=np.float64)
    assert np.allclose(oracle.arp(s, one, n), ranks)
