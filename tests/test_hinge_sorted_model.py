"""The algorithm of hinge_sorted_kernel (csrc/ltr_hinge_sorted.cuh) restated in numpy and checked on
the CPU against the O(n^2) float32 restatement of the reference (oracle, loss/pairwise_additive.py:
107-113): sorting by score, the exact float32 kink predicate fl(s_i - s_j) <= 1 located by binary
search, and per-class-boundary prefix counts / score sums must reproduce the integer-valued gradients
bit for bit -- including pairs exactly on the kink, tied scores and negative / sparse grades.  (The GPU
parity tests check the CUDA code; this pins the algorithm itself, on every CPU run.)"""
import numpy as np
import pytest

import oracle


def hinge_sorted_model(s, y, n):
    s = np.asarray(s, dtype=np.float32)
    y = np.asarray(y, dtype=np.int64)
    B, L = s.shape
    loss = np.zeros(B)
    grad = np.zeros((B, L))
    one = np.float32(1.0)
    for b in range(B):
        nb = int(n[b])
        if nb < 2:
            continue
        order = np.argsort(s[b, :nb], kind="stable")          # ascending score
        ss, yy = s[b, order], y[b, order]
        classes = {g: k for k, g in enumerate(sorted(set(yy.tolist())))}
        cl = np.array([classes[g] for g in yy])
        lo = np.zeros(nb, dtype=np.int64)
        hi = np.zeros(nb, dtype=np.int64)
        for p in range(nb):
            a, c = 0, nb                                         # first q with fl(s_p - s_q) <= 1
            while a < c:
                mid = (a + c) // 2
                if np.float32(ss[p] - ss[mid]) <= one:
                    c = mid
                else:
                    a = mid + 1
            lo[p] = a
            a, c = 0, nb                                         # last q with fl(s_q - s_p) <= 1
            while a < c:
                mid = (a + c) // 2
                if np.float32(ss[mid] - ss[p]) <= one:
                    a = mid + 1
                else:
                    c = mid
            hi[p] = a - 1
        g = np.zeros(nb)
        for c in range(1, len(classes)):
            below = cl < c
            A = np.concatenate([[0], np.cumsum(below)])
            S = np.concatenate([[0.0], np.cumsum(np.where(below, ss.astype(np.float64), 0.0))])
            for p in range(nb):
                if cl[p] == c:
                    cnt = A[nb] - A[lo[p]]
                    loss[b] += cnt * (1.0 - float(ss[p])) + (S[nb] - S[lo[p]])
                    g[p] -= cnt
                elif cl[p] == c - 1:
                    g[p] += (hi[p] + 1) - A[hi[p] + 1]
        grad[b, order] = g
    return loss, grad


@pytest.mark.parametrize("seed,B,L,grid", [(0, 6, 40, None), (1, 5, 130, None), (2, 6, 64, 0.25), (3, 4, 90, 0.5),
                                           (4, 3, 257, 0.125)])
def test_sorted_hinge_algorithm_equals_pairwise_reference(seed, B, L, grid):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((B, L)).astype(np.float32)
    if grid:
        s = (np.round(s / grid) * grid).astype(np.float32)      # many pairs exactly on the kink, many ties
    y = rng.integers(0, 5, size=(B, L))
    y[0] = rng.integers(-3, 50, size=L)                          # sparse / negative grades
    n = rng.integers(L // 2, L + 1, size=B)
    n[-1] = L
    if B > 2:
        n[1] = 1
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    ref_loss, ref_grad = oracle.pairwise_additive("hinge", s, y, n, f32=True)
    loss, grad = hinge_sorted_model(s, y, n)
    assert np.array_equal(grad, ref_grad)
    assert np.allclose(loss, ref_loss, rtol=1e-5, atol=1e-6)
