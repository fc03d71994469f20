"""CPU oracle for the pytorchltr loss / metric hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/ltr_oracle.c`` (a plain-C restatement of the
reference's algorithm, each function citing the reference file:line it follows).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; nothing under
``pytorchltr_b200/`` does, and the product path never falls back to it.

Parity status: pinned against the reference's known-answer tests and against
fixtures generated from the unmodified reference (``tests/golden``), except
``listnet`` which has no reference counterpart ("parity unpinned").

All entry points take numpy arrays (``scores`` float32 ``(B, L)``, ``relevance``
int64 ``(B, L)``, ``n`` int64 ``(B,)``) and return float64 numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libltr_oracle.so")

ADDITIVE_MODES = {"hinge": 0, "dcg_hinge": 1, "logistic": 2}
LAMBDA_MODES = {"arp1": 0, "arp2": 1, "ndcg1": 2, "ndcg2": 3}

_lib = None


def build(force: bool = False) -> str:
    """Compiles ``ltr_oracle.c`` with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "ltr_oracle.c")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libltr_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _i64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64))


def _prep(scores, relevance, n):
    s = _f32(scores)
    y = _i64(relevance)
    if s.ndim == 3:
        s = s.reshape(s.shape[0], s.shape[1])
    if y.ndim == 3:
        y = y.reshape(y.shape[0], y.shape[1])
    nn = _i64(n)
    assert s.ndim == 2 and s.shape == y.shape and nn.shape == (s.shape[0],)
    return s, y, nn


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def set_threads(t: int):
    lib().ltr_oracle_set_threads(ctypes.c_int(int(t)))


def max_threads() -> int:
    return int(lib().ltr_oracle_max_threads())


def pairwise_additive(mode, scores, relevance, n, sigma=1.0, f32=False, want_grad=True):
    """-> (loss (B,), grad (B, L) | None); reference: loss/pairwise_additive.py:51-163."""
    s, y, nn = _prep(scores, relevance, n)
    B, L = s.shape
    loss = np.zeros(B, dtype=np.float64)
    grad = np.zeros((B, L), dtype=np.float64) if want_grad else None
    rc = lib().ltr_oracle_pairwise_additive(
        ctypes.c_int(ADDITIVE_MODES[mode]), ctypes.c_int(int(f32)), _p(s), _p(y), _p(nn),
        ctypes.c_int(B), ctypes.c_int(L), ctypes.c_double(sigma), _p(loss), _p(grad))
    assert rc == 0
    return loss, grad


def lambda_loss(mode, scores, relevance, n, sigma=1.0, want_grad=True, want_ranking=False):
    """-> (loss, grad | None[, ranking]); reference: loss/pairwise_lambda.py:50-241."""
    s, y, nn = _prep(scores, relevance, n)
    B, L = s.shape
    loss = np.zeros(B, dtype=np.float64)
    grad = np.zeros((B, L), dtype=np.float64) if want_grad else None
    ranking = np.zeros((B, L), dtype=np.int64) if want_ranking else None
    rc = lib().ltr_oracle_lambda(
        ctypes.c_int(LAMBDA_MODES[mode]), ctypes.c_int(0), _p(s), _p(y), _p(nn),
        ctypes.c_int(B), ctypes.c_int(L), ctypes.c_double(sigma), _p(loss), _p(grad),
        _p(ranking))
    assert rc == 0
    if want_ranking:
        return loss, grad, ranking
    return loss, grad


def listnet(scores, relevance, n, want_grad=True):
    """-> (loss, grad | None).  No reference counterpart: parity unpinned."""
    s, y, nn = _prep(scores, relevance, n)
    B, L = s.shape
    loss = np.zeros(B, dtype=np.float64)
    grad = np.zeros((B, L), dtype=np.float64) if want_grad else None
    rc = lib().ltr_oracle_listnet(_p(s), _p(y), _p(nn), ctypes.c_int(B), ctypes.c_int(L),
                                  _p(loss), _p(grad))
    assert rc == 0
    return loss, grad


def rank_by_score(scores, n):
    """-> ranking (B, L) int64; reference: utils/tensor_operations.py:48-64."""
    s = _f32(scores)
    if s.ndim == 3:
        s = s.reshape(s.shape[0], s.shape[1])
    nn = _i64(n)
    B, L = s.shape
    ranking = np.zeros((B, L), dtype=np.int64)
    rc = lib().ltr_oracle_rank_by_score(_p(s), _p(nn), ctypes.c_int(B), ctypes.c_int(L),
                                        _p(ranking))
    assert rc == 0
    return ranking


def dcg(scores, relevance, n, k=None, exp=True, normalized=False):
    """-> (B,) if k else (B, L); reference: evaluation/dcg.py:41-99 (ndcg: :8-38)."""
    s, y, nn = _prep(scores, relevance, n)
    B, L = s.shape
    if k is not None and k <= 0:
        raise IndexError("k must be positive")
    out = np.zeros(B if k is not None else (B, L), dtype=np.float64)
    rc = lib().ltr_oracle_dcg(_p(s), _p(y), _p(nn), ctypes.c_int(B), ctypes.c_int(L),
                              ctypes.c_int(0 if k is None else int(k)), ctypes.c_int(int(exp)),
                              ctypes.c_int(int(normalized)), _p(out))
    assert rc == 0
    return out


def ndcg(scores, relevance, n, k=None, exp=True):
    return dcg(scores, relevance, n, k=k, exp=exp, normalized=True)


def arp(scores, relevance, n):
    """-> (B,); reference: evaluation/arp.py:7-42."""
    s, y, nn = _prep(scores, relevance, n)
    B, L = s.shape
    out = np.zeros(B, dtype=np.float64)
    rc = lib().ltr_oracle_arp(_p(s), _p(y), _p(nn), ctypes.c_int(B), ctypes.c_int(L), _p(out))
    assert rc == 0
    return out


def linear_listnet(features, weight, bias, relevance, n):
    """Linear scorer + ListNet + gradients in float64 (numpy); test infrastructure like the rest
    of this package.  Scorer: ``torch.nn.Linear(F, 1)`` as used by the reference's training loop
    (examples/01-basic-usage.py:44,72); loss: the ListNet statement of SURVEY.md 8(a) A19
    (parity unpinned, no reference file).  Gradients are those of ``loss.sum()``.
    -> (scores (B, L), loss (B,), dscores (B, L), dweight (F,), dbias, gscale (F,)) where
    ``gscale[f] = sum |dscores * features[..., f]|`` is the magnitude the float32 accumulation
    error of dweight scales with."""
    X = np.asarray(features, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64).reshape(-1)
    b0 = 0.0 if bias is None else float(np.asarray(bias, dtype=np.float64).reshape(-1)[0])
    y = np.asarray(relevance, dtype=np.float64)
    nn = np.asarray(n, dtype=np.int64)
    B, L, F = X.shape
    s = X @ w + b0
    valid = np.arange(L)[None, :] < nn[:, None]
    loss = np.zeros(B)
    d = np.zeros((B, L))
    for b in range(B):
        if nn[b] <= 0:
            continue
        v = valid[b]
        sb, yb = s[b, v], y[b, v]
        q = sb - sb.max()
        logz = np.log(np.exp(q).sum())
        p = np.exp(yb - yb.max())
        p /= p.sum()
        loss[b] = -(p * (q - logz)).sum()
        d[b, v] = np.exp(q - logz) - p
    dweight = np.einsum("bl,blf->f", d, X)
    gscale = np.einsum("bl,blf->f", np.abs(d), np.abs(X))
    return s, loss, d, dweight, d.sum(), gscale


def collate(items, indices, max_list_size=None):
    """CPU restatement of SVMRankDataset.collate_fn(ListSampler(max_list_size)) for dense items
    (datasets/svmrank/svmrank.py:135-205, datasets/list_sampler.py:5-19): test infrastructure.
    `items[i]` = (features (n_i, F), relevance (n_i,), qid).  -> (features (B, L, F) f32, relevance
    (B, L) i64, n (B,) i64, qid (B,) i64)."""
    batch = [items[i] for i in indices]
    sizes = [b[1].shape[0] if max_list_size is None else min(max_list_size, b[1].shape[0]) for b in batch]
    L = max(sizes)                                                     # svmrank.py:142-143
    F = batch[0][0].shape[1]
    feats = np.zeros((len(batch), L, F), dtype=np.float32)             # :149-150
    rel = np.zeros((len(batch), L), dtype=np.int64)
    n = np.zeros(len(batch), dtype=np.int64)
    qid = np.zeros(len(batch), dtype=np.int64)
    for b, (x, y, q) in enumerate(batch):
        k = min(x.shape[0], L)                                         # ListSampler: arange(list_size), :161,176
        feats[b, :k] = x[:k]
        rel[b, :k] = y[:k]
        n[b] = k                                                       # min(sample.n, list_size), :194
        qid[b] = q
    return feats, rel, n, qid


def pbm_probabilities(rankings, ys, n, relevance_probs, cutoff=None, eta=1.0):
    """CPU restatement of the deterministic part of simulate_pbm (click_simulation/pbm.py:32-53,
    returned in document order like :55-62): -> (click probability, propensity), (B, L) float64."""
    rk = np.asarray(rankings, dtype=np.int64)
    y = np.asarray(ys, dtype=np.int64)
    nn = np.asarray(n, dtype=np.int64).copy()
    rp = np.asarray(relevance_probs, dtype=np.float64)
    B, L = rk.shape
    if cutoff is not None:
        nn = np.minimum(nn, cutoff)                                       # :33-34
    obs = 1.0 / (1.0 + (1.0 + np.arange(L))) ** eta                      # :37-39
    obs = np.where(np.arange(L)[None, :] < nn[:, None], obs[None, :], 0.0)   # :40-42
    cp = np.zeros((B, L))
    pr = np.zeros((B, L))
    for b in range(B):
        ranked_y = y[b, rk[b]]                                            # :45
        pr[b, rk[b]] = obs[b]                                             # :58-62 (inverse permutation)
        cp[b, rk[b]] = rp[ranked_y] * obs[b]                              # :48-53
    return cp, pr


# ---- MLP ranker (docs/source/getting-started.rst:42-51: l1 -> relu -> l2 -> relu -> l3), float64 numpy ----
def tf32_trunc(a):
    """float32 values with the low 13 mantissa bits cleared: what a kind::tf32 tensor-core operand keeps."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return (a.view(np.int32) & np.int32(~0x1FFF)).view(np.float32)


def mlp_scores(features, w1, b1, w2, b2, w3, b3, tf32=False):
    """Scores of the documented model on ``features (rows, F)``: float64 arithmetic on the float32 inputs.
    ``tf32=True`` restates the kernel's operand precision: layer 1 multiplies TF32-truncated features and
    weights (everything else exact)."""
    x = np.asarray(features, dtype=np.float32).reshape(-1, np.shape(features)[-1])
    xa, w1a = (tf32_trunc(x), tf32_trunc(w1)) if tf32 else (x, np.asarray(w1, dtype=np.float32))
    d = np.float64
    z1 = xa.astype(d) @ w1a.astype(d).T + np.asarray(b1, d)
    h1 = np.maximum(z1, 0.0)
    z2 = h1 @ np.asarray(w2, d).T + np.asarray(b2, d)
    h2 = np.maximum(z2, 0.0)
    return (h2 @ np.asarray(w3, d).reshape(-1, 1)).reshape(-1) + float(np.asarray(b3, d).reshape(-1)[0])


def mlp_grads(features, w1, b1, w2, b2, w3, b3, dscores, tf32=False):
    """Parameter gradients ``(dW1, db1, dW2, db2, dW3, db3)`` for an upstream ``dscores (rows,)``.
    ``tf32=True`` restates the backward kernel's operand precision: every tensor-core product multiplies
    TF32-truncated operands (X and W1; H1 and W2 for the layer-2 pre-activation that decides the ReLU mask;
    dZ2 and W2 for dH1; dZ1 and X for dW1; dZ1 for db1), the remaining sums are exact."""
    d = np.float64
    x = np.asarray(features, dtype=np.float32).reshape(-1, np.shape(features)[-1])
    t = tf32_trunc if tf32 else (lambda a: np.asarray(a, dtype=np.float32))
    w1f, w2f, w3f = (np.asarray(a, dtype=np.float32) for a in (w1, w2, w3))
    g = np.asarray(dscores, d).reshape(-1, 1)
    z1 = t(x).astype(d) @ t(w1f).astype(d).T + np.asarray(b1, d)
    h1 = np.maximum(z1, 0.0)
    h1_op = t(h1.astype(np.float32)).astype(d) if tf32 else h1
    z2 = h1_op @ t(w2f).astype(d).T + np.asarray(b2, d)
    h2 = np.maximum(z2, 0.0)
    dw3 = (g * h2).sum(0).reshape(1, -1)
    db3 = g.sum().reshape(1)
    dz2 = g * w3f.astype(d).reshape(1, -1) * (z2 > 0)
    dw2 = dz2.T @ h1
    db2 = dz2.sum(0)
    dz2_op = t(dz2.astype(np.float32)).astype(d) if tf32 else dz2
    dz1 = (dz2_op @ t(w2f).astype(d)) * (z1 > 0)
    dz1_op = t(dz1.astype(np.float32)).astype(d) if tf32 else dz1
    db1 = dz1_op.sum(0)
    dw1 = dz1_op.T @ t(x).astype(d)
    return dw1, db1, dw2, db2, dw3, db3


def mlp_grads_kept(features, w1, b1, w2, b2, w3, b3, dscores):
    """Restatement of ``ltr_mlp_backward`` started from the activations the forward kernel kept (``hz``): layer 1
    with TF32 operands, layer 2 exact (the forward kernel runs it in float32), the forward pass's own ReLU masks,
    TF32 operands in the three tensor-core products dH1 = dZ2 W2, [dW1, db1] = dZ1^T [X, 1], [dW2, db2] = dZ2^T [H1, 1]."""
    d = np.float64
    t = tf32_trunc
    x = np.asarray(features, dtype=np.float32).reshape(-1, np.shape(features)[-1])
    w1f, w2f, w3f = (np.asarray(a, dtype=np.float32) for a in (w1, w2, w3))
    g = np.asarray(dscores, d).reshape(-1, 1)
    z1 = t(x).astype(d) @ t(w1f).astype(d).T + np.asarray(b1, d)
    h1 = np.maximum(z1, 0.0)
    z2 = h1 @ w2f.astype(d).T + np.asarray(b2, d)
    h2 = np.maximum(z2, 0.0)
    dz2 = g * w3f.astype(d).reshape(1, -1) * (z2 > 0)
    dz2_op = t(dz2.astype(np.float32)).astype(d)
    dz1 = (dz2_op @ t(w2f).astype(d)) * (z1 > 0)
    dz1_op = t(dz1.astype(np.float32)).astype(d)
    return (dz1_op.T @ t(x).astype(d), dz1_op.sum(0), dz2_op.T @ t(h1.astype(np.float32)).astype(d), dz2_op.sum(0),
            (g * h2).sum(0).reshape(1, -1), g.sum().reshape(1))
