"""The float64 oracle of the fused linear scorer + ListNet against torch autograd (float64) of the
caller-side composite: torch.nn.functional.linear followed by the masked softmax cross entropy of
SURVEY.md 8(a) A19 (LogSoftmax over mask_padded_values(scores, n), utils/tensor_operations.py:81-87).
ListNet has no reference file: parity of the loss itself is unpinned; this pins the plumbing
(masking, gradients through the scorer) of the oracle."""
import numpy as np
import torch

import oracle


def _torch_composite(X, w, b, y, n):
    X = torch.as_tensor(X, dtype=torch.float64)
    w = torch.as_tensor(w, dtype=torch.float64).requires_grad_(True)
    b = torch.as_tensor(b, dtype=torch.float64).requires_grad_(True)
    y = torch.as_tensor(y, dtype=torch.float64)
    n = torch.as_tensor(n)
    B, L, F = X.shape
    s = torch.nn.functional.linear(X, w.reshape(1, F), b).reshape(B, L)
    s.retain_grad()
    mask = torch.arange(L)[None, :] < n[:, None]
    neg = torch.full_like(s, float("-inf"))
    q = torch.log_softmax(torch.where(mask, s, neg), dim=1)
    p = torch.softmax(torch.where(mask, y, neg), dim=1)
    per_doc = torch.where(mask, p * q, torch.zeros_like(s))
    loss = -per_doc.sum(dim=1)
    loss = torch.where(n > 0, loss, torch.zeros_like(loss))
    loss.sum().backward()
    return s.detach().numpy(), loss.detach().numpy(), s.grad.numpy(), w.grad.numpy(), float(b.grad)


def test_oracle_linear_listnet_matches_torch_float64_autograd():
    rng = np.random.default_rng(0)
    B, L, F = 9, 13, 8
    X = rng.standard_normal((B, L, F))
    w = rng.standard_normal(F) * 0.5
    b = rng.standard_normal(1)
    y = rng.integers(0, 5, size=(B, L))
    n = rng.integers(1, L + 1, size=B)
    n[0] = L
    n[1] = 1
    s, loss, d, dw, db, gscale = oracle.linear_listnet(X, w, b, y, n)
    ts, tl, td, tdw, tdb = _torch_composite(X, w, b, y, n)
    assert np.allclose(s, ts, rtol=1e-12, atol=1e-12)
    assert np.allclose(loss, tl, rtol=1e-10, atol=1e-12)
    assert np.allclose(d, np.nan_to_num(td), rtol=1e-10, atol=1e-12)
    assert np.allclose(dw, tdw, rtol=1e-9, atol=1e-11)
    assert abs(db - tdb) < 1e-10
    assert np.all(gscale >= np.abs(dw) - 1e-12)


def test_oracle_linear_listnet_agrees_with_c_oracle_listnet():
    """Same loss as the C restatement of ListNet fed with the float32-rounded scores."""
    rng = np.random.default_rng(1)
    B, L, F = 6, 40, 12
    X = rng.standard_normal((B, L, F)).astype(np.float32)
    w = (rng.standard_normal(F) * 0.3).astype(np.float32)
    y = rng.integers(0, 5, size=(B, L))
    n = rng.integers(0, L + 1, size=B)
    s, loss, d, _, _, _ = oracle.linear_listnet(X, w, None, y, n)
    y0 = y.copy()
    y0[np.arange(L)[None, :] >= n[:, None]] = 0
    closs, cgrad = oracle.listnet(s.astype(np.float32), y0, n)
    assert np.allclose(loss, closs, rtol=1e-5, atol=1e-6)
    assert np.allclose(d, cgrad, rtol=1e-4, atol=1e-6)
