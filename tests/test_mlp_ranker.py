"""N1 (SURVEY.md 8 f): the reference's documented MLP ranker (docs/source/getting-started.rst:42-51) on the
tcgen05 kernels -- ltr_mlp_scores / ltr_mlp_backward through the C ABI and pytorchltr_b200.fused.MLPRanker.

The oracle is a float64 numpy restatement (oracle.mlp_scores / oracle.mlp_grads); with ``tf32=True`` it keeps
exactly the operand bits a kind::tf32 MMA keeps, so the comparison is tight (float32 accumulation error only);
against the unrounded float64 model the tolerance is the TF32 operand precision (2^-10 relative per product).
"""
import numpy as np
import pytest
import torch

import oracle
from pytorchltr_b200 import _lib

gpu = pytest.mark.gpu


def _params(F, H1, H2, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(H1, F, generator=g) / F ** 0.5
    b1 = torch.randn(H1, generator=g) * 0.1
    w2 = torch.randn(H2, H1, generator=g) / H1 ** 0.5
    b2 = torch.randn(H2, generator=g) * 0.1
    w3 = torch.randn(1, H2, generator=g) / H2 ** 0.5
    b3 = torch.randn(1, generator=g) * 0.1
    return [t.to(device) for t in (w1, b1, w2, b2, w3, b3)]


# ---- CPU: the restatement itself ---------------------------------------------------------------------
def test_oracle_mlp_matches_torch_float64():
    F, H1, H2, rows = 20, 7, 3, 50
    p = _params(F, H1, H2, 0)
    x = torch.randn(rows, F, generator=torch.Generator().manual_seed(1))
    ds = torch.randn(rows, generator=torch.Generator().manual_seed(2))
    pd = [t.double().requires_grad_() for t in p]
    h1 = torch.relu(x.double() @ pd[0].t() + pd[1])
    h2 = torch.relu(h1 @ pd[2].t() + pd[3])
    s = (h2 @ pd[4].t() + pd[5]).reshape(-1)
    s.backward(ds.double())
    ref = oracle.mlp_scores(x.numpy(), *[t.numpy() for t in p])
    np.testing.assert_allclose(ref, s.detach().numpy(), rtol=1e-12, atol=1e-12)
    grads = oracle.mlp_grads(x.numpy(), *[t.numpy() for t in p], ds.numpy())
    for got, want in zip(grads, pd):
        np.testing.assert_allclose(got.reshape(want.shape), want.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_oracle_kept_activation_restatement_is_the_exact_gradient_on_tf32_exact_operands():
    """With every tensor-core operand exactly representable in TF32 the restatement of the kept-activation
    backward pass equals the plain float64 gradient (masks included)."""
    rng = np.random.default_rng(3)
    F, H1, H2, rows = 12, 5, 3, 64
    q = lambda a: np.round(np.asarray(a) * 8) / 8                    # few mantissa bits: products of TF32-exact values
    x = q(rng.standard_normal((rows, F))).astype(np.float32)
    p = [q(rng.standard_normal(s)).astype(np.float32) for s in ((H1, F), (H1,), (H2, H1), (H2,), (1, H2), (1,))]
    ds = q(rng.standard_normal(rows)).astype(np.float32)
    a = oracle.mlp_grads(x, *p, ds)
    b = oracle.mlp_grads_kept(x, *p, ds)
    c = oracle.mlp_grads(x, *p, ds, tf32=True)
    for u, v, w in zip(a, b, c):
        np.testing.assert_allclose(v.reshape(u.shape), u, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(w.reshape(u.shape), u, rtol=1e-12, atol=1e-12)


def test_tf32_trunc_keeps_19_bits():
    a = np.array([1.0, 1.0 + 2.0 ** -10, 1.0 + 2.0 ** -11, -3.1415927, 1e-30, 65504.0], dtype=np.float32)
    t = oracle.tf32_trunc(a)
    assert t[0] == 1.0 and t[1] == a[1] and t[2] == 1.0
    assert np.all(np.abs(t) <= np.abs(a)) and np.all(np.abs(a - t) <= np.abs(a) * 2.0 ** -10)


def test_mlp_symbols_exported():
    import ctypes
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in ("ltr_mlp_scores", "ltr_mlp_backward", "ltr_mlp_grad_len", "ltr_mlp_workspace_bytes", "ltr_mlp_hz_pitch"):
        assert hasattr(lib, sym)
    lib.ltr_mlp_grad_len.restype = ctypes.c_size_t
    lib.ltr_mlp_grad_len.argtypes = [ctypes.c_int] * 3
    assert lib.ltr_mlp_grad_len(136, 50, 10) == 50 * 136 + 50 + 10 * 50 + 10 + 10 + 1


def test_mlp_ranker_state_dict_matches_documented_model():
    from pytorchltr_b200.fused import MLPRanker

    class Model(torch.nn.Module):            # docs/source/getting-started.rst:42-51
        def __init__(self, in_features):
            super().__init__()
            self.l1 = torch.nn.Linear(in_features, 50)
            self.l2 = torch.nn.Linear(50, 10)
            self.l3 = torch.nn.Linear(10, 1)

    m, r = Model(136), MLPRanker(136)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in r.state_dict().items()]
    r.load_state_dict(m.state_dict())
    with pytest.raises(RuntimeError):        # no CPU compute path
        r(torch.zeros(2, 5, 136))


# ---- GPU ------------------------------------------------------------------------------------------------
def _call_scores(lib, x, p, hz=None):
    rows, F = x.shape
    out = torch.full((rows,), float("nan"), device=x.device)
    rc = lib.ltr_mlp_scores(x.data_ptr(), rows, F, p[0].data_ptr(), p[1].data_ptr(), p[0].shape[0], p[2].data_ptr(),
                            p[3].data_ptr(), p[2].shape[0], p[4].data_ptr(), p[5].data_ptr(), out.data_ptr(),
                            None if hz is None else hz.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return rc, out


def _kept_activations(lib, x, p):
    """forward pass keeping [H1 | Z2] (None when the shape has no such path)"""
    pitch = lib.ltr_mlp_hz_pitch(p[0].shape[0], p[2].shape[0])
    if pitch == 0:
        return None
    hz = torch.full((x.shape[0], pitch), float("nan"), device=x.device)
    rc, _ = _call_scores(lib, x, p, hz)
    assert rc == 0
    return hz


def _call_backward(lib, x, p, ds, hz=None):
    rows, F = x.shape
    H1, H2 = p[0].shape[0], p[2].shape[0]
    n = lib.ltr_mlp_grad_len(F, H1, H2)
    out = torch.full((n,), float("nan"), device=x.device)
    wsb = lib.ltr_mlp_workspace_bytes(F, H1, H2)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    rc = lib.ltr_mlp_backward(x.data_ptr(), rows, F, p[0].data_ptr(), p[1].data_ptr(), H1, p[2].data_ptr(),
                              p[3].data_ptr(), H2, p[4].data_ptr(), p[5].data_ptr(), None if hz is None else hz.data_ptr(),
                              ds.data_ptr(), out.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    return rc, out


def _split(out, F, H1, H2):
    o = np.cumsum([0, H1 * F, H1, H2 * H1, H2, H2, 1])
    shapes = [(H1, F), (H1,), (H2, H1), (H2,), (1, H2), (1,)]
    return [out[o[k]:o[k + 1]].reshape(shapes[k]) for k in range(6)]


SHAPES = [(128, 136, 50, 10), (1000, 136, 50, 10), (40000, 136, 50, 10), (5000, 128, 50, 10), (5000, 32, 50, 10),
          (3000, 24, 50, 10), (700, 8, 50, 10), (5000, 48, 64, 16), (5000, 100, 20, 5), (9000, 64, 32, 8), (1, 136, 50, 10),
          (4000, 136, 40, 10), (4000, 64, 50, 4)]
# rows wider than 288 features: W1 streams through the stage ring beside the tile (forward), the dW1 columns come
# from several launches (backward from kept activations)
WIDE_SHAPES = [(3000, 700, 50, 10), (1000, 320, 50, 10), (2100, 1024, 20, 5), (700, 292, 32, 8), (1, 700, 50, 10),
               (129, 320, 32, 8), (260, 4096, 50, 10)]
KEPT_SHAPES = ([s_ for s_ in SHAPES if s_[2] <= 50 and s_[3] <= 10] + [(3000, 220, 50, 10)]
               + WIDE_SHAPES)


@gpu
@pytest.mark.parametrize("rows,F,H1,H2", SHAPES + [(5000, 220, 20, 5)] + WIDE_SHAPES)
def test_mlp_scores_vs_oracle(rows, F, H1, H2):
    lib = _lib.lib()
    p = _params(F, H1, H2, 3, "cuda")
    x = torch.randn(rows, F, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    rc, out = _call_scores(lib, x, p)
    assert rc == 0
    torch.cuda.synchronize()
    pn = [t.cpu().numpy() for t in p]
    ref_t = oracle.mlp_scores(x.cpu().numpy(), *pn, tf32=True)
    ref = oracle.mlp_scores(x.cpu().numpy(), *pn)
    got = out.cpu().numpy().astype(np.float64)
    scale = max(1.0, np.abs(ref).max())
    # same operand bits as the tensor core: float32 accumulation error only
    assert np.abs(got - ref_t).max() <= 2e-5 * scale
    # TF32 operand precision against the exact model (2^-10 per product, 1.2e-3 observed at F = 136)
    assert np.abs(got - ref).max() <= 4e-3 * scale


def _near_kink(x, pn):
    """documents with a pre-activation next to a ReLU kink, whose mask the float64 restatement cannot reproduce:
    layer 1 within the float32 accumulation error of zero; layer 2 within 2e-3, because the kernel's H1 (float32
    sum) and the restatement's (float64 sum) fall on different sides of a TF32 truncation step for ~0.1 % of the
    units, which moves Z2 by up to one TF32 ulp of a term"""
    d = np.float64
    t = oracle.tf32_trunc
    z1 = t(x).astype(d) @ t(pn[0]).astype(d).T + pn[1].astype(d)
    h1 = np.maximum(z1, 0.0)
    z2 = t(h1.astype(np.float32)).astype(d) @ t(pn[2]).astype(d).T + pn[3].astype(d)
    return (np.abs(z2) < 2e-3).any(axis=1) | ((np.abs(z1) < 2e-5) & (np.abs(z1) > 0)).any(axis=1)


@gpu
@pytest.mark.parametrize("rows,F,H1,H2", SHAPES)
def test_mlp_backward_vs_oracle(rows, F, H1, H2):
    lib = _lib.lib()
    p = _params(F, H1, H2, 5, "cuda")
    gen = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(rows, F, device="cuda", generator=gen)
    ds = torch.randn(rows, device="cuda", generator=gen) * (torch.rand(rows, device="cuda", generator=gen) > 0.2)
    pn = [t.cpu().numpy() for t in p]
    # take the documents on a ReLU kink out of the comparison (zero upstream gradient)
    ds[torch.from_numpy(_near_kink(x.cpu().numpy(), pn)).cuda()] = 0.0
    rc, out = _call_backward(lib, x, p, ds)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    got = _split(out.cpu().numpy().astype(np.float64), F, H1, H2)
    ref_t = oracle.mlp_grads(x.cpu().numpy(), *pn, ds.cpu().numpy(), tf32=True)
    ref = oracle.mlp_grads(x.cpu().numpy(), *pn, ds.cpu().numpy())
    for name, g, rt, r in zip(("dW1", "db1", "dW2", "db2", "dW3", "db3"), got, ref_t, ref):
        scale = max(np.abs(r).max(), 1e-6)
        # TF32-operand restatement: float32 accumulation order and one-ulp TF32 roundings of dZ1 / dZ2 remain
        err = np.abs(g - rt.reshape(g.shape)).max()
        assert err <= 3e-4 * scale, (name, err, scale)
        # against the exact float64 model: TF32 operand noise flips the ReLU mask of the (document, unit) pairs
        # whose pre-activation is within ~1e-3 of zero, each flip moving a sum by one document's term, so this
        # is a norm-wise sanity bound only -- the comparison above is the parity check
        assert np.linalg.norm(g - r.reshape(g.shape)) <= 0.1 * max(np.linalg.norm(r), 1e-6), name
    # bit-reproducible
    rc, out2 = _call_backward(lib, x, p, ds)
    assert rc == 0 and torch.equal(out, out2)


def _near_kink_exact(x, pn):
    """documents with a layer-1 or layer-2 pre-activation within float32 accumulation error of zero"""
    d = np.float64
    t = oracle.tf32_trunc
    z1 = t(x).astype(d) @ t(pn[0]).astype(d).T + pn[1].astype(d)
    z2 = np.maximum(z1, 0.0) @ pn[2].astype(d).T + pn[3].astype(d)
    return (np.abs(z2) < 2e-5).any(axis=1) | (np.abs(z1) < 2e-5).any(axis=1)


@gpu
@pytest.mark.parametrize("rows,F,H1,H2", KEPT_SHAPES)
def test_mlp_backward_from_kept_activations_vs_oracle(rows, F, H1, H2):
    """ltr_mlp_scores(hz_out) + ltr_mlp_backward(hz): the activation rows hold relu(Z1 + b1) and Z2 of the forward
    pass (float32 layer 2), the masks are the forward pass's own, and dW1 / dW2 / db1 / db2 come out of one
    tensor-core product with TF32 operands (dZ1, dZ2, X, H1)."""
    lib = _lib.lib()
    p = _params(F, H1, H2, 9, "cuda")
    gen = torch.Generator(device="cuda").manual_seed(10)
    x = torch.randn(rows, F, device="cuda", generator=gen)
    ds = torch.randn(rows, device="cuda", generator=gen) * (torch.rand(rows, device="cuda", generator=gen) > 0.2)
    pn = [t.cpu().numpy() for t in p]
    xn = x.cpu().numpy()
    ds[torch.from_numpy(_near_kink_exact(xn, pn)).cuda()] = 0.0
    hz = _kept_activations(lib, x, p)
    assert hz is not None
    # the kept rows themselves: [H1 | 0 | Z2 | 0]
    d = np.float64
    t = oracle.tf32_trunc
    z1 = t(xn).astype(d) @ t(pn[0]).astype(d).T + pn[1].astype(d)
    h1 = np.maximum(z1, 0.0)
    z2 = h1 @ pn[2].astype(d).T + pn[3].astype(d)
    hzn = hz.cpu().numpy().astype(d)
    big = H1 > 32 or H2 > 8                           # kernel instantiation: 50-10, or 32-8 (hidden units padded)
    z0 = 52 if big else 32
    assert np.abs(hzn[:, :H1] - h1).max() <= 2e-5 * max(1.0, np.abs(h1).max())
    assert np.abs(hzn[:, z0:z0 + H2] - z2).max() <= 2e-5 * max(1.0, np.abs(z2).max())
    one = z0 + (10 if big else 8)                     # the column of ones that yields db1 / db2
    assert (hzn[:, H1:z0] == 0).all() and (hzn[:, z0 + H2:one] == 0).all()
    assert (hzn[:, one] == 1).all() and (hzn[:, one + 1:] == 0).all()
    assert hzn.shape[1] == (64 if big else 44)
    rc, out = _call_backward(lib, x, p, ds, hz)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    got = _split(out.cpu().numpy().astype(np.float64), F, H1, H2)
    # restatement of this path: exact masks, TF32 operands in dH1 = dZ2 W2, dW1 = dZ1^T X, dW2 = dZ2^T H1, db1, db2
    ref = oracle.mlp_grads_kept(xn, *pn, ds.cpu().numpy())
    exact = oracle.mlp_grads(xn, *pn, ds.cpu().numpy())
    for name, a, r, e in zip(("dW1", "db1", "dW2", "db2", "dW3", "db3"), got, ref, exact):
        scale = max(np.abs(e).max(), 1e-6)
        assert np.abs(a - r.reshape(a.shape)).max() <= 3e-4 * scale, name
        assert np.linalg.norm(a - e.reshape(a.shape)) <= 0.1 * max(np.linalg.norm(e), 1e-6), name
    rc, out2 = _call_backward(lib, x, p, ds, hz)
    assert rc == 0 and torch.equal(out, out2)
    # and the two backward paths agree with each other to the TF32 operand precision (where the recompute
    # kernel takes the shape: its two copies of a feature tile limit it to F <= ~140)
    rc, out3 = _call_backward(lib, x, p, ds)
    assert rc in (0, -2)
    if rc == 0:
        assert (out - out3).norm().item() <= 0.1 * out3.norm().item()


@gpu
def test_mlp_backward_is_additive_over_documents():
    """Size-independent property: the gradients of a launch whose CTAs each stream several tiles equal the sum
    of single-tile-per-CTA launches over slices of the same documents (every document's term is the same bits
    either way; only the float32 summation order differs)."""
    lib = _lib.lib()
    F, H1, H2, rows = 136, 50, 10, 128 * 148 * 3 + 77
    p = _params(F, H1, H2, 7, "cuda")
    gen = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(rows, F, device="cuda", generator=gen)
    ds = torch.randn(rows, device="cuda", generator=gen)
    hz = _kept_activations(lib, x, p)
    for kept in (None, hz):
        rc, full = _call_backward(lib, x, p, ds, kept)
        assert rc == 0
        acc = torch.zeros_like(full, dtype=torch.float64)
        step = 128 * 100
        for r0 in range(0, rows, step):
            rc, part = _call_backward(lib, x[r0:r0 + step].contiguous(), p, ds[r0:r0 + step].contiguous(),
                                      None if kept is None else kept[r0:r0 + step].contiguous())
            assert rc == 0
            acc += part.double()
        assert (full.double() - acc).abs().max().item() <= 2e-6 * acc.abs().max().item()


@gpu
def test_mlp_unsupported_shapes_return_code():
    lib = _lib.lib()
    p = _params(30, 50, 10, 0, "cuda")
    x = torch.randn(10, 30, device="cuda")
    assert _call_scores(lib, x, p)[0] == -2                  # F % 4 != 0
    p = _params(32, 70, 10, 0, "cuda")
    assert _call_scores(lib, torch.randn(10, 32, device="cuda"), p)[0] == -2     # H1 > 64
    # the recomputing backward keeps W1 and a whole tile in shared memory: rows beyond that need kept activations
    p = _params(700, 50, 10, 0, "cuda")
    x = torch.randn(10, 700, device="cuda")
    n = lib.ltr_mlp_grad_len(700, 50, 10)
    out = torch.empty(n, device="cuda")
    wsb = lib.ltr_mlp_workspace_bytes(700, 50, 10)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    ds = torch.randn(10, device="cuda")
    rc = lib.ltr_mlp_backward(x.data_ptr(), 10, 700, p[0].data_ptr(), p[1].data_ptr(), 50, p[2].data_ptr(),
                              p[3].data_ptr(), 10, p[4].data_ptr(), p[5].data_ptr(), None, ds.data_ptr(),
                              out.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    assert rc == -2


@gpu
@pytest.mark.parametrize("loss_name", ["PairwiseLogisticLoss", "LambdaNDCGLoss2", "ListNetLoss"])
def test_mlp_ranker_matches_unfused_modules(loss_name):
    """MLPRanker + a loss of this package == the documented torch model + the same loss: scores within the TF32
    operand precision, parameter gradients of loss.mean() likewise (examples/01-basic-usage.py:72-74 form)."""
    import pytorchltr_b200.loss as L
    from pytorchltr_b200.fused import MLPRanker
    torch.manual_seed(0)
    B, Lq, F = 64, 40, 136
    xs = torch.randn(B, Lq, F, device="cuda")
    ys = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    fused = MLPRanker(F).cuda()
    plain = torch.nn.Sequential(torch.nn.Linear(F, 50), torch.nn.ReLU(), torch.nn.Linear(50, 10), torch.nn.ReLU(),
                                torch.nn.Linear(10, 1)).cuda()
    with torch.no_grad():
        for a, b in zip((plain[0], plain[2], plain[4]), (fused.l1, fused.l2, fused.l3)):
            a.weight.copy_(b.weight)
            a.bias.copy_(b.bias)
    loss_fn = getattr(L, loss_name)()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s_f, s_p = fused(xs), plain(xs)
        assert s_f.shape == s_p.shape == (B, Lq, 1)
        assert (s_f - s_p).abs().max().item() <= 4e-3 * max(1.0, s_p.abs().max().item())
        loss_fn(s_f, ys, n).mean().backward()
        loss_fn(s_p, ys, n).mean().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    for a, b in zip((plain[0], plain[2], plain[4]), (fused.l1, fused.l2, fused.l3)):
        for ga, gb in ((a.weight.grad, b.weight.grad), (a.bias.grad, b.bias.grad)):
            assert gb is not None and torch.isfinite(gb).all()
            # TF32 operands in the scorer (and the ReLU-mask flips they cause next to a kink): norm-wise bound
            assert (ga - gb).norm().item() <= 0.1 * ga.norm().item() + 1e-6, loss_name   # (d/db3 of these losses is 0)


@gpu
@pytest.mark.parametrize("F", [46, 220, 699])
def test_mlp_ranker_other_feature_widths(F):
    """MQ2007 (46: padded to 48 on the fly), Istella (220: kernels as they are), Yahoo (699: padded to 700, W1
    streamed beside the tile, dW1 in four column slabs) -- all on the library's kernels, no fallback warning."""
    import warnings
    from pytorchltr_b200.fused import MLPRanker
    torch.manual_seed(2)
    B, Lq = 16, 40
    xs = torch.randn(B, Lq, F, device="cuda")
    fused = MLPRanker(F).cuda()
    plain = torch.nn.Sequential(torch.nn.Linear(F, 50), torch.nn.ReLU(), torch.nn.Linear(50, 10), torch.nn.ReLU(),
                                torch.nn.Linear(10, 1)).cuda()
    with torch.no_grad():
        for a, b in zip((plain[0], plain[2], plain[4]), (fused.l1, fused.l2, fused.l3)):
            a.weight.copy_(b.weight)
            a.bias.copy_(b.bias)
    g = torch.randn(B, Lq, 1, device="cuda")
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            s_f = fused(xs)
        s_p = plain(xs)
        s_f.backward(g)
        s_p.backward(g)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert (s_f - s_p).abs().max().item() <= 4e-3 * max(1.0, s_p.abs().max().item())
    for a, b in zip((plain[0], plain[2], plain[4]), (fused.l1, fused.l2, fused.l3)):
        assert a.weight.grad.shape == b.weight.grad.shape
        assert (a.weight.grad - b.weight.grad).norm().item() <= 0.1 * a.weight.grad.norm().item() + 1e-6
        assert (a.bias.grad - b.bias.grad).norm().item() <= 0.1 * a.bias.grad.norm().item() + 1e-6


@gpu
def test_mlp_ranker_trains():
    """A few SGD steps of the getting-started loop (rst:83-138) reduce the loss, fused and plain alike."""
    import pytorchltr_b200.loss as L
    from pytorchltr_b200.fused import MLPRanker
    torch.manual_seed(1)
    B, Lq, F = 128, 32, 136
    xs = torch.randn(B, Lq, F, device="cuda")
    w = torch.randn(F, device="cuda")
    ys = ((xs @ w) > 8).long() + ((xs @ w) > 0).long()
    n = torch.full((B,), Lq, device="cuda")
    model = MLPRanker(F).cuda()
    opt = torch.optim.Adagrad(model.parameters(), lr=0.1)
    loss_fn = L.PairwiseHingeLoss()
    first = last = None
    for step in range(30):
        loss = loss_fn(model(xs), ys, n).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    assert last < 0.6 * first


@gpu
def test_mlp_ranker_rejects_feature_gradients_and_captures():
    from pytorchltr_b200.fused import MLPRanker, mlp_scores
    m = MLPRanker(136).cuda()
    x = torch.randn(4, 8, 136, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        m(x)
    # forward + backward are capturable (no allocation / synchronisation inside the library calls)
    x = torch.randn(16, 64, 136, device="cuda")
    ds = torch.randn(16 * 64, device="cuda")
    lib = _lib.lib()
    p = [t.detach() for t in (m.l1.weight, m.l1.bias, m.l2.weight, m.l2.bias, m.l3.weight, m.l3.bias)]
    rc, want = _call_backward(lib, x.reshape(-1, 136), p, ds)
    assert rc == 0
    n = lib.ltr_mlp_grad_len(136, 50, 10)
    out = torch.zeros(n, device="cuda")
    wsb = lib.ltr_mlp_workspace_bytes(136, 50, 10)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            rc = lib.ltr_mlp_backward(x.data_ptr(), 16 * 64, 136, p[0].data_ptr(), p[1].data_ptr(), 50, p[2].data_ptr(),
                                      p[3].data_ptr(), 10, p[4].data_ptr(), p[5].data_ptr(), None, ds.data_ptr(),
                                      out.data_ptr(), ws.data_ptr(), wsb, s.cuda_stream)
            assert rc == 0
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)


@gpu
def test_vector_exchange_and_fused_backward_on_one_rank():
    """world = 1 (the rank is its own peer): the vector all-reduce is the identity, at any length and under graph
    replay; ltr_mlp_backward_allreduce returns ltr_mlp_backward's bits.  (Two ranks: tests/test_gpu_round2.py.)"""
    import ctypes
    lib = _lib.lib()
    h = ctypes.c_void_p()
    blob = (ctypes.c_ubyte * 64)()
    _lib.check(lib.ltr_p2p_create(0, 1, ctypes.byref(h), blob))
    try:
        _lib.check(lib.ltr_p2p_connect(h, blob))
        st = torch.cuda.current_stream().cuda_stream
        for k in (1, 77, 65536, 65537, 300001):
            v = torch.randn(k, device="cuda")
            ref = v.clone()
            for _ in range(3):
                _lib.check(lib.ltr_p2p_allreduce_vec(h, v.data_ptr(), k, st))
            assert torch.equal(v, ref)
        v = torch.randn(5000, device="cuda")
        ref = v.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _lib.check(lib.ltr_p2p_allreduce_vec(h, v.data_ptr(), 5000, side.cuda_stream))
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            _lib.check(lib.ltr_p2p_allreduce_vec(h, v.data_ptr(), 5000, torch.cuda.current_stream().cuda_stream))
        for _ in range(4):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(v, ref) and lib.ltr_p2p_error(h) == 0
        for (rows, F, H1, H2) in [(5000, 136, 50, 10), (700, 700, 50, 10), (300, 1400, 50, 10), (0, 136, 50, 10)]:
            p = _params(F, H1, H2, 5, "cuda")
            x = torch.randn(max(rows, 1), F, device="cuda")[:rows]
            ds = torch.randn(rows, device="cuda")
            hz = _kept_activations(lib, x, p) if rows else None
            rc, ref = _call_backward(lib, x, p, ds, hz)
            assert rc == 0
            n = lib.ltr_mlp_grad_len(F, H1, H2)
            out = torch.full((n,), float("nan"), device="cuda")
            wsb = lib.ltr_mlp_workspace_bytes(F, H1, H2)
            ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
            rc = lib.ltr_mlp_backward_allreduce(x.data_ptr(), rows, F, p[0].data_ptr(), p[1].data_ptr(), H1,
                                                p[2].data_ptr(), p[3].data_ptr(), H2, p[4].data_ptr(), p[5].data_ptr(),
                                                None if hz is None else hz.data_ptr(), ds.data_ptr(), out.data_ptr(),
                                                ws.data_ptr(), wsb, h, st)
            assert rc == 0
            torch.cuda.synchronize()
            assert torch.equal(out, ref)
        assert lib.ltr_p2p_error(h) == 0
    finally:
        lib.ltr_p2p_destroy(h)
