"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, the reference's
known answers and the committed reference fixtures.

Tolerance contract: per query
    |loss - ref64| <= 1e-5 * |ref64| + 1e-6
    |grad - ref64| <= 1e-5 * max_j |ref64[b, j]|   (+1e-7 absolute)
where ref64 is the reference evaluated with float64 scores (oracle / fixtures).
Integer work is bit-exact: rankings on tie-free inputs, hinge gradients.
"""
import glob
import os

import numpy as np
import pytest
import torch
from pytest import approx

import oracle
from known_answers import LOSS_CASES, METRIC_CASES

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))
ADDITIVE = ("hinge", "dcg_hinge", "logistic")
LAMBDA = ("arp1", "arp2", "ndcg1", "ndcg2")
ALL_LOSSES = ADDITIVE + LAMBDA + ("listnet",)


def _loss_module(mode, sigma=1.0):
    from pytorchltr_b200 import loss as L
    return {
        "hinge": lambda: L.PairwiseHingeLoss(),
        "dcg_hinge": lambda: L.PairwiseDCGHingeLoss(),
        "logistic": lambda: L.PairwiseLogisticLoss(sigma),
        "arp1": lambda: L.LambdaARPLoss1(sigma),
        "arp2": lambda: L.LambdaARPLoss2(sigma),
        "ndcg1": lambda: L.LambdaNDCGLoss1(sigma),
        "ndcg2": lambda: L.LambdaNDCGLoss2(sigma),
        "listnet": lambda: L.ListNetLoss(),
    }[mode]()


def _oracle_loss(mode, s, y, n, sigma=1.0, f32=False):
    if mode in ADDITIVE:
        # hinge family: compare with the float32 restatement -- a pair sitting on the kink
        # (1 - (s_i - s_j) == 0 after float32 rounding) is active in float32 and may not be in
        # float64, which changes an integer gradient count by one
        return oracle.pairwise_additive(mode, s, y, n, sigma=sigma, f32=f32 or "hinge" in mode)
    if mode == "listnet":
        return oracle.listnet(s, y, n)
    return oracle.lambda_loss(mode, s, y, n, sigma=sigma)


def _run_cuda(mode, s, y, n, sigma=1.0, weights=None):
    dev = torch.device("cuda", 0)
    st = torch.as_tensor(s).to(dev).requires_grad_(True)
    out = _loss_module(mode, sigma)(st, torch.as_tensor(y).to(dev), torch.as_tensor(n).to(dev))
    if weights is None:
        out.sum().backward()
    else:
        (out * torch.as_tensor(weights, dtype=torch.float32, device=dev)).sum().backward()
    return out.detach().cpu().double().numpy(), st.grad.cpu().double().numpy()


def _assert_parity(loss, grad, ref_loss, ref_grad, loss_rel=1e-5, grad_rel=1e-5):
    assert loss.shape == ref_loss.shape
    bad = np.abs(loss - ref_loss) > loss_rel * np.abs(ref_loss) + 1e-6
    assert not bad.any(), (loss[bad][:4], ref_loss[bad][:4])
    gmax = np.abs(ref_grad).max(axis=1, keepdims=True) if ref_grad.size else 0.0
    badg = np.abs(grad - ref_grad) > grad_rel * gmax + 1e-7
    assert not badg.any(), (grad[badg][:4], ref_grad[badg][:4])


def make_batch(seed, B, L, grades=5, full=False):
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn(B, L, generator=g, dtype=torch.float32)
    if full:
        n = torch.full((B,), L, dtype=torch.int64)
    else:
        n = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int64)
    rel = torch.randint(0, grades, (B, L), generator=g, dtype=torch.int64)
    rel[torch.arange(L)[None, :] >= n[:, None]] = 0
    return scores.numpy(), rel.numpy(), n.numpy()


# ------------------------------------------------------------------ known answers
@pytest.mark.parametrize("case", LOSS_CASES, ids=lambda c: f"{c['mode']}@{c['src']}")
def test_cuda_loss_known_answer(case):
    loss, grad = _run_cuda(case["mode"], case["scores"], case["rel"], case["n"], case["sigma"])
    if case["loss"].shape != ():
        assert loss == approx(case["loss"], rel=1e-5, abs=2e-6)
    if case["grad"] is not None:
        assert grad == approx(case["grad"], rel=1e-5, abs=1e-6)


@pytest.mark.parametrize("case", METRIC_CASES, ids=lambda c: f"{c['metric']}@{c['src']}")
def test_cuda_metric_known_answer(case):
    from pytorchltr_b200.evaluation import arp, dcg, ndcg
    dev = torch.device("cuda", 0)
    s = torch.as_tensor(case["scores"]).to(dev)
    y = torch.as_tensor(case["rel"]).to(dev)
    n = torch.as_tensor(case["n"]).to(dev)
    if case["metric"] == "arp":
        out = arp(s, y, n)
    elif case["metric"] == "dcg":
        out = dcg(s, y, n, k=case["k"], exp=case["exp"])
    else:
        out = ndcg(s, y, n, k=case["k"], exp=case["exp"])
    assert out.dtype == torch.float32 and out.shape == (s.shape[0],)
    assert out.cpu().double().numpy() == approx(case["expected"], rel=1e-5, abs=1e-6)


def test_cuda_hinge_doc_example_shapes():
    """(B, L, 1) scores and relevance are accepted (test_pairwise_additive.py:11-30)."""
    dev = torch.device("cuda", 0)
    loss_fn = _loss_module("hinge")
    s = torch.tensor([[0.0, 0.0, 1.0, 2.0, 1.0]], device=dev)
    y = torch.tensor([[[0], [0], [1], [2], [1]]], device=dev)
    n = torch.tensor([5], device=dev)
    assert loss_fn(s, y, n).item() == approx(0.0)
    s3 = s.reshape(1, 5, 1).clone().requires_grad_(True)
    out = loss_fn(s3, y.reshape(1, 5), n)
    out.sum().backward()
    assert out.shape == (1,) and s3.grad.shape == (1, 5, 1)


# ------------------------------------------------------------------ reference fixtures
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
def test_cuda_loss_vs_reference_fixture(path, mode):
    g = dict(np.load(path))
    loss, grad = _run_cuda(mode, g["scores"], g["relevance"], g["n"], float(g["sigma"]))
    _assert_parity(loss, grad, g[f"{mode}_loss64"], g[f"{mode}_grad64"])
    if mode == "hinge":
        # integer-valued gradient: bit-exact against the reference's float32 run
        assert np.array_equal(grad, g["hinge_grad32"].astype(np.float64))


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_cuda_metrics_vs_reference_fixture(path):
    from pytorchltr_b200.evaluation import arp, dcg, ndcg
    from pytorchltr_b200.utils import rank_by_score
    g = dict(np.load(path))
    dev = torch.device("cuda", 0)
    s = torch.as_tensor(g["scores"]).to(dev)
    y = torch.as_tensor(g["relevance"]).to(dev)
    n = torch.as_tensor(g["n"]).to(dev)
    L = s.shape[1]
    rk = rank_by_score(s, n).cpu().numpy()
    assert rk.dtype == np.int64
    for b, nb in enumerate(g["n"]):
        assert np.array_equal(rk[b, :nb], g["ranking"][b, :nb])
        assert sorted(rk[b, nb:]) == list(range(nb, L))
    for exp in (True, False):
        tag = "exp" if exp else "lin"
        assert dcg(s, y, n, exp=exp).cpu().numpy() == approx(g[f"dcg_all_{tag}"], rel=1e-5, abs=1e-6)
        assert ndcg(s, y, n, exp=exp).cpu().numpy() == approx(g[f"ndcg_all_{tag}"], rel=1e-5, abs=1e-6)
        for k in (1, 3, 10, 1000):
            assert dcg(s, y, n, k=k, exp=exp).cpu().numpy() == approx(
                g[f"dcg_k{k}_{tag}"], rel=1e-5, abs=1e-6)
            assert ndcg(s, y, n, k=k, exp=exp).cpu().numpy() == approx(
                g[f"ndcg_k{k}_{tag}"], rel=1e-5, abs=1e-6)
    assert arp(s, y, n).cpu().numpy() == approx(g["arp"], rel=1e-5, abs=1e-6)


# ------------------------------------------------------------------ random vs oracle
SHAPES = [(7, 1), (9, 2), (16, 5), (12, 31), (12, 32), (10, 33), (8, 37), (6, 64), (5, 100),
          (40, 128), (6, 129), (5, 200), (4, 256), (3, 500), (3, 512), (2, 1000), (2, 1024),
          (1, 2048)]


@pytest.mark.parametrize("mode", ALL_LOSSES)
@pytest.mark.parametrize("B,L", SHAPES)
def test_cuda_loss_vs_oracle_random(mode, B, L):
    if mode in ("arp1", "ndcg1") and L > 1024:
        pytest.skip("oracle too slow")
    s, y, n = make_batch(100 + L, B, L)
    # edge list sizes in the first rows: n = 0, 1, L (SURVEY 8(a) edge behaviour)
    for i, v in enumerate([0, 1, L][:B]):
        n[i] = min(v, L)
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    sigma = 1.0 if L % 2 else 0.5
    rng = np.random.default_rng(L)
    wts = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    loss, grad = _run_cuda(mode, s, y, n, sigma, weights=wts)
    ref_loss, ref_grad = _oracle_loss(mode, s, y, n, sigma)
    _assert_parity(loss, grad, ref_loss, ref_grad * wts[:, None].astype(np.float64))
    assert np.all(grad[np.arange(L)[None, :] >= n[:, None]] == 0.0)


@pytest.mark.parametrize("B,L", [(16, 5), (8, 37), (40, 128), (5, 200), (2, 1024)])
def test_cuda_hinge_gradient_bit_exact(B, L):
    s, y, n = make_batch(7 + L, B, L)
    # scores on a coarse grid so that many pairs sit exactly on the hinge kink (l == 0)
    s = np.round(s * 2.0) / 2.0
    loss, grad = _run_cuda("hinge", s, y, n)
    ref_loss, ref_grad = oracle.pairwise_additive("hinge", s, y, n, f32=True)
    assert np.array_equal(grad, ref_grad)
    assert loss == approx(ref_loss, rel=1e-6, abs=1e-6)


@pytest.mark.parametrize("B,L", [(16, 5), (8, 37), (64, 128), (9, 200), (4, 1024), (2, 4096)])
def test_cuda_rank_by_score_bit_exact(B, L):
    from pytorchltr_b200.utils import rank_by_score
    s, _, n = make_batch(55 + L, B, L)
    n[0] = 0
    dev = torch.device("cuda", 0)
    rk = rank_by_score(torch.as_tensor(s).to(dev).reshape(B, L, 1), torch.as_tensor(n).to(dev))
    assert rk.dtype == torch.int64 and rk.shape == (B, L)
    assert np.array_equal(rk.cpu().numpy(), oracle.rank_by_score(s, n))


def test_cuda_rank_by_score_ties_and_specials():
    """Ties -> lowest index first; -0.0 == +0.0; +/-inf ordered; padding last in index order."""
    from pytorchltr_b200.utils import rank_by_score
    dev = torch.device("cuda", 0)
    s = torch.tensor([[1.0, 1.0, -0.0, 0.0, float("inf"), -float("inf"), 1.0, 9.0]], device=dev)
    n = torch.tensor([7], device=dev)
    assert rank_by_score(s, n).cpu().tolist() == [[4, 0, 1, 6, 2, 3, 5, 7]]


@pytest.mark.parametrize("B,L", [(16, 5), (8, 37), (64, 128), (9, 200), (3, 1024)])
@pytest.mark.parametrize("exp", [True, False])
def test_cuda_metrics_vs_oracle_random(B, L, exp):
    from pytorchltr_b200.evaluation import arp, dcg, ndcg
    s, y, n = make_batch(900 + L, B, L)
    n[0] = 0
    n[1] = 1
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    dev = torch.device("cuda", 0)
    st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    tol = dict(rel=1e-5, abs=1e-6)
    assert dcg(st, yt, nt, exp=exp).cpu().numpy() == approx(oracle.dcg(s, y, n, exp=exp), **tol)
    assert ndcg(st, yt, nt, exp=exp).cpu().numpy() == approx(oracle.ndcg(s, y, n, exp=exp), **tol)
    for k in (1, 10, L, L + 5):
        assert dcg(st, yt, nt, k=k, exp=exp).cpu().numpy() == approx(
            oracle.dcg(s, y, n, k=k, exp=exp), **tol)
        assert ndcg(st, yt, nt, k=k, exp=exp).cpu().numpy() == approx(
            oracle.ndcg(s, y, n, k=k, exp=exp), **tol)
    assert arp(st, yt, nt).cpu().numpy() == approx(oracle.arp(s, y, n), **tol)
    with pytest.raises(IndexError):
        dcg(st, yt, nt, k=0)


# ------------------------------------------------------------------ boundary behaviour
def test_cuda_int32_inputs_and_cpu_tensors():
    """int32 relevance / n are accepted (tests/utils/test_tensor_operations.py:27) and CPU
    tensors are staged through the GPU and come back as CPU tensors."""
    s, y, n = make_batch(3, 6, 50)
    ref_loss, ref_grad = oracle.lambda_loss("ndcg2", s, y, n)
    dev = torch.device("cuda", 0)
    st = torch.as_tensor(s).to(dev).requires_grad_(True)
    out = _loss_module("ndcg2")(st, torch.as_tensor(y).to(dev).int(), torch.as_tensor(n).to(dev).int())
    out.sum().backward()
    _assert_parity(out.detach().cpu().double().numpy(), st.grad.cpu().double().numpy(), ref_loss, ref_grad)
    sc = torch.as_tensor(s).clone().requires_grad_(True)
    out = _loss_module("ndcg2")(sc, torch.as_tensor(y), torch.as_tensor(n))
    assert out.device.type == "cpu"
    out.sum().backward()
    assert sc.grad.device.type == "cpu"
    _assert_parity(out.detach().double().numpy(), sc.grad.double().numpy(), ref_loss, ref_grad)


def test_cuda_no_grad_and_empty_batch():
    dev = torch.device("cuda", 0)
    s, y, n = make_batch(4, 5, 20)
    with torch.no_grad():
        out = _loss_module("arp2")(torch.as_tensor(s).to(dev), torch.as_tensor(y).to(dev),
                                   torch.as_tensor(n).to(dev))
    assert not out.requires_grad
    assert out.cpu().double().numpy() == approx(oracle.lambda_loss("arp2", s, y, n)[0], rel=1e-5)
    e = _loss_module("hinge")(torch.zeros(0, 7, device=dev), torch.zeros(0, 7, dtype=torch.long, device=dev),
                              torch.zeros(0, dtype=torch.long, device=dev))
    assert e.shape == (0,)


def test_cabi_error_codes_and_host_entry():
    import ctypes
    from pytorchltr_b200 import _lib
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    B, L = 33, 77
    s, y, n = make_batch(11, B, L)
    ds, dy, dn = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    loss = torch.empty(B, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (ds.data_ptr(), dy.data_ptr(), 8, dn.data_ptr(), 8, B, L, 1.0, loss.data_ptr(), None, None, None, st)
    assert lib.ltr_lambda(9, *args) == -1                                   # bad mode
    assert lib.ltr_lambda(3, ds.data_ptr(), dy.data_ptr(), 3, *args[3:]) == -1   # bad dtype width (8, 4, 2, 1 are valid)
    big = (ds.data_ptr(), dy.data_ptr(), 8, dn.data_ptr(), 8, B, _lib.MAX_LIST_SIZE + 1, 1.0,
           loss.data_ptr(), None, None, None, st)
    assert lib.ltr_lambda(3, *big) == -2                                    # unsupported L
    assert lib.ltr_lambda(3, None, *args[1:]) == -1                          # NULL scores
    assert b"invalid" in lib.ltr_strerror(-1)
    with pytest.raises(_lib.LtrError):
        _lib.check(-1)

    # host-buffer entry point: pinned host memory in, pinned host memory out
    hs = torch.as_tensor(s).pin_memory()
    hy = torch.as_tensor(y).pin_memory()
    hn = torch.as_tensor(n).pin_memory()
    hl = torch.empty(B, dtype=torch.float32).pin_memory()
    hg = torch.empty(B, L, dtype=torch.float32).pin_memory()
    nbytes = lib.ltr_host_workspace_bytes(B, L)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for family, mode, name in ((0, 2, "logistic"), (1, 3, "ndcg2"), (2, 0, "listnet")):
        rc = lib.ltr_loss_host(family, mode, hs.data_ptr(), hy.data_ptr(), hn.data_ptr(), B, L, 1.0,
                               hl.data_ptr(), hg.data_ptr(), ws.data_ptr(), ctypes.c_size_t(nbytes), st)
        assert rc == 0
        torch.cuda.synchronize()
        ref_loss, ref_grad = _oracle_loss(name, s, y, n)
        _assert_parity(hl.double().numpy(), hg.double().numpy(), ref_loss, ref_grad)
    assert lib.ltr_loss_host(1, 3, hs.data_ptr(), hy.data_ptr(), hn.data_ptr(), B, L, 1.0, hl.data_ptr(),
                             hg.data_ptr(), ws.data_ptr(), ctypes.c_size_t(16), st) == -1


def test_cuda_graph_capture_and_loss_sum():
    """The launches are capturable (no sync, no allocation inside the library) and the fused
    per-device loss sum equals the sum of the per-query losses."""
    from pytorchltr_b200 import _lib, _ops
    dev = torch.device("cuda", 0)
    B, L = 256, 128
    s, y, n = make_batch(21, B, L)
    ds, dy, dn = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    total = torch.zeros(1, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            _ops.launch_loss(_lib.FAMILY_LAMBDA, _lib.LAM_NDCG2, ds, dy, dn, 1.0, True)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        total.zero_()
        loss, grad, _ = _ops.launch_loss(_lib.FAMILY_LAMBDA, _lib.LAM_NDCG2, ds, dy, dn, 1.0, True,
                                         loss_sum=total)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ref_loss, ref_grad = oracle.lambda_loss("ndcg2", s, y, n)
    _assert_parity(loss.cpu().double().numpy(), grad.cpu().double().numpy(), ref_loss, ref_grad)
    assert total.item() == approx(ref_loss.sum(), rel=1e-5)


# ------------------------------------------------------------------ full BASELINE sizes
@pytest.mark.parametrize("mode,B,L", [("ndcg2", 4096, 128), ("dcg_hinge", 1024, 1024),
                                      ("ndcg2", 512, 1024), ("listnet", 8192, 200)])
def test_cuda_full_size_properties(mode, B, L):
    """At BASELINE.json sizes the oracle is too slow for every query: check size-independent
    properties on the whole batch and the oracle on a sample of queries."""
    s, y, n = make_batch(1234, B, L)
    loss, grad = _run_cuda(mode, s, y, n)
    assert np.isfinite(loss).all() and np.isfinite(grad).all()
    # antisymmetric pair gradients / softmax - P: every query's gradient sums to zero
    scale = np.abs(grad).sum(axis=1) + 1e-30
    assert np.all(np.abs(grad.sum(axis=1)) <= 2e-5 * scale + 1e-6)
    assert np.all(grad[np.arange(L)[None, :] >= n[:, None]] == 0.0)
    # padded scores must not matter
    s2 = s.copy()
    s2[np.arange(L)[None, :] >= n[:, None]] = 1e3
    loss2, grad2 = _run_cuda(mode, s2, y, n)
    if L <= 128 or mode in ("arp1", "ndcg1", "listnet"):
        assert np.array_equal(loss, loss2) and np.array_equal(grad, grad2)
    else:
        # the CTA-per-query tile kernel hands tiles to warps dynamically and merges them with
        # shared-memory float atomics: equal up to float32 summation order
        assert loss2 == approx(loss, rel=2e-6, abs=1e-6)
        assert np.all(np.abs(grad2 - grad) <= 2e-6 * np.abs(grad).max(axis=1, keepdims=True) + 1e-7)
    # permutation equivariance: shuffling the valid documents of a query permutes the
    # gradient and leaves the loss unchanged (up to float32 summation order)
    rng = np.random.default_rng(0)
    sp, yp, perms = s.copy(), y.copy(), []
    for b in range(B):
        p = rng.permutation(n[b])
        perms.append(p)
        sp[b, :n[b]] = s[b, p]
        yp[b, :n[b]] = y[b, p]
    lossp, gradp = _run_cuda(mode, sp, yp, n)
    assert lossp == approx(loss, rel=2e-5, abs=1e-5)
    for b in range(0, B, max(1, B // 64)):
        assert gradp[b, :n[b]] == approx(grad[b, perms[b]], rel=1e-4, abs=1e-5 * np.abs(grad[b]).max() + 1e-7)
    # oracle on a sample
    idx = np.arange(0, B, max(1, B // 32))
    ref_loss, ref_grad = _oracle_loss(mode, s[idx], y[idx], n[idx])
    _assert_parity(loss[idx], grad[idx], ref_loss, ref_grad)


def test_cuda_full_size_ranking_and_ndcg():
    from pytorchltr_b200.evaluation import ndcg
    from pytorchltr_b200.utils import rank_by_score
    B, L = 8192, 200
    s, y, n = make_batch(4321, B, L)
    dev = torch.device("cuda", 0)
    st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    rk = rank_by_score(st, nt)
    assert torch.equal(torch.sort(rk, dim=1).values, torch.arange(L, device=dev).expand(B, L))
    sorted_scores = torch.gather(st, 1, rk)
    valid = torch.arange(L, device=dev)[None, :] < nt[:, None]
    d = sorted_scores[:, 1:] - sorted_scores[:, :-1]
    assert bool((d[valid[:, 1:]] <= 0).all())                  # sortedness of the valid prefix
    assert bool((rk[~valid] >= nt[:, None].expand(B, L)[~valid]).all())   # padding last
    out = ndcg(st, yt, nt, k=10).cpu().numpy()
    assert np.all((out >= 0) & (out <= 1 + 1e-6))
    idx = np.arange(0, B, 64)
    assert out[idx] == approx(oracle.ndcg(s[idx], y[idx], n[idx], k=10), rel=1e-5, abs=1e-6)
    # idempotence: ranking by the (negated) rank position reproduces the ranking
    pos = torch.empty_like(st)
    pos.scatter_(1, rk, -torch.arange(L, device=dev, dtype=torch.float32).expand(B, L))
    rk2 = rank_by_score(pos, nt)
    assert torch.equal(rk2[valid], rk[valid])


# ------------------------------------------------------------------ slow paths of the tile kernels
@pytest.mark.parametrize("mode", ["logistic", "arp1", "arp2", "ndcg1", "ndcg2"])
@pytest.mark.parametrize("B,L", [(6, 37), (8, 128), (4, 300), (2, 1024)])
def test_cuda_large_score_range_uses_stable_form(mode, B, L):
    """sigma * (max - min) * log2(e) > 64: the factored exponential is replaced by exp(-|x|)."""
    s, y, n = make_batch(31 + L, B, L)
    s = (s * 40.0).astype(np.float32)
    loss, grad = _run_cuda(mode, s, y, n, sigma=1.0)
    ref_loss, ref_grad = _oracle_loss(mode, s, y, n, 1.0)
    _assert_parity(loss, grad, ref_loss, ref_grad)


@pytest.mark.parametrize("B,L", [(8, 128), (6, 100), (3, 300)])
def test_cuda_near_tied_scores_take_exact_sort(B, L):
    """Scores a few ulps apart cannot be separated by the 25-bit packed sort key: the exact
    64-bit network must kick in and the ranking stay bit-exact."""
    from pytorchltr_b200 import _lib, _ops
    rng = np.random.default_rng(5)
    s = np.empty((B, L), dtype=np.float32)
    for b in range(B):
        steps = rng.permutation(L).astype(np.int64)
        base = np.float32(1.0 + b)
        s[b] = (base.view(np.int32) + steps * rng.integers(1, 4)).astype(np.int32).view(np.float32)
    _, y, n = make_batch(77, B, L)
    n[0] = L
    dev = torch.device("cuda", 0)
    st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    loss, grad, ranking = _ops.launch_loss(_lib.FAMILY_LAMBDA, _lib.LAM_NDCG2, st, yt, nt, 1.0, True,
                                           want_ranking=True)
    ref_loss, ref_grad, ref_rank = oracle.lambda_loss("ndcg2", s, y, n, want_ranking=True)
    assert np.array_equal(ranking.cpu().numpy(), ref_rank)
    _assert_parity(loss.cpu().double().numpy(), grad.cpu().double().numpy() , ref_loss, ref_grad)


@pytest.mark.parametrize("B,L", [(8, 64), (5, 128), (3, 400)])
def test_cuda_wide_relevance_grades(B, L):
    """Grades outside [0, 31] leave the histogram fast path of the ideal DCG."""
    s, y, n = make_batch(13 + L, B, L)
    y = np.where(y == 4, 40, np.where(y == 3, 33, y)).astype(np.int64)
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    loss, grad = _run_cuda("ndcg2", s, y, n)
    ref_loss, ref_grad = oracle.lambda_loss("ndcg2", s, y, n)
    _assert_parity(loss, grad, ref_loss, ref_grad)


def test_cuda_lambda_ranking_out_matches_rank_by_score():
    from pytorchltr_b200 import _lib, _ops
    from pytorchltr_b200.utils import rank_by_score
    dev = torch.device("cuda", 0)
    for B, L in ((16, 100), (4, 700)):
        s, y, n = make_batch(3 + L, B, L)
        st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
        for mode in (_lib.LAM_ARP1, _lib.LAM_ARP2, _lib.LAM_NDCG1, _lib.LAM_NDCG2):
            _, _, ranking = _ops.launch_loss(_lib.FAMILY_LAMBDA, mode, st, yt, nt, 1.0, True, want_ranking=True)
            assert torch.equal(ranking, rank_by_score(st, nt))


@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
@pytest.mark.parametrize("B,L", [(9, 5), (6, 100), (3, 300)])
def test_cuda_generic_kernel_vs_oracle(mode, B, L, monkeypatch):
    """LTR_KERNEL=generic routes every loss through the one-CTA-per-query reference kernel (the
    second, independent CUDA implementation of the pair losses)."""
    monkeypatch.setenv("LTR_KERNEL", "generic")
    s, y, n = make_batch(500 + L, B, L)
    n[0] = 0
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    loss, grad = _run_cuda(mode, s, y, n, sigma=0.8)
    ref_loss, ref_grad = _oracle_loss(mode, s, y, n, 0.8)
    _assert_parity(loss, grad, ref_loss, ref_grad)


def test_cuda_warp_kernel_dynamic_queue_many_queries():
    """More queries than resident warps: the warp-per-query kernel switches to its device-wide
    work queue (self-resetting counter).  Run it several times back to back and check a sample."""
    B, L = 20000, 64
    s, y, n = make_batch(99, B, L)
    idx = np.arange(0, B, 97)
    ref_loss, ref_grad = oracle.lambda_loss("ndcg2", s[idx], y[idx], n[idx])
    for _ in range(3):
        loss, grad = _run_cuda("ndcg2", s, y, n)
        assert np.isfinite(loss).all()
        _assert_parity(loss[idx], grad[idx], ref_loss, ref_grad)
    # every query was visited exactly once: untouched outputs would keep torch.empty garbage,
    # so compare two full runs bit for bit (the kernel is deterministic per query)
    loss2, grad2 = _run_cuda("ndcg2", s, y, n)
    assert np.array_equal(loss, loss2) and np.array_equal(grad, grad2)


def test_cuda_rank_by_plackettluce_statistics():
    """Plackett-Luce sampling (reference tests/utils/test_tensor_operations.py:24-127 style):
    the first rank follows softmax(scores), padded documents always come last."""
    from pytorchltr_b200.utils import rank_by_plackettluce
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(42)
    runs = 4000
    scores = torch.tensor([[5.0, 3.0, 2.0, 1.0, 9.0]], device=dev).repeat(runs, 1)
    n = torch.full((runs,), 4, dtype=torch.int32, device=dev)
    rk = rank_by_plackettluce(scores.reshape(runs, 5, 1), n, generator=gen)
    assert rk.shape == (runs, 5) and rk.dtype == torch.int64
    assert bool((rk[:, 4] == 4).all())                       # the padded document is last
    first = torch.bincount(rk[:, 0], minlength=5).float().cpu().numpy()[:4] / runs
    expected = torch.softmax(torch.tensor([5.0, 3.0, 2.0, 1.0]), dim=0).numpy()
    assert first == approx(expected, abs=0.03)
    # second rank given the first: P(second = j) = sum_i p_i p_j / (1 - p_i)
    second = torch.bincount(rk[:, 1], minlength=5).float().cpu().numpy()[:4] / runs
    p = expected.astype(np.float64)
    exp2 = np.array([sum(p[i] * p[j] / (1 - p[i]) for i in range(4) if i != j) for j in range(4)])
    assert second == approx(exp2, abs=0.03)


@pytest.mark.parametrize("mode", ADDITIVE + LAMBDA)
@pytest.mark.parametrize("B,L", [(4, 200), (2, 1000)])
def test_cuda_tile_kernel_vs_oracle(mode, B, L, monkeypatch):
    """LTR_KERNEL=tiles keeps the 128 x 128 rank-tile kernel (the path of lists longer than 1024)
    for every L > 128, so it stays covered at sizes the oracle finishes quickly."""
    monkeypatch.setenv("LTR_KERNEL", "tiles")
    s, y, n = make_batch(700 + L, B, L)
    n[0] = 1
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    loss, grad = _run_cuda(mode, s, y, n, sigma=1.2)
    ref_loss, ref_grad = _oracle_loss(mode, s, y, n, 1.2)
    _assert_parity(loss, grad, ref_loss, ref_grad)


@pytest.mark.parametrize("mode", ["ndcg2", "hinge", "arp1"])
def test_cuda_ring_kernel_scheduled_many_queries(mode):
    """More queries than resident CTAs: the ring kernel orders the queries longest-first (counting
    sort in the scheduling workspace) and hands them out through the device-wide queue.  Every
    query must be visited exactly once, whatever n it has (0, 1, short, long)."""
    B, L = 6000, 160
    s, y, n = make_batch(321, B, L)
    n[::7] = torch.randint(0, 40, n[::7].shape, generator=torch.Generator().manual_seed(5))
    n[1], n[2] = 0, 1
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    idx = np.arange(0, B, 53)
    ref_loss, ref_grad = _oracle_loss(mode, s[idx], y[idx], n[idx])
    loss, grad = _run_cuda(mode, s, y, n)
    assert np.isfinite(loss).all()
    _assert_parity(loss[idx], grad[idx], ref_loss, ref_grad)
    loss2, grad2 = _run_cuda(mode, s, y, n)
    # per query the kernel is deterministic (static split over the warps, no atomics)
    assert np.array_equal(loss, loss2) and np.array_equal(grad, grad2)


@pytest.mark.parametrize("B,L", [(9000, 96), (3000, 300)])
def test_cabi_schedule_workspace_is_optional_and_does_not_change_results(B, L):
    """ltr_lambda (batch order) and ltr_lambda_ws (longest-first) return the same bits; a
    workspace that is too small or NULL falls back to batch order."""
    from pytorchltr_b200 import _lib
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    s, y, n = make_batch(11 + L, B, L)
    st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    stream = torch.cuda.current_stream(dev).cuda_stream
    outs = []
    need = lib.ltr_schedule_workspace_bytes(B)
    assert need >= 4 * B
    for ws_bytes in (None, need, 8):
        loss = torch.empty(B, device=dev)
        grad = torch.empty(B, L, device=dev)
        if ws_bytes is None:
            rc = lib.ltr_lambda(_lib.LAM_NDCG2, st.data_ptr(), yt.data_ptr(), 8, nt.data_ptr(), 8, B, L, 1.0,
                                loss.data_ptr(), grad.data_ptr(), None, None, stream)
        else:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            rc = lib.ltr_lambda_ws(_lib.LAM_NDCG2, st.data_ptr(), yt.data_ptr(), 8, nt.data_ptr(), 8, B, L, 1.0,
                                   loss.data_ptr(), grad.data_ptr(), None, None, ws.data_ptr(), ws_bytes, stream)
        assert rc == 0
        torch.cuda.synchronize()
        outs.append((loss.cpu().numpy(), grad.cpu().numpy()))
    for l, g in outs[1:]:
        assert np.array_equal(l, outs[0][0]) and np.array_equal(g, outs[0][1])


@pytest.mark.parametrize("B,L", [(12, 5), (9, 37), (33, 128), (17, 200), (9, 500), (8, 1024), (8, 202), (8, 96),
                                 (8, 150), (8, 190), (8, 250), (8, 300), (8, 700)])
def test_cuda_topk_metrics_edge_cases(B, L, monkeypatch):
    """dcg@k / ndcg@k for k <= 32 take the top-k selection kernel: tied scores (lowest index first,
    incl. the massive-tie fallback), lists shorter than k, unmasked padded relevance (dcg.py:85),
    wide and negative grades, and bit-for-bit the same ranking decisions as the full-sort kernel."""
    from pytorchltr_b200.evaluation import dcg, ndcg
    rng = np.random.default_rng(L)
    s, y, n = (np.array(a) for a in make_batch(77 + L, B, L))
    s[0] = 0.25                                  # every score tied
    s[1] = np.round(s[1] * 2) / 2                # a handful of distinct scores
    s[2, : L // 2] = s[2, L // 2: 2 * (L // 2)]  # exact duplicates
    n[3] = min(3, L)                             # fewer valid documents than k
    n[4] = 0
    y[5] = rng.integers(-3, 40, size=L)          # grades outside 0..7
    y[6] = rng.integers(0, 3, size=L) * 9
    # padded relevance is deliberately NOT zeroed in rows 7..: the reference does not mask it
    y[:7][np.arange(L)[None, :] >= n[:7, None]] = 0
    dev = torch.device("cuda", 0)
    st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
    tol = dict(rel=1e-5, abs=1e-6)
    for exp in (True, False):
        for k in (1, 2, 5, 10, 32):
            got_d = dcg(st, yt, nt, k=k, exp=exp).cpu().numpy()
            got_n = ndcg(st, yt, nt, k=k, exp=exp).cpu().numpy()
            assert got_d == approx(oracle.dcg(s, y, n, k=k, exp=exp), **tol), (k, exp)
            assert got_n == approx(oracle.ndcg(s, y, n, k=k, exp=exp), **tol), (k, exp)
    # int32 inputs take the same path
    got = ndcg(st, yt.int(), nt.int(), k=10).cpu().numpy()
    assert got == approx(oracle.ndcg(s, y, n, k=10), **tol)


# ------------------------------------------------------------------ fused linear scorer + ListNet (N1)
def _fused_batch(seed, B, L, F):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(B, L, F, generator=g, dtype=torch.float32)
    w = torch.randn(F, generator=g, dtype=torch.float32) * 0.3
    b = torch.randn(1, generator=g, dtype=torch.float32)
    n = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int64)
    y = torch.randint(0, 5, (B, L), generator=g, dtype=torch.int64)
    return X, w, b, y, n


@pytest.mark.parametrize("B,L,F", [(5, 8, 4), (37, 200, 136), (300, 40, 64), (3, 100, 256), (4, 50, 12),
                                   (6, 30, 10), (2, 300, 136), (3, 1000, 136), (2, 700, 45), (2, 64, 700),
                                   (300, 33, 7)])
def test_cuda_fused_linear_listnet_vs_oracle(B, L, F):
    """ltr_linear_listnet (one pass over the features) against the float64 oracle: loss and dscores
    within 1e-5 (relative to the row maximum), dweight within 1e-5 of sum |dscores * x| (the scale
    its float32 accumulation error grows with).  Shapes the TMA-staged kernel does not take (F % 4 != 0,
    a feature block larger than shared memory, more features than threads) run the tiled kernel
    (linear_listnet_tiled_kernel): still one launch, no library GEMV."""
    from pytorchltr_b200.fused import linear_listnet
    X, w, b, y, n = _fused_batch(B * 1000 + L + F, B, L, F)
    n[0] = 0
    if B > 1:
        n[1] = 1
    dev = torch.device("cuda", 0)
    wd = w.to(dev).requires_grad_(True)
    bd = b.to(dev).requires_grad_(True)
    loss = linear_listnet(X.to(dev), wd, bd, y.to(dev), n.to(dev))
    loss.sum().backward()
    s_ref, l_ref, d_ref, dw_ref, db_ref, gscale = oracle.linear_listnet(X.numpy(), w.numpy(), b.numpy(),
                                                                        y.numpy(), n.numpy())
    got = loss.detach().cpu().double().numpy()
    assert np.all(np.abs(got - l_ref) <= 1e-5 * np.abs(l_ref) + 2e-6), np.abs(got - l_ref).max()
    gw = wd.grad.cpu().double().numpy()
    assert np.all(np.abs(gw - dw_ref) <= 1e-5 * gscale + 1e-6), (np.abs(gw - dw_ref) / (gscale + 1e-30)).max()
    assert abs(float(bd.grad.cpu()) - db_ref) <= 1e-5 * np.abs(d_ref).sum() + 1e-6
    # .mean() scales the same gradients; a per-query upstream gradient takes the second pass
    wd.grad = None
    bd.grad = None
    wts = torch.rand(B, generator=torch.Generator().manual_seed(3)) + 0.5
    (linear_listnet(X.to(dev), wd, bd, y.to(dev), n.to(dev)) * wts.to(dev)).sum().backward()
    dw_w = np.einsum("bl,blf->f", d_ref * wts.double().numpy()[:, None], X.double().numpy())
    gw = wd.grad.cpu().double().numpy()
    assert np.all(np.abs(gw - dw_w) <= 2e-5 * gscale * 1.5 + 1e-6)


def test_cuda_fused_linear_listnet_module_matches_unfused_modules():
    """LinearListNet(F) == ListNetLoss()(Linear(F, 1)(xs), ys, n) with the same parameters: loss,
    weight and bias gradients (float32 on both sides, so the tolerance is the accumulation order)."""
    from pytorchltr_b200.fused import LinearListNet
    from pytorchltr_b200.loss import ListNetLoss
    B, L, F = 64, 200, 136
    X, w, b, y, n = _fused_batch(7, B, L, F)
    dev = torch.device("cuda", 0)
    fused = LinearListNet(F).to(dev)
    with torch.no_grad():
        fused.linear.weight.copy_(w.reshape(1, F))
        fused.linear.bias.copy_(b)
    plain = torch.nn.Linear(F, 1).to(dev)
    plain.load_state_dict(fused.linear.state_dict())
    Xd, yd, nd = X.to(dev), y.to(dev), n.to(dev)
    lf = fused(Xd, yd, nd)
    lf.mean().backward()
    lp = ListNetLoss()(plain(Xd), yd, nd)
    lp.mean().backward()
    assert torch.allclose(lf, lp, rtol=1e-5, atol=1e-5)
    assert torch.allclose(fused.linear.weight.grad, plain.weight.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(fused.linear.bias.grad, plain.bias.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mode", ["hinge", "dcg_hinge"])
@pytest.mark.parametrize("B,L", [(6, 200), (3, 1000), (40, 130), (3, 1500), (2, 4096)])
def test_cuda_sorted_hinge_matches_pair_kernels_and_oracle(mode, B, L, monkeypatch):
    """Lists longer than 384 take the O(n log n) sorted hinge kernel (shorter ones the pair kernels: the sorted
    form only pays beyond the measured crossover); LTR_HINGE forces either.  Scores on a 1/4 grid put many
    pairs exactly on the kink (s_i - s_j == 1: active, gradient -1 / +1) and create ties; the integer
    valued gradients must equal the float32 restatement of the reference bit for bit, and the O(n^2)
    pair kernels (LTR_HINGE=pairs) must agree."""
    rng = np.random.default_rng(B * L)
    s, y, n = (np.array(a) for a in make_batch(31 + L, B, L))
    s[0] = np.round(s[0] * 4) / 4
    s[1] = np.round(s[1] * 2) / 2
    if B > 2:
        s[2] = 0.5                      # every score tied
    n[0] = L
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    if B > 4:
        y[4, : n[4]] = rng.integers(-5, 60, size=n[4])   # grades outside 0..31: in-kernel O(n^2) fallback
    wts = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    ref_loss, ref_grad = oracle.pairwise_additive(mode, s, y, n, f32=True)
    loss0, grad0 = _run_cuda(mode, s, y, n, weights=wts)                  # the form the list size selects
    _assert_parity(loss0, grad0, ref_loss, ref_grad * wts[:, None].astype(np.float64))
    monkeypatch.setenv("LTR_HINGE", "sorted")
    loss, grad = _run_cuda(mode, s, y, n, weights=wts)
    _assert_parity(loss, grad, ref_loss, ref_grad * wts[:, None].astype(np.float64))
    if mode == "hinge":
        assert np.array_equal(grad, ref_grad * wts[:, None].astype(np.float64).astype(np.float32)) or \
            np.array_equal(grad.astype(np.float32), (ref_grad.astype(np.float32) * wts[:, None]))
        assert np.array_equal(grad, grad0)
    monkeypatch.setenv("LTR_HINGE", "pairs")
    loss2, grad2 = _run_cuda(mode, s, y, n, weights=wts)
    assert np.allclose(loss, loss2, rtol=1e-5, atol=1e-6)
    if mode == "hinge":
        assert np.array_equal(grad, grad2)
    else:
        assert np.allclose(grad, grad2, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("mode", ["ndcg2", "hinge", "logistic", "arp1"])
@pytest.mark.parametrize("L", [130, 203, 512])
def test_cuda_long_lists_int32_cpu_tensors_and_unaligned_rows(mode, L):
    """Lists longer than 128 (ring / sorted-hinge kernels): int32 relevance and n (4-byte TMA rows),
    list sizes that are not a multiple of 4 (plain-load path instead of TMA), CPU tensors through
    ltr_loss_host, and a non-contiguous scores view (made contiguous by the wrapper)."""
    B = 5
    s, y, n = make_batch(1700 + L, B, L)
    n[0] = 3
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    ref_loss, ref_grad = _oracle_loss(mode, s, y, n)
    dev = torch.device("cuda", 0)
    # int32 inputs
    st = torch.as_tensor(s).to(dev).requires_grad_(True)
    out = _loss_module(mode)(st, torch.as_tensor(y).to(dev).int(), torch.as_tensor(n).to(dev).int())
    out.sum().backward()
    _assert_parity(out.detach().cpu().double().numpy(), st.grad.cpu().double().numpy(), ref_loss, ref_grad)
    # CPU tensors (host-buffer entry point), (B, L, 1) scores like a model output
    sc = torch.as_tensor(s).clone().reshape(B, L, 1).requires_grad_(True)
    out = _loss_module(mode)(sc, torch.as_tensor(y), torch.as_tensor(n))
    assert out.device.type == "cpu"
    out.mean().backward()
    assert sc.grad.shape == (B, L, 1) and sc.grad.device.type == "cpu"
    _assert_parity(out.detach().double().numpy(), sc.grad.reshape(B, L).double().numpy() * B, ref_loss, ref_grad)
    # a strided view: every second column of a wider tensor
    wide = torch.zeros(B, 2 * L, device=dev)
    wide[:, ::2] = torch.as_tensor(s).to(dev)
    sv = wide[:, ::2].detach().requires_grad_(True)
    out = _loss_module(mode)(sv, torch.as_tensor(y).to(dev), torch.as_tensor(n).to(dev))
    out.sum().backward()
    _assert_parity(out.detach().cpu().double().numpy(), sv.grad.cpu().double().numpy(), ref_loss, ref_grad)
