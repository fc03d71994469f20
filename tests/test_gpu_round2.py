"""GPU tests of the round-2 surface: narrow relevance dtypes, the pipelined host-buffer entry points,
`.mean().backward()` of CPU callers, opt-in random tie-break, static schedule under graph capture, the
sharded mean (single rank and 2-rank NCCL), and the coverage holes of round 1 (lambda losses at the
header's maximum list size, ARP1 / NDCG1 at L = 2048, full-size properties at (4096, 1024) and a
sample of (65536, 512)).  Same tolerance contract as test_gpu_parity.py.
"""
import os
import socket

import numpy as np
import pytest
import torch

import oracle
from test_gpu_parity import (ADDITIVE, LAMBDA, _assert_parity, _loss_module, _oracle_loss, _run_cuda,
                             make_batch)

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _cuda_loss_grad(mode, s, y, n):
    st = torch.as_tensor(s).to(DEV).requires_grad_(True)
    out = _loss_module(mode)(st, y.to(DEV), torch.as_tensor(n).to(DEV))
    out.sum().backward()
    return out.detach().cpu().numpy(), st.grad.cpu().numpy()


# ------------------------------------------------------------------ narrow relevance dtypes
@pytest.mark.parametrize("mode", ("ndcg2", "arp1", "logistic", "hinge", "dcg_hinge", "listnet"))
@pytest.mark.parametrize("L", (37, 128, 300, 512, 1100))
@pytest.mark.parametrize("dtype", (torch.uint8, torch.int16, torch.int32))
def test_narrow_relevance_is_bit_identical_to_int64(mode, L, dtype):
    """uint8 / int16 / int32 relevance (extension) takes the same kernels: results identical to int64,
    on the vector, TMA and scalar row paths (L % 16 decides TMA for 1-byte labels)."""
    s, y, n = make_batch(11 + L, 9, L)
    ref = _cuda_loss_grad(mode, s, torch.as_tensor(y), n)
    got = _cuda_loss_grad(mode, s, torch.as_tensor(y).to(dtype), n)
    if L > 1024 and mode in ("ndcg2", "arp1", "logistic"):
        # the 128 x 128 tile kernel merges its tiles with shared float atomics: not bit-reproducible
        _assert_parity(got[0].astype(np.float64), got[1].astype(np.float64), ref[0].astype(np.float64),
                       ref[1].astype(np.float64))
    else:
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])


@pytest.mark.parametrize("L", (37, 200, 700))
@pytest.mark.parametrize("dtype", (torch.uint8, torch.int16))
def test_narrow_relevance_metrics(L, dtype):
    from pytorchltr_b200.evaluation import arp, dcg, ndcg
    s, y, n = make_batch(5 + L, 7, L)
    st, nt = torch.as_tensor(s).to(DEV), torch.as_tensor(n).to(DEV)
    y64, yn = torch.as_tensor(y).to(DEV), torch.as_tensor(y).to(dtype).to(DEV)
    for fn, kw in ((ndcg, {"k": 10}), (ndcg, {}), (dcg, {"k": 5}), (dcg, {}), (arp, {})):
        assert torch.equal(fn(st, y64, nt, **kw), fn(st, yn, nt, **kw))


def test_negative_grades_with_padding():
    """Negative relevance labels next to padding: a padded column must never win a pair."""
    rng = np.random.default_rng(3)
    for L in (9, 70, 128, 200, 600):
        B = 6
        s = rng.standard_normal((B, L)).astype(np.float32)
        n = rng.integers(1, L + 1, size=B)
        n[0] = L
        y = rng.integers(-3, 3, size=(B, L))
        y[np.arange(L)[None, :] >= n[:, None]] = 0
        for mode in ("arp2", "logistic", "ndcg2"):
            loss, grad = _run_cuda(mode, s, y, n)
            rl, rg = _oracle_loss(mode, s, y, n)
            try:
                _assert_parity(loss, grad, rl, rg, loss_rel=2e-5)
            except AssertionError as e:
                raise AssertionError(f"L={L} mode={mode} n={n.tolist()}: {e}") from None


def test_fractional_labels_are_rejected():
    from pytorchltr_b200.loss import LambdaARPLoss2
    s = torch.randn(2, 5, device=DEV)
    n = torch.tensor([5, 4], device=DEV)
    ok = LambdaARPLoss2()(s, torch.tensor([[0., 1., 2., 0., 1.]] * 2, device=DEV), n)
    assert ok.shape == (2,)
    with pytest.raises(ValueError):
        LambdaARPLoss2()(s, torch.tensor([[0.5, 1., 2., 0., 1.]] * 2, device=DEV), n)


# ------------------------------------------------------------------ host-buffer entry points
def test_host_path_mean_backward_matches_device_path():
    """`loss_fn(cpu tensors).mean().backward()` (the reference's own training-loop form): same numbers as
    the device path, for a batch small enough for one chunk and one large enough to be pipelined."""
    for (B, L) in ((64, 128), (16384, 512)):
        s, y, n = make_batch(3, B, L)
        mod = _loss_module("ndcg2")
        sd = torch.as_tensor(s).to(DEV).requires_grad_(True)
        ld = mod(sd, torch.as_tensor(y).to(DEV), torch.as_tensor(n).to(DEV))
        ld.mean().backward()
        for yh, nh in ((torch.as_tensor(y), torch.as_tensor(n)),
                       (torch.as_tensor(y).to(torch.uint8), torch.as_tensor(n).to(torch.int32))):
            sh = torch.as_tensor(s).pin_memory().requires_grad_(True)
            lh = mod(sh, yh.pin_memory(), nh.pin_memory())
            assert not lh.is_cuda
            lh.mean().backward()
            assert not sh.grad.is_cuda
            assert torch.equal(lh.detach(), ld.detach().cpu())
            assert torch.equal(sh.grad, sd.grad.cpu())
        # a non-uniform upstream gradient takes the generic path
        w = torch.rand(B)
        sh = torch.as_tensor(s).requires_grad_(True)
        (mod(sh, torch.as_tensor(y), torch.as_tensor(n)) * w).sum().backward()
        sd.grad = None
        (mod(sd, torch.as_tensor(y).to(DEV), torch.as_tensor(n).to(DEV)) * w.to(DEV)).sum().backward()
        assert torch.equal(sh.grad, sd.grad.cpu())


def test_scale_rows_host_and_loss_host_ex_c_abi():
    from pytorchltr_b200 import _lib
    lib = _lib.lib()
    B, L = 9000, 1024     # 9000 * 1024 * 12 B = 110 MB: seven chunks, the last one partial
    s, y, n = make_batch(8, B, L)
    hs, hy, hn = (torch.as_tensor(a).pin_memory() for a in (s, y.astype(np.int16), n.astype(np.int32)))
    hl = torch.empty(B, pin_memory=True)
    hg = torch.empty(B, L, pin_memory=True)
    ws = torch.empty(lib.ltr_host_workspace_bytes(B, L), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ltr_loss_host_ex(_lib.FAMILY_ADDITIVE, _lib.ADD_LOGISTIC, hs.data_ptr(), hy.data_ptr(), 2,
                                    hn.data_ptr(), 4, B, L, 1.0, hl.data_ptr(), hg.data_ptr(), 1, ws.data_ptr(),
                                    ws.numel(), st))
    torch.cuda.synchronize()
    loss, grad = _run_cuda("logistic", s, y, n)
    assert np.array_equal(hl.numpy(), loss.astype(np.float32))
    assert np.array_equal(hg.numpy(), grad.astype(np.float32))
    off = lib.ltr_host_workspace_dscores_offset(B, L)
    gd = ws[off:off + 4 * B * L].view(torch.float32).view(B, L)
    h2 = torch.empty(B, L, pin_memory=True)
    scratch = torch.empty_like(gd)
    _lib.check(lib.ltr_scale_rows_host(0.25, gd.data_ptr(), h2.data_ptr(), B, L, scratch.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(h2.numpy(), 0.25 * hg.numpy())
    _lib.check(lib.ltr_scale_rows_host(1.0, gd.data_ptr(), h2.data_ptr(), B, L, None, st))
    torch.cuda.synchronize()
    assert np.array_equal(h2.numpy(), hg.numpy())
    assert lib.ltr_loss_host_ex(_lib.FAMILY_ADDITIVE, 0, hs.data_ptr(), hy.data_ptr(), 3, hn.data_ptr(), 4, B, L, 1.0,
                                hl.data_ptr(), None, 0, ws.data_ptr(), ws.numel(), st) == -1


# ------------------------------------------------------------------ tie-break, capture, misc
def test_rank_by_score_random_tiebreak_is_opt_in():
    from pytorchltr_b200.utils import rank_by_score
    B, L = 64, 50
    s = torch.zeros(B, L, device=DEV)            # every document tied
    n = torch.full((B,), 40, device=DEV)
    det = rank_by_score(s, n)
    assert torch.equal(det, torch.arange(L, device=DEV).expand(B, L))
    g = torch.Generator(device=DEV).manual_seed(1)
    rnd = rank_by_score(s, n, generator=g)
    assert torch.equal(rnd.sort(dim=1).values, torch.arange(L, device=DEV).expand(B, L))
    assert (rnd[:, :40] < 40).all() and (rnd[:, 40:] >= 40).all()     # padding stays last
    assert not torch.equal(rnd, det)
    # tie-free scores: the permutation does not change the ranking of the valid documents
    s2 = torch.randn(B, L, device=DEV)
    a, b = rank_by_score(s2, n), rank_by_score(s2, n, generator=g)
    assert torch.equal(a[:, :40], b[:, :40])


@pytest.mark.parametrize("L", (128, 512))
def test_captured_launch_without_workspace_matches_eager(L):
    """A captured launch of the non-workspace entry point takes the queries in a static stride (no
    device-wide queue slot is baked into the graph); more queries than resident slots."""
    from pytorchltr_b200 import _lib
    lib = _lib.lib()
    B = 40000 if L == 128 else 3000
    s, y, n = make_batch(4, B, L)
    ds, dy, dn = (torch.as_tensor(a).to(DEV) for a in (s, y, n))
    loss_e, grad_e = torch.empty(B, device=DEV), torch.empty(B, L, device=DEV)
    loss_g, grad_g = torch.empty(B, device=DEV), torch.empty(B, L, device=DEV)

    def launch(lo, gr):
        _lib.check(lib.ltr_lambda(_lib.LAM_NDCG2, ds.data_ptr(), dy.data_ptr(), 8, dn.data_ptr(), 8, B, L, 1.0,
                                  lo.data_ptr(), gr.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream))

    launch(loss_e, grad_e)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        launch(loss_g, grad_g)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss_e, loss_g) and torch.equal(grad_e, grad_g)


def test_fused_linear_listnet_refuses_feature_gradients():
    from pytorchltr_b200.fused import linear_listnet
    x = torch.randn(4, 8, 16, device=DEV, requires_grad=True)
    w = torch.randn(16, device=DEV, requires_grad=True)
    y = torch.randint(0, 3, (4, 8), device=DEV)
    n = torch.full((4,), 8, device=DEV)
    with pytest.raises(NotImplementedError):
        linear_listnet(x, w, None, y, n)


# ------------------------------------------------------------------ sharded mean
def test_sharded_mean_loss_single_rank_uses_the_kernel_epilogue():
    from pytorchltr_b200.distributed import sharded_mean_loss
    s, y, n = make_batch(21, 300, 200)
    mod = _loss_module("ndcg2")
    a = torch.as_tensor(s).to(DEV).requires_grad_(True)
    b = torch.as_tensor(s).to(DEV).requires_grad_(True)
    yd, nd = torch.as_tensor(y).to(DEV), torch.as_tensor(n).to(DEV)
    m1 = sharded_mean_loss(mod, a, yd, nd)
    m1.backward()
    m2 = mod(b, yd, nd).mean()
    m2.backward()
    assert m1.item() == pytest.approx(m2.item(), rel=2e-6)      # float32 atomics: summation order
    assert torch.equal(a.grad, b.grad)
    # the caller knows the global query count: scalar-only exchange, same numbers
    c = torch.as_tensor(s).to(DEV).requires_grad_(True)
    m3 = sharded_mean_loss(mod, c, yd, nd, global_count=300)
    m3.backward()
    assert m3.item() == pytest.approx(m2.item(), rel=2e-6)
    assert torch.allclose(c.grad, b.grad, rtol=1e-6, atol=0)


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _nccl_worker(rank, world, port, B, L, out_dir):
    import traceback
    try:
        _nccl_worker_body(rank, world, port, B, L, out_dir)
    except BaseException:  # noqa: B902  (reported through a file: the parent kills a wedged peer)
        with open(os.path.join(out_dir, f"error{rank}.txt"), "w") as f:
            f.write(traceback.format_exc())
        os._exit(1)
    os._exit(0)


def _nccl_worker_body(rank, world, port, B, L, out_dir):
    import torch.distributed as dist
    from pytorchltr_b200.distributed import global_mean, shard_batch, shard_bounds, sharded_mean_loss
    from pytorchltr_b200.evaluation import ndcg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dist.all_reduce(torch.zeros(2, device=dev))          # communicator set-up before any capture
    torch.cuda.synchronize()
    s, y, n = make_batch(5, B, L)
    st, yt, nt = shard_batch(torch.as_tensor(s), torch.as_tensor(y), torch.as_tensor(n))
    st = st.to(dev).requires_grad_(True)
    yt, nt = yt.to(dev), nt.to(dev)
    mod = _loss_module("ndcg2")
    mean = sharded_mean_loss(mod, st, yt, nt)
    mean.backward()
    gm = global_mean(ndcg(st.detach(), yt, nt, k=10))
    torch.cuda.synchronize()
    # the same step captured into a CUDA graph (collective included), replayed twice
    st2 = st.detach().clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            st2.grad = None
            sharded_mean_loss(mod, st2, yt, nt).backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    st2.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        gmean = sharded_mean_loss(mod, st2, yt, nt, global_count=B)
        gmean.backward()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    # NVLink peer-memory exchange (ltr_p2p_*): same sums as NCCL, eager and captured, many steps in a row
    from pytorchltr_b200.distributed import PeerScalarExchange
    ex = PeerScalarExchange()
    p2p_ok = True
    for it in range(300):
        k = 1 + it % 4
        v = torch.arange(k, device=dev, dtype=torch.float32) * (rank + 1) + it * 0.5 + rank
        ref = v.clone()
        dist.all_reduce(ref)
        ex.all_reduce_(v)
        p2p_ok = p2p_ok and bool(torch.equal(v, ref))
    st3 = st.detach().clone().requires_grad_(True)
    with torch.cuda.stream(side):
        for _ in range(3):
            st3.grad = None
            sharded_mean_loss(mod, st3, yt, nt, global_count=B, exchange=ex).backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    st3.grad = None
    graph3 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph3):
        pmean = sharded_mean_loss(mod, st3, yt, nt, global_count=B, exchange=ex)
        pmean.backward()
    for _ in range(5):
        graph3.replay()
    torch.cuda.synchronize()
    # vector form: the same sums as NCCL at lengths around the mailbox capacity (65536), many in a row
    vec_ok = True
    for it, k in enumerate((1, 1000, 65536, 65537, 200003, 7371, 7371, 7371)):
        g = torch.Generator(device=dev).manual_seed(100 * rank + it)
        v = torch.randn(k, device=dev, generator=g)
        ref = v.clone()
        dist.all_reduce(ref)
        ex.all_reduce_vec_(v)
        vec_ok = vec_ok and bool(torch.equal(v, ref))             # two ranks: a + b in either order
    # data-parallel MLP ranker: every rank scores its shard; backward returns the gradient summed over the ranks
    # (exchange fused into the final reduction), equal to local gradient + NCCL all-reduce; eager and captured
    from pytorchltr_b200.fused import MLPRanker
    import pytorchltr_b200.loss as Lmod
    torch.manual_seed(7)                                  # the same initial model on every rank
    Fm = 136
    local = MLPRanker(Fm).to(dev)
    shared = MLPRanker(Fm, exchange=ex).to(dev)
    shared.load_state_dict(local.state_dict())
    gx = torch.Generator(device=dev).manual_seed(50 + rank)
    nq = st.shape[0]
    xs = torch.randn(nq, L, Fm, device=dev, generator=gx)
    hinge = Lmod.PairwiseHingeLoss()

    def step(model):
        for p_ in model.parameters():
            p_.grad = None
        (hinge(model(xs), yt, nt).sum() / B).backward()
        return torch.cat([p_.grad.reshape(-1) for p_ in model.parameters()])
    ref_g = step(local).clone()
    dist.all_reduce(ref_g)
    got_g = step(shared).clone()
    mlp_ok = bool(torch.equal(got_g, ref_g))
    with torch.cuda.stream(side):
        for _ in range(3):
            step(shared)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph4 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph4):
        cap_g = step(shared)
    for _ in range(3):
        graph4.replay()
    torch.cuda.synchronize()
    mlp_ok = mlp_ok and bool(torch.equal(cap_g, ref_g))
    # wide rows: gradient longer than one mailbox piece (unfused vector exchange after the reduction)
    wide_l = MLPRanker(1400).to(dev)
    wide_s = MLPRanker(1400, exchange=ex).to(dev)
    wide_s.load_state_dict(wide_l.state_dict())
    xw = torch.randn(64, 16, 1400, device=dev, generator=gx)
    gw = torch.randn(64, 16, 1, device=dev, generator=gx)
    wide_l(xw).backward(gw)
    wide_s(xw).backward(gw)
    rw = torch.cat([p_.grad.reshape(-1) for p_ in wide_l.parameters()])
    dist.all_reduce(rw)
    mlp_ok = mlp_ok and bool(torch.equal(torch.cat([p_.grad.reshape(-1) for p_ in wide_s.parameters()]), rw))
    p2p_ok = p2p_ok and not ex.timed_out()
    lo, hi = shard_bounds(B, rank, world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), mean=mean.item(), grad=st.grad.cpu().numpy(),
             gmean=gmean.item(), ggrad=st2.grad.cpu().numpy(), lo=lo, hi=hi, gm=gm.item(),
             p2p_ok=p2p_ok, pmean=pmean.item(), pgrad=st3.grad.cpu().numpy(), vec_ok=vec_ok, mlp_ok=mlp_ok)
    dist.barrier()
    torch.cuda.synchronize()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_mean_loss_two_ranks_nccl(tmp_path):
    import time
    import torch.multiprocessing as mp
    B, L, world = 1001, 256, 2
    ctx = mp.start_processes(_nccl_worker, args=(world, _free_port(), B, L, str(tmp_path)), nprocs=world,
                             join=False, start_method="spawn")
    deadline = time.time() + 150
    while time.time() < deadline and any(p.is_alive() for p in ctx.processes):
        time.sleep(0.5)
    wedged = [p for p in ctx.processes if p.is_alive()]
    for p in wedged:
        p.kill()
    errors = [open(tmp_path / f).read() for f in os.listdir(tmp_path) if f.startswith("error")]
    assert not errors, errors[0]
    assert not wedged, "a rank did not finish within 150 s"
    s, y, n = make_batch(5, B, L)
    loss, grad = oracle.lambda_loss("ndcg2", s, y, n)
    ndcg = oracle.ndcg(s, y, n, k=10)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        assert float(d["mean"]) == pytest.approx(loss.mean(), rel=1e-5)
        assert float(d["gmean"]) == pytest.approx(loss.mean(), rel=1e-5)
        assert float(d["gm"]) == pytest.approx(ndcg.mean(), rel=1e-5)
        ref = grad[lo:hi] / B
        gmax = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(d["grad"] - ref) <= 1e-5 * gmax + 1e-9).all()
        assert np.allclose(d["grad"], d["ggrad"], rtol=1e-6, atol=0)
        assert bool(d["p2p_ok"]), "peer-memory all-reduce disagrees with NCCL (or timed out)"
        assert bool(d["vec_ok"]), "peer-memory vector all-reduce disagrees with NCCL"
        assert bool(d["mlp_ok"]), "MLPRanker(exchange=...) gradients differ from local gradients + NCCL all-reduce"
        assert float(d["pmean"]) == pytest.approx(loss.mean(), rel=1e-5)
        assert np.array_equal(d["pgrad"], d["ggrad"])


# ------------------------------------------------------------------ round-1 coverage holes
@pytest.mark.parametrize("mode", LAMBDA + ("logistic",))
def test_lambda_losses_at_maximum_list_size(mode):
    """L = 4096 = LTR_MAX_LIST_SIZE (pair_cta_kernel) vs the oracle on B = 1 (+ a short query)."""
    s, y, n = make_batch(9, 2, 4096)
    n[0], n[1] = 4096, 2500
    y[1, 2500:] = 0
    loss, grad = _run_cuda(mode, s, y, n)
    rl, rg = _oracle_loss(mode, s, y, n)
    _assert_parity(loss, grad, rl, rg, loss_rel=2e-5)


@pytest.mark.parametrize("mode", ("arp1", "ndcg1"))
def test_ordered_pair_losses_at_2048(mode):
    s, y, n = make_batch(10, 1, 2048, full=True)
    loss, grad = _run_cuda(mode, s, y, n)
    rl, rg = _oracle_loss(mode, s, y, n)
    _assert_parity(loss, grad, rl, rg, loss_rel=2e-5)


def test_full_size_properties_north_star_and_c5_sample():
    """(4096, 1024) and (65536, 512): per-query gradient sums to zero, padded scores are irrelevant
    (bit-exact), and a 64-query sample agrees with the oracle."""
    mod = _loss_module("ndcg2")
    for (B, L, seed) in ((4096, 1024, 31), (65536, 512, 32)):
        s, y, n = make_batch(seed, B, L)
        st = torch.as_tensor(s).to(DEV).requires_grad_(True)
        yt, nt = torch.as_tensor(y).to(DEV), torch.as_tensor(n).to(DEV)
        out = mod(st, yt, nt)
        out.sum().backward()
        g = st.grad
        gsum = g.double().sum(dim=1).abs()
        gabs = g.double().abs().sum(dim=1)
        assert (gsum <= 2e-5 * gabs + 1e-9).all()
        pad = torch.arange(L, device=DEV)[None, :] >= nt[:, None]
        assert (g[pad] == 0).all()
        s2 = st.detach().clone()
        s2[pad] = 123.0
        s2.requires_grad_(True)
        out2 = mod(s2, yt, nt)
        out2.sum().backward()
        assert torch.equal(out2, out) and torch.equal(s2.grad, g)
        idx = np.linspace(0, B - 1, 64).astype(np.int64)
        rl, rg = oracle.lambda_loss("ndcg2", s[idx], y[idx], n[idx])
        _assert_parity(out.detach().cpu().double().numpy()[idx], g.cpu().double().numpy()[idx], rl, rg)


def test_ours_vs_ref32_beside_ref32_vs_ref64():
    """the reference's own float32 result drifts from its float64 result (cancellation in
    log2(sigmoid ** tiny)); ours must sit closer to ref64 than ref32 does, or within the contract."""
    import glob
    rows = []
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz"))):
        g = dict(np.load(path))
        for mode in LAMBDA + ADDITIVE:
            if f"{mode}_loss32" not in g:
                continue
            loss, _ = _run_cuda(mode, g["scores"], g["relevance"], g["n"], float(g["sigma"]))
            ref64, ref32 = g[f"{mode}_loss64"], g[f"{mode}_loss32"].astype(np.float64)
            scale = np.abs(ref64) + 1e-6
            ours = float((np.abs(loss - ref64) / scale).max())
            theirs = float((np.abs(ref32 - ref64) / scale).max())
            o32 = float((np.abs(loss - ref32) / scale).max())
            rows.append((os.path.basename(path), mode, ours, theirs, o32))
            assert ours <= max(1e-5, theirs), (path, mode, ours, theirs)
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "ref32_vs_ref64.txt"), "w") as f:
            f.write("fixture mode max_rel|ours-ref64| max_rel|ref32-ref64| max_rel|ours-ref32|\n")
            for r in rows:
                f.write("%s %s %.3e %.3e %.3e\n" % r)


# ------------------------------------------------------------------ close scores: exactness of the packed sort
@pytest.mark.parametrize("L", (20, 50, 100, 200, 256))
@pytest.mark.parametrize("kind", ("bucket", "grid", "pairs"))
def test_rank_by_score_with_scores_closer_than_the_packed_key(L, kind):
    """Scores that share the key bits of the 32-bit packed sort: the ranking must still be the exact
    (descending score, lowest index first) order, and arp / dcg must agree with the oracle."""
    from pytorchltr_b200.evaluation import arp, dcg
    from pytorchltr_b200.utils import rank_by_score
    rng = np.random.default_rng(L)
    B = 33
    if kind == "bucket":       # every score inside one bucket, random order, with exact ties
        s = (1.0 + rng.integers(0, 64, size=(B, L)) * 2.0 ** -22).astype(np.float32)
    elif kind == "grid":       # normal scores on a coarse grid: many exact ties
        s = (np.round(rng.standard_normal((B, L)) * 8) / 8).astype(np.float32)
    else:                      # random scores, a few pairs one ulp apart in reversed index order
        s = rng.standard_normal((B, L)).astype(np.float32)
        for b in range(B):
            i, j = sorted(rng.choice(L // 2, size=2, replace=False))
            s[b, i] = np.nextafter(s[b, j], np.float32(-np.inf))
    n = rng.integers(L // 2, L + 1, size=B)
    y = rng.integers(0, 5, size=(B, L))
    y[np.arange(L)[None, :] >= n[:, None]] = 0
    st, yt, nt = (torch.as_tensor(a).to(DEV) for a in (s, y, n))
    got = rank_by_score(st, nt).cpu().numpy()
    for b in range(B):
        nb = int(n[b])
        want = np.lexsort((np.arange(nb), -s[b, :nb].astype(np.float64)))
        assert np.array_equal(got[b, :nb], want), (kind, L, b)
        assert np.array_equal(got[b, nb:], np.arange(nb, L))
    assert arp(st, yt, nt).cpu().double().numpy() == pytest.approx(oracle.arp(s, y, n), rel=1e-5, abs=1e-6)
    assert dcg(st, yt, nt).cpu().double().numpy() == pytest.approx(oracle.dcg(s, y, n), rel=1e-5, abs=1e-5)
