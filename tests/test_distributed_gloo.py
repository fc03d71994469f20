"""World-size-2 gloo tests (CPU) of the query-sharding host logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pytorchltr_b200.distributed import (global_mean, shard_batch, shard_bounds,
                                         sharded_mean_loss)


def test_shard_bounds_cover_exactly():
    for B in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(B, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_loss_fn(scores, relevance, n):
    """Stand-in for the CUDA loss on the CPU-only box: per-query loss with the oracle's
    gradient attached, so the sharding logic (not the kernel) is what is tested."""
    loss, grad = oracle.lambda_loss("ndcg2", scores.detach().numpy(), relevance.numpy(), n.numpy())
    grad_t = torch.from_numpy(grad).to(scores.dtype)

    class _F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, s):
            return torch.from_numpy(loss).to(s.dtype)

        @staticmethod
        def backward(ctx, g):
            return g[:, None] * grad_t

    return _F.apply(scores)


def _worker(rank, world, port, B, L, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        scores = torch.randn(B, L, generator=g, dtype=torch.float64)
        n = torch.randint(L // 2, L + 1, (B,), generator=g)
        rel = torch.randint(0, 5, (B, L), generator=g)
        rel[torch.arange(L)[None, :] >= n[:, None]] = 0
        s, y, nn = shard_batch(scores, rel, n)
        s = s.clone().requires_grad_(True)
        mean = sharded_mean_loss(_oracle_loss_fn, s, y, nn)
        mean.backward()
        metric = torch.from_numpy(oracle.ndcg(s.detach().float().numpy(), y.numpy(), nn.numpy(), k=10))
        gm = global_mean(metric)
        lo, hi = shard_bounds(B, rank, world)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), mean=mean.detach().numpy(), grad=s.grad.numpy(),
                 lo=lo, hi=hi, gm=gm.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_mean_loss_matches_single_process(tmp_path):
    B, L, world = 11, 24, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, L, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    scores = torch.randn(B, L, generator=g, dtype=torch.float64)
    n = torch.randint(L // 2, L + 1, (B,), generator=g)
    rel = torch.randint(0, 5, (B, L), generator=g)
    rel[torch.arange(L)[None, :] >= n[:, None]] = 0
    loss, grad = oracle.lambda_loss("ndcg2", scores.float().numpy(), rel.numpy(), n.numpy())
    ndcg = oracle.ndcg(scores.float().numpy(), rel.numpy(), n.numpy(), k=10)
    seen = 0
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert d["mean"] == pytest.approx(loss.mean(), rel=1e-12)
        assert d["gm"] == pytest.approx(ndcg.mean(), rel=1e-12)
        lo, hi = int(d["lo"]), int(d["hi"])
        assert d["grad"] == pytest.approx(grad[lo:hi] / B, rel=1e-12, abs=1e-15)
        seen += hi - lo
    assert seen == B


def test_peer_exchange_rejects_what_it_cannot_carry():
    """Host-side argument checks of the peer-memory exchange (no GPU, no process group): CPU tensors, wrong dtypes
    and empty vectors never reach the library; the exchange cannot be built without a process group."""
    from pytorchltr_b200.distributed import PeerExchange, PeerScalarExchange
    assert PeerExchange is PeerScalarExchange
    ex = object.__new__(PeerScalarExchange)
    for bad in (torch.zeros(8), torch.zeros(0), torch.zeros(8, dtype=torch.float64)):
        with pytest.raises(ValueError):
            ex.all_reduce_vec_(bad)
        with pytest.raises(ValueError):
            ex.all_reduce_(bad)
    if not (dist.is_available() and dist.is_initialized()):
        with pytest.raises(RuntimeError):
            PeerScalarExchange()


def test_mlp_ranker_takes_an_exchange_and_keeps_the_documented_state_dict():
    from pytorchltr_b200.fused import MLPRanker
    m = MLPRanker(136, exchange=None)
    assert m.exchange is None and sorted(m.state_dict()) == ["l1.bias", "l1.weight", "l2.bias", "l2.weight",
                                                              "l3.bias", "l3.weight"]
    marker = object()
    assert MLPRanker(46, hidden=(20, 5), exchange=marker).exchange is marker
