"""GPU collation against the reference's own collate_fn: tests/golden/collate_ref.npz
holds the batches the UNMODIFIED SVMRankDataset.collate_fn(ListSampler(max_list_size)) produced for a
seeded ragged dataset (tests/golden/make_collate_golden.py).  Copies are bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "collate_ref.npz")
KEYS = [f"{name}_{mls}" for mls in (None, 7, 40) for name in ("all", "some", "one", "rep")]


def _items(g):
    offs = np.concatenate([[0], np.cumsum(g["counts"])])
    return [(g["features"][offs[q]:offs[q + 1]], g["relevance"][offs[q]:offs[q + 1]], int(g["qids"][q]))
            for q in range(len(g["counts"]))], offs


@pytest.mark.parametrize("key", KEYS)
def test_oracle_collate_matches_reference_collate_fn(key):
    g = np.load(GOLDEN)
    items, _ = _items(g)
    mls = None if key.endswith("None") else int(key.rsplit("_", 1)[1])
    f, r, n, q = oracle.collate(items, g[f"{key}_idx"].tolist(), mls)
    assert np.array_equal(f, g[f"{key}_features"])
    assert np.array_equal(r, g[f"{key}_relevance"])
    assert np.array_equal(n, g[f"{key}_n"])
    assert np.array_equal(q, g[f"{key}_qid"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_cuda_collate_matches_reference_collate_fn(key):
    from pytorchltr_b200.datasets import DeviceRankingDataset
    g = np.load(GOLDEN)
    _, offs = _items(g)
    ds = DeviceRankingDataset(torch.from_numpy(g["features"]), torch.from_numpy(g["relevance"]),
                              torch.from_numpy(offs), torch.from_numpy(g["qids"]))
    mls = None if key.endswith("None") else int(key.rsplit("_", 1)[1])
    b = ds.collate(g[f"{key}_idx"].tolist(), mls)
    assert b.features.is_cuda and not b.sparse
    assert np.array_equal(b.features.cpu().numpy(), g[f"{key}_features"])
    assert np.array_equal(b.relevance.cpu().numpy(), g[f"{key}_relevance"])
    assert np.array_equal(b.n.cpu().numpy(), g[f"{key}_n"])
    assert np.array_equal(b.qid.cpu().numpy(), g[f"{key}_qid"])


@pytest.mark.gpu
@pytest.mark.parametrize("Q,F,maxn", [(300, 136, 220), (50, 10, 33), (40, 4, 3000)])
def test_cuda_collate_random_vs_oracle_and_feeds_the_losses(Q, F, maxn):
    """Larger ragged sets (incl. F % 4 != 0 and rows longer than one slab) against the oracle, and the
    batch goes straight into a loss (the (features, relevance, n) the path consumes)."""
    from pytorchltr_b200.datasets import DeviceRankingDataset
    from pytorchltr_b200.fused import LinearListNet
    rng = np.random.default_rng(Q)
    counts = rng.integers(1, maxn + 1, size=Q)
    offs = np.concatenate([[0], np.cumsum(counts)])
    X = rng.standard_normal((offs[-1], F)).astype(np.float32)
    y = rng.integers(0, 5, size=offs[-1])
    items = [(X[offs[q]:offs[q + 1]], y[offs[q]:offs[q + 1]], q) for q in range(Q)]
    ds = DeviceRankingDataset(torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(offs))
    idx = rng.permutation(Q)[: max(1, Q // 2)].tolist()
    for mls in (None, 64):
        b = ds.collate(idx, mls)
        f, r, n, q = oracle.collate(items, idx, mls)
        assert np.array_equal(b.features.cpu().numpy(), f)
        assert np.array_equal(b.relevance.cpu().numpy(), r)
        assert np.array_equal(b.n.cpu().numpy(), n)
        assert np.array_equal(b.qid.cpu().numpy(), q)
    model = LinearListNet(F).to(b.features.device)
    loss = model(b.features, b.relevance, b.n)
    loss.mean().backward()
    assert torch.isfinite(loss).all() and torch.isfinite(model.linear.weight.grad).all()
