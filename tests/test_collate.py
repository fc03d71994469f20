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


# ------------------------------------------------------------------ list samplers on the device
# Statistical tests restated from the reference's tests/datasets/test_list_sampler.py:65-97, 158-197: one
# collate call over 1000 copies of the same query draws 1000 independent samples.
def _single_query_dataset(relevance):
    from pytorchltr_b200.datasets import DeviceRankingDataset
    n = len(relevance)
    feats = torch.arange(n, dtype=torch.float32).reshape(n, 1).repeat(1, 4)     # feature = document index
    return DeviceRankingDataset(feats, torch.tensor(relevance), torch.tensor([0, n]))


def _grade_hist(rel_col, grades=3):
    return np.bincount(rel_col, minlength=grades)[:grades] / len(rel_col)


@pytest.mark.gpu
def test_uniform_sampler_statistics():
    from pytorchltr_b200.datasets import UniformSampler
    relevance = [0, 0, 1, 0, 0, 0, 2, 1]
    ds = _single_query_dataset(relevance)
    gen = torch.Generator(device="cuda").manual_seed(1608637542)
    for mls in (1, 5):
        b = ds.collate([0] * 1000, list_sampler=UniformSampler(mls, generator=gen))
        assert b.features.shape == (1000, mls, 4) and (b.n == mls).all()
        doc = b.features[:, :, 0].long().cpu().numpy()
        rel = b.relevance.cpu().numpy()
        assert np.array_equal(rel, np.asarray(relevance)[doc])          # relevance follows the sampled rows
        assert all(len(set(row)) == mls for row in doc)                 # without replacement
        for pos in range(mls):
            assert _grade_hist(rel[:, pos]) == pytest.approx([5 / 8, 2 / 8, 1 / 8], abs=0.05)
    # lists no longer than list_size are copied in order, whatever the sampler (svmrank.py:159-190)
    b = ds.collate([0, 0], list_sampler=UniformSampler(None, generator=gen))
    assert np.array_equal(b.features[:, :, 0].cpu().numpy(), np.tile(np.arange(8.0), (2, 1)))
    # a CPU generator (the reference's convention) works too
    b = ds.collate([0] * 4, list_sampler=UniformSampler(3, generator=torch.Generator().manual_seed(1)))
    assert b.features.shape == (4, 3, 4)


@pytest.mark.gpu
def test_balanced_relevance_sampler_statistics():
    from pytorchltr_b200.datasets import BalancedRelevanceSampler
    gen = torch.Generator(device="cuda").manual_seed(1608637542)
    ds = _single_query_dataset([0, 0, 1, 0, 0, 0, 2, 1])
    b = ds.collate([0] * 1000, list_sampler=BalancedRelevanceSampler(1, generator=gen))
    assert _grade_hist(b.relevance[:, 0].cpu().numpy()) == pytest.approx([1 / 3, 1 / 3, 1 / 3], abs=0.05)
    relevance = [0, 0, 1, 0, 0, 0, 2, 1, 0, 0, 0]
    ds = _single_query_dataset(relevance)
    b = ds.collate([0] * 1000, list_sampler=BalancedRelevanceSampler(7, generator=gen))
    doc = b.features[:, :, 0].long().cpu().numpy()
    rel = b.relevance.cpu().numpy()
    assert np.array_equal(rel, np.asarray(relevance)[doc])
    assert all(len(set(row)) == 7 for row in doc)
    expected = [[1 / 3, 1 / 3, 1 / 3]] * 3 + [[1 / 2, 1 / 2, 0.0]] * 2 + [[1.0, 0.0, 0.0]] * 2
    for pos in range(7):
        assert _grade_hist(rel[:, pos]) == pytest.approx(expected[pos], abs=0.05)


@pytest.mark.gpu
def test_sparse_collate_matches_dense_collate():
    """CSR dataset -> torch sparse (B, list_size, F) batch == the dense batch, for the default sampler and,
    with the same generator state, for the random samplers."""
    from pytorchltr_b200.datasets import BalancedRelevanceSampler, DeviceRankingDataset, UniformSampler
    rng = np.random.default_rng(0)
    Q, F = 40, 23
    counts = rng.integers(1, 60, size=Q)
    offs = np.concatenate([[0], np.cumsum(counts)])
    X = rng.standard_normal((offs[-1], F)).astype(np.float32)
    X[rng.random(X.shape) < 0.8] = 0.0
    y = rng.integers(0, 4, size=offs[-1])
    dense = DeviceRankingDataset(torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(offs))
    xs = torch.from_numpy(X).to_sparse_csr()
    sp = DeviceRankingDataset.from_csr(xs.crow_indices(), xs.col_indices(), xs.values(), F, torch.from_numpy(y),
                                       torch.from_numpy(offs))
    idx = rng.permutation(Q)[:17].tolist()
    for make in (lambda g: None, lambda g: UniformSampler(9, generator=g),
                 lambda g: BalancedRelevanceSampler(9, generator=g)):
        g1 = torch.Generator(device="cuda").manual_seed(7)
        g2 = torch.Generator(device="cuda").manual_seed(7)
        kw1 = {"max_list_size": 20} if make(g1) is None else {"list_sampler": make(g1)}
        kw2 = {"max_list_size": 20} if make(g2) is None else {"list_sampler": make(g2)}
        a, b = dense.collate(idx, **kw1), sp.collate(idx, **kw2)
        assert b.sparse and b.features.is_sparse and not a.sparse
        assert torch.equal(b.features.to_dense(), a.features)
        assert torch.equal(a.relevance, b.relevance) and torch.equal(a.n, b.n) and torch.equal(a.qid, b.qid)
