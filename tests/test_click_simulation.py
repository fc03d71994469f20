"""PBM click simulation (click_simulation/pbm.py): the oracle and the CUDA kernel
against the reference's own known answers (tests/click_simulation/test_pbm.py:8-140, restated below
with their source lines) and against outputs of the unmodified reference
(tests/golden/pbm_ref.npz, tests/golden/make_pbm_golden.py)."""
import os

import numpy as np
import pytest
import torch
from pytest import approx

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pbm_ref.npz")

# tests/click_simulation/test_pbm.py:9-21
RANKINGS = np.array([[3, 4, 0, 2, 1], [1, 0, 2, 4, 3]])
YS = np.array([[1, 0, 4, 0, 2], [4, 3, 0, 0, 0]])
N = np.array([5, 3])
PERFECT = [0.0, 0.2, 0.4, 0.8, 1.0]          # pbm.py:81-82
POSITION = [0.1, 0.1, 0.1, 1.0, 1.0]         # pbm.py:103-104
REL_PERFECT = np.array([[0.2, 0.0, 1.0, 0.0, 0.4], [1.0, 0.8, 0.0, 0.0, 0.0]])      # test_pbm.py:49-52
REL_POSITION = np.array([[0.1, 0.1, 1.0, 0.1, 0.1], [1.0, 1.0, 0.1, 0.1, 0.1]])     # test_pbm.py:100-103
PROPS_POSITION = np.array([[1 / 4, 1 / 6, 1 / 5, 1 / 2, 1 / 3], [1 / 3, 1 / 2, 1 / 4, 0.0, 0.0]])   # :104-107
KNOWN = [
    # (relevance probs, cutoff, eta, expected propensities, expected relevance term, source)
    (PERFECT, None, 0.0, np.array([[1.0] * 5, [1, 1, 1, 0, 0]]), REL_PERFECT, "test_pbm.py:46-57"),
    (PERFECT, 3, 0.0, np.array([[1, 0, 0, 1, 1], [1, 1, 1, 0, 0]]), REL_PERFECT, "test_pbm.py:60-76"),
    (PERFECT, 2, 0.0, np.array([[0, 0, 0, 1, 1], [1, 1, 0, 0, 0]]), REL_PERFECT, "test_pbm.py:79-95"),
    (POSITION, None, 1.0, PROPS_POSITION, REL_POSITION, "test_pbm.py:98-110"),
    (POSITION, None, 2.0, PROPS_POSITION ** 2.0, REL_POSITION, "test_pbm.py:113-128"),
]


@pytest.mark.parametrize("case", KNOWN, ids=lambda c: c[5])
def test_oracle_pbm_known_answers(case):
    probs, cutoff, eta, props, rel, _ = case
    cp, pr = oracle.pbm_probabilities(RANKINGS, YS, N, probs, cutoff, eta)
    assert pr == approx(props, abs=1e-12)
    assert cp == approx(rel * props, abs=1e-12)


def _golden_cases():
    g = np.load(GOLDEN)
    for name in ("small", "mid", "long"):
        for cutoff in (None, 3, 10):
            for eta in (0.0, 1.0, 2.0):
                yield g, name, cutoff, eta


def test_oracle_pbm_matches_reference_outputs():
    for g, name, cutoff, eta in _golden_cases():
        _, pr = oracle.pbm_probabilities(g[f"{name}_rankings"], g[f"{name}_ys"], g[f"{name}_n"],
                                         [0.05, 0.3, 0.5, 0.7, 0.95], cutoff, eta)
        assert pr == approx(g[f"{name}_props_c{cutoff}_e{eta}"], rel=1e-6, abs=1e-9)
    g = np.load(GOLDEN)
    for name in ("small", "mid", "long"):
        for cutoff in (None, 5):
            cp, _ = oracle.pbm_probabilities(g[f"{name}_rankings"], g[f"{name}_ys"], g[f"{name}_n"],
                                             [0.0, 1.0, 0.0, 1.0, 1.0], cutoff, 0.0)
            assert np.array_equal(cp.astype(np.int64), g[f"{name}_clicks_c{cutoff}"])


@pytest.mark.gpu
def test_cuda_pbm_matches_reference_outputs_and_oracle():
    from pytorchltr_b200.click_simulation.pbm import pbm_probabilities, simulate_pbm
    dev = torch.device("cuda", 0)
    for g, name, cutoff, eta in _golden_cases():
        rk, ys, n = (torch.from_numpy(g[f"{name}_{k}"]).to(dev) for k in ("rankings", "ys", "n"))
        probs = torch.tensor([0.05, 0.3, 0.5, 0.7, 0.95])
        cp, pr = pbm_probabilities(rk, ys, n, probs, cutoff, eta)
        assert pr.cpu().numpy() == approx(g[f"{name}_props_c{cutoff}_e{eta}"], rel=2e-6, abs=1e-9)
        ocp, _ = oracle.pbm_probabilities(g[f"{name}_rankings"], g[f"{name}_ys"], g[f"{name}_n"],
                                          probs.numpy(), cutoff, eta)
        assert cp.cpu().numpy() == approx(ocp, rel=2e-6, abs=1e-9)
    g = np.load(GOLDEN)
    for name in ("small", "mid", "long"):
        rk, ys, n = (torch.from_numpy(g[f"{name}_{k}"]).to(dev) for k in ("rankings", "ys", "n"))
        for cutoff in (None, 5):
            clicks, _ = simulate_pbm(rk, ys.int(), n.int(), torch.tensor([0.0, 1.0, 0.0, 1.0, 1.0]), cutoff, 0.0)
            assert clicks.dtype == torch.int64
            assert np.array_equal(clicks.cpu().numpy(), g[f"{name}_clicks_c{cutoff}"])


@pytest.mark.gpu
def test_cuda_pbm_monte_carlo_like_the_reference_tests():
    """tests/click_simulation/test_pbm.py:24-43: averages over repeated simulations approach the
    expected click rates and propensities (abs 0.1 there; 0.05 here with 400 runs)."""
    from pytorchltr_b200.click_simulation import simulate_perfect, simulate_position
    dev = torch.device("cuda", 0)
    rk, ys, n = (torch.from_numpy(a).to(dev) for a in (RANKINGS, YS, N))
    torch.manual_seed(4200)
    for fn, kwargs, rel, props in ((simulate_perfect, {}, REL_PERFECT, np.array([[1.0] * 5, [1, 1, 1, 0, 0]])),
                                   (simulate_position, {}, REL_POSITION, PROPS_POSITION),
                                   (simulate_position, {"eta": 2.0}, REL_POSITION, PROPS_POSITION ** 2.0)):
        acc = torch.zeros(2, 5, device=dev)
        for _ in range(400):
            clicks, pr = fn(rk, ys, n, **kwargs)
            acc += clicks.float()
        assert (acc / 400).cpu().numpy() == approx(rel * props, abs=0.05)
        assert pr.cpu().numpy() == approx(props, rel=1e-5, abs=1e-7)
    # the ranking may come straight from the ranking kernel
    from pytorchltr_b200.utils import rank_by_score
    scores = torch.randn(2, 5, device=dev)
    clicks, pr = simulate_perfect(rank_by_score(scores, n), ys, n)
    assert clicks.shape == (2, 5) and bool((pr[1].sum() == 3.0).item())
