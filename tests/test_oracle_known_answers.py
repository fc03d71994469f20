"""Pins the CPU oracle against the reference's own known-answer vectors (CPU only)."""
import numpy as np
import pytest
from pytest import approx

import oracle
from known_answers import LOSS_CASES, METRIC_CASES


def _run_loss(case, **kw):
    if case["kind"] == "additive":
        return oracle.pairwise_additive(case["mode"], case["scores"], case["rel"], case["n"],
                                        sigma=case["sigma"], **kw)
    return oracle.lambda_loss(case["mode"], case["scores"], case["rel"], case["n"],
                              sigma=case["sigma"])


@pytest.mark.parametrize("case", LOSS_CASES, ids=lambda c: f"{c['mode']}@{c['src']}")
def test_oracle_loss_known_answer(case):
    loss, grad = _run_loss(case)
    if case["loss"].dtype != object and case["loss"].shape != ():
        kw = {} if case["abs_tol"] is None else {"abs": case["abs_tol"], "rel": 1e-6}
        assert loss == approx(case["loss"], **kw)
    if case["grad"] is not None:
        # SURVEY 8(c) gradients were printed with 7 significant digits
        assert grad == approx(case["grad"], rel=2e-6, abs=2e-7)


@pytest.mark.parametrize("case", [c for c in LOSS_CASES if c["kind"] == "additive"],
                         ids=lambda c: f"{c['mode']}@{c['src']}")
def test_oracle_additive_f32_mode_known_answer(case):
    """The float32 restatement (reference op order) hits the same known answers."""
    loss, grad = _run_loss(case, f32=True)
    if case["loss"].shape != ():
        assert loss == approx(case["loss"], rel=1e-6, abs=1e-6)
    if case["grad"] is not None:
        assert grad == approx(case["grad"], rel=2e-6, abs=2e-7)


@pytest.mark.parametrize("case", METRIC_CASES, ids=lambda c: f"{c['metric']}@{c['src']}")
def test_oracle_metric_known_answer(case):
    if case["metric"] == "arp":
        out = oracle.arp(case["scores"], case["rel"], case["n"])
    else:
        out = oracle.dcg(case["scores"], case["rel"], case["n"], k=case["k"], exp=case["exp"],
                         normalized=case["metric"] == "ndcg")
    assert out == approx(case["expected"], rel=1e-6, abs=1e-7)


def test_oracle_accepts_3d_inputs():
    # (B, L, 1) auto-reshape: test_pairwise_additive.py:11-30
    s = np.asarray([[0.0, 0.0, 1.0, 2.0, 1.0]], dtype=np.float32)
    y = np.asarray([[[0], [0], [1], [2], [1]]], dtype=np.int64)
    loss, _ = oracle.pairwise_additive("hinge", s, y, [5])
    assert loss == approx([0.0])
    loss, _ = oracle.pairwise_additive("hinge", s.reshape(1, 5, 1), y.reshape(1, 5), [5])
    assert loss == approx([0.0])


def test_oracle_edge_cases_n0_n1():
    # SURVEY 8(a) edge behaviour: n=0 -> 0; n=1 -> 0 for masked-pair losses,
    # rel_0 for ARP1 and G_0 (== 1 / D_0 = 1 when rel_0 > 0) for NDCG1.
    s = np.asarray([[0.3, -1.0, 2.0], [0.3, -1.0, 2.0]], dtype=np.float32)
    y = np.asarray([[2, 1, 0], [2, 1, 0]], dtype=np.int64)
    n = np.asarray([0, 1], dtype=np.int64)
    for mode in ("hinge", "logistic"):
        loss, grad = oracle.pairwise_additive(mode, s, y, n)
        assert loss == approx([0.0, 0.0]) and np.all(grad == 0)
    loss, _ = oracle.pairwise_additive("dcg_hinge", s, y, n)
    assert loss == approx([-1 / np.log(2.0)] * 2)
    loss, grad = oracle.lambda_loss("arp1", s, y, n)
    assert loss == approx([0.0, 2.0]) and np.all(grad == 0)
    loss, grad = oracle.lambda_loss("ndcg1", s, y, n)
    assert loss == approx([0.0, 1.0]) and np.all(grad == 0)
    for mode in ("arp2", "ndcg2"):
        loss, grad = oracle.lambda_loss(mode, s, y, n)
        assert loss == approx([0.0, 0.0]) and np.all(grad == 0)
    loss, grad = oracle.listnet(s, y, n)
    assert loss == approx([0.0, 0.0]) and np.all(grad == 0)
