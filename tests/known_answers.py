"""Known-answer vectors that pin the hot path.

Every case below restates a known answer held by the reference's own tests or
docs (cited per block, paths relative to the reference root).  The same cases
are run against the CPU oracle (``-m "not gpu"``) and, through the C-ABI, against
the CUDA path (``-m gpu``).  ``approx`` tolerances are pytest's defaults
(rel 1e-6, abs 1e-12) unless the reference test states another one.
"""
from math import exp, log, log2

import numpy as np

F = np.float32
I = np.int64


def _case(kind, mode, scores, rel, n, loss, sigma=1.0, grad=None, abs_tol=None, src=""):
    return dict(kind=kind, mode=mode, sigma=sigma,
                scores=np.asarray(scores, dtype=F), rel=np.asarray(rel, dtype=I),
                n=np.asarray(n, dtype=I), loss=np.asarray(loss, dtype=np.float64),
                grad=None if grad is None else np.asarray(grad, dtype=np.float64),
                abs_tol=abs_tol, src=src)


LOSS_CASES = []

# --- tests/loss/test_pairwise_additive.py ------------------------------------
_YS = [[0, 0, 1, 2, 1]]
LOSS_CASES += [
    _case("additive", "hinge", [[0.0, 0.0, 1.0, 2.0, 1.0]], _YS, [5], [0.0],
          src="test_pairwise_additive.py:33-41"),
    _case("additive", "hinge", [[0.0, 0.0, 1.0, 1.0, 1.0]], _YS, [5], [2.0],
          src="test_pairwise_additive.py:44-52"),
    _case("additive", "hinge", [[0.0, 0.0, 1.0, -5.0, 1.0]], _YS, [5], [7.0 + 7.0 + 6.0 + 6.0],
          src="test_pairwise_additive.py:55-63"),
    _case("additive", "hinge",
          [[0.0, 10.0, 1.0, 0.5, 1.0], [1.0, 3.5, 6.0, 4.3, 10.0]],
          [[0, 2, 1, 2, 1], [1, 2, 2, 1, 0]], [5, 4], [0.5 + 1.5 + 1.5, 5.3 - 3.5],
          src="test_pairwise_additive.py:66-79"),
    _case("additive", "hinge", [[1.0, 3.5, 6.0, 4.3, 8.0]], [[1, 2, 2, 1, 0]], [3], [0.0],
          src="test_pairwise_additive.py:82-97"),
    _case("additive", "hinge", [[1.0, 3.5, 6.0, 4.3, 8.0]], [[1, 2, 2, 1, 0]], [4], [5.3 - 3.5],
          src="test_pairwise_additive.py:82-97"),
    _case("additive", "hinge", [[1.0, 3.5, 6.0, 4.3, 8.0]], [[1, 2, 2, 1, 0]], [5],
          [(5.3 - 3.5) + (9.0 - 4.3) + (9.0 - 6.0) + (9.0 - 3.5) + (9.0 - 1.0)],
          src="test_pairwise_additive.py:82-97"),
    _case("additive", "dcg_hinge", [[0.0, 0.0, 1.0, 2.0, 1.0]], _YS, [5], [-1.0 / log(2.0)],
          src="test_pairwise_additive.py:100-109"),
    _case("additive", "dcg_hinge", [[3.0, 3.0, 1.0, 0.0, 1.0]], _YS, [5], [-1.0 / log(26.0)],
          src="test_pairwise_additive.py:112-121"),
    _case("additive", "logistic", [[0.0, 0.0, 1.0, 2.0, 1.0]], _YS, [5],
          [log2(1.0 + exp(-1.0)) * 2 + log2(1.0 + exp(-2.0)) * 2 + log2(1.0 + exp(-1.0)) * 4],
          src="test_pairwise_additive.py:124-139"),
    _case("additive", "logistic", [[3.0, 3.0, 1.0, 0.0, 1.0]], _YS, [5],
          [log2(1.0 + exp(3.0)) * 2 + log2(1.0 + exp(1.0)) * 2 + log2(1.0 + exp(2.0)) * 4],
          src="test_pairwise_additive.py:142-157"),
]

# --- docs/source/loss.rst:26-37 doctest + SURVEY 8(c) gradients captured from the
# reference's autograd on the same batch -------------------------------------
_DS = [[0.5, 2.0, 1.0], [0.9, -1.2, 0.0]]
_DY = [[2, 0, 1], [0, 1, 0]]
_DN = [3, 2]
LOSS_CASES += [
    _case("additive", "hinge", _DS, _DY, _DN, [6.0, 3.1],
          grad=[[-2, 2, 0], [1, -1, 0]], src="docs/source/loss.rst:30-37"),
    _case("additive", "dcg_hinge", _DS, _DY, _DN,
          [-1.0 / log(8.0), -1.0 / log(5.1)],
          grad=[[-0.05781581, 0.05781581, 0], [0.07386853, -0.07386853, 0]], src="SURVEY 8(c)"),
    _case("additive", "logistic", _DS, _DY, _DN, None,
          grad=[[-2.07753, 2.234205, -0.1566755], [1.285302, -1.285302, 0]], src="SURVEY 8(c)"),
    # tests/loss/test_pairwise_lambda.py:11-40
    _case("lambda", "arp1", _DS, _DY, _DN, [13.298417091369629, 4.196318626403809],
          grad=[[-3.610384, 3.413716, 0.1966677], [1.285302, -1.285301, 0]],
          src="test_pairwise_lambda.py:22-25"),
    _case("lambda", "arp2", _DS, _DY, _DN, [8.209173202514648, 3.1963188648223877],
          grad=[[-3.25704, 3.413716, -0.1566755], [1.285302, -1.285302, 0]],
          src="test_pairwise_lambda.py:27-30"),
    _case("lambda", "ndcg1", _DS, _DY, _DN, [2.629549503326416, 2.647582530975342],
          grad=[[-0.7636176, 0.6705456, 0.09307203], [0.8109351, -0.810935, 0]],
          src="test_pairwise_lambda.py:32-35"),
    _case("lambda", "ndcg2", _DS, _DY, _DN, [0.3102627396583557, 0.4184933304786682],
          grad=[[-0.1323237, 0.1055911, 0.02673253], [0.1682843, -0.1682843, 0]],
          src="test_pairwise_lambda.py:37-40"),
]


# --- tests/loss/test_pairwise_lambda.py:43-366: the reference re-derives the
# expected value with explicit Python double loops (and hard-coded sort orders).
def _arp1_expected(scores, ys, n):
    e = 0.0
    for i in range(n):
        for j in range(n):
            inner = 1.0 / (1.0 + exp(-1.0 * (scores[i] - scores[j])))
            e -= log2(inner ** float(ys[i]))
    return e


def _arp2_expected(scores, ys, n):
    e = 0.0
    for i in range(n):
        for j in range(n):
            if ys[i] > ys[j]:
                e += abs(float(ys[i] - ys[j])) * log2(1.0 + exp(-1.0 * (scores[i] - scores[j])))
    return e


def _ndcg1_expected(scores, ys, n, sorting, max_dcg):
    discounts = [log2(2.0 + i) for i in range(5)]
    gains = [((2 ** float(ys[i])) - 1.0) / max_dcg for i in range(5)]
    e = 0.0
    for i in range(n):
        for j in range(n):
            si, sj = sorting[i], sorting[j]
            inner = 1.0 + exp(-1.0 * (scores[si] - scores[sj]))
            e -= log2((1.0 / inner) ** (gains[si] / discounts[i]))
    return e


def _ndcg2_expected(scores, ys, n, sorting, max_dcg):
    discounts = [log2(2.0 + i) for i in range(6)]
    gains = [((2 ** float(ys[i])) - 1.0) / max_dcg for i in range(5)]
    e = 0.0
    for i in range(n):
        for j in range(n):
            si, sj = sorting[i], sorting[j]
            if ys[si] > ys[sj]:
                inner = 1.0 / (1.0 + exp(-1.0 * (scores[si] - scores[sj])))
                delta = abs(1.0 / discounts[abs(i - j)] - 1.0 / discounts[abs(i - j) + 1])
                e -= log2(inner ** (delta * abs(gains[si] - gains[sj])))
    return e


_Y5 = [0, 0, 1, 2, 1]
_MD3 = (2 ** 2.0 - 1.0) / log2(2.0) + (2 ** 1.0 - 1.0) / log2(3.0) + (2 ** 1.0 - 1.0) / log2(4.0)
_MD2 = (2 ** 2.0 - 1.0) / log2(2.0) + (2 ** 1.0 - 1.0) / log2(3.0)
_PERFECT = [0.0, 0.0, 10.0, 20.0, 10.0]
_WORST = [4.0, 4.0, 2.0, 0.0, 2.0]
_MID_ARP = [0.0, 1.0, 1.0, -2.0, 0.0]
_MID_NDCG = [0.0, 1.0, 1.5, -2.0, 0.0]
LOSS_CASES += [
    _case("lambda", "arp1", [[0.0, 0.0, 1.0, 2.0, 1.0]], [_Y5], [5],
          [_arp1_expected([0.0, 0.0, 1.0, 2.0, 1.0], _Y5, 5)], src="test_pairwise_lambda.py:43-78"),
    _case("lambda", "arp1", [_PERFECT], [_Y5], [5], [_arp1_expected(_PERFECT, _Y5, 5)],
          src="test_pairwise_lambda.py:81-97"),
    _case("lambda", "arp1", [_WORST], [_Y5], [5], [_arp1_expected(_WORST, _Y5, 5)],
          src="test_pairwise_lambda.py:100-116"),
    _case("lambda", "arp1", [_MID_ARP], [_Y5], [4], [_arp1_expected(_MID_ARP, _Y5, 4)],
          src="test_pairwise_lambda.py:119-135"),
    _case("lambda", "arp2", [_PERFECT], [_Y5], [5], [_arp2_expected(_PERFECT, _Y5, 5)],
          abs_tol=1e-6, src="test_pairwise_lambda.py:138-153"),
    _case("lambda", "arp2", [_WORST], [_Y5], [5], [_arp2_expected(_WORST, _Y5, 5)],
          src="test_pairwise_lambda.py:156-171"),
    _case("lambda", "arp2", [_MID_ARP], [_Y5], [4], [_arp2_expected(_MID_ARP, _Y5, 4)],
          src="test_pairwise_lambda.py:174-189"),
    _case("lambda", "ndcg1", [_PERFECT], [_Y5], [5],
          [_ndcg1_expected(_PERFECT, _Y5, 5, [3, 4, 2, 0, 1], _MD3)],
          src="test_pairwise_lambda.py:192-217"),
    _case("lambda", "ndcg1", [_WORST], [_Y5], [5],
          [_ndcg1_expected(_WORST, _Y5, 5, [1, 0, 2, 4, 3], _MD3)],
          src="test_pairwise_lambda.py:220-245"),
    _case("lambda", "ndcg1", [_MID_NDCG], [_Y5], [4],
          [_ndcg1_expected(_MID_NDCG, _Y5, 4, [2, 1, 0, 3, 4], _MD2)],
          src="test_pairwise_lambda.py:248-270"),
    _case("lambda", "ndcg2", [_PERFECT], [_Y5], [5],
          [_ndcg2_expected(_PERFECT, _Y5, 5, [3, 4, 2, 0, 1], _MD3)], abs_tol=1e-7,
          src="test_pairwise_lambda.py:273-303"),
    _case("lambda", "ndcg2", [_WORST], [_Y5], [5],
          [_ndcg2_expected(_WORST, _Y5, 5, [1, 0, 2, 4, 3], _MD3)],
          src="test_pairwise_lambda.py:306-336"),
    _case("lambda", "ndcg2", [_MID_NDCG], [_Y5], [4],
          [_ndcg2_expected(_MID_NDCG, _Y5, 4, [2, 1, 0, 3, 4], _MD2)],
          src="test_pairwise_lambda.py:339-366"),
]

# --- tests/evaluation/test_dcg.py, test_arp.py, docs/source/evaluation.rst -----
_MS = np.asarray([[10.0, 5.0, 2.0, 3.0, 4.0], [5.0, 6.0, 4.0, 2.0, 5.5]], dtype=F)
_MY = np.asarray([[0, 1, 1, 0, 1], [3, 1, 0, 1, 0]], dtype=I)
_MN = np.asarray([5, 4], dtype=I)
_ONES = np.ones((2, 5), dtype=I)
_ZEROS = np.zeros((2, 5), dtype=I)
_ALLREL = float(sum(1.0 / log2(2.0 + i) for i in range(5)))


def _m(metric, scores, rel, n, expected, k=None, exp=True, src=""):
    return dict(metric=metric, scores=np.asarray(scores, dtype=F), rel=np.asarray(rel, dtype=I),
                n=np.asarray(n, dtype=I), k=k, exp=exp,
                expected=np.asarray(expected, dtype=np.float64), src=src)


METRIC_CASES = [
    _m("dcg", _MS, _MY, _MN, [1.1309297535714575, 5.4165082750002025], k=3, src="test_dcg.py:20-30"),
    _m("ndcg", _MS, _MY, _MN, [1.1309297535714575 / 2.1309297535714578,
                               5.4165082750002025 / 8.130929753571458], k=3, src="test_dcg.py:33-43"),
    _m("dcg", _MS, _MY, _MN, [1.5177825608059992, 5.847184833073595], k=5, src="test_dcg.py:46-56"),
    _m("ndcg", _MS, _MY, _MN, [1.5177825608059992 / 2.1309297535714578,
                               5.847184833073595 / 8.130929753571458], k=5, src="test_dcg.py:59-69"),
    _m("dcg", _MS, _MY, _MN, [1.5177825608059992, 3.3234658187877653], k=5, exp=False,
       src="test_dcg.py:72-82"),
    _m("ndcg", _MS, _MY, _MN, [1.5177825608059992 / 2.1309297535714578,
                               3.3234658187877653 / 4.130929753571458], k=5, exp=False,
       src="test_dcg.py:85-95"),
    _m("dcg", _MS, _ONES, _MN, [_ALLREL, _ALLREL], k=5, exp=False, src="test_dcg.py:98-116"),
    _m("ndcg", _MS, _ONES, _MN, [1.0, 1.0], k=5, exp=False, src="test_dcg.py:119-134"),
    _m("dcg", _MS, _ZEROS, _MN, [0.0, 0.0], k=5, exp=False, src="test_dcg.py:137-152"),
    _m("ndcg", _MS, _ZEROS, _MN, [0.0, 0.0], k=5, exp=False, src="test_dcg.py:155-170"),
    _m("arp", _MS, [[0, 1, 1, 0, 1], [1, 1, 0, 0, 0]], _MN, [3.333333333, 1.5],
       src="test_arp.py:6-21"),
    _m("arp", _MS, _ONES, _MN, [3.0, 2.5], src="test_arp.py:24-39"),
    _m("arp", _MS, _ZEROS, _MN, [0.0, 0.0], src="test_arp.py:42-57"),
    _m("ndcg", [[1.0, 0.0, 1.5], [1.5, 0.2, 0.5]], [[0, 1, 0], [0, 1, 1]], [3, 3],
       [0.5, (1.0 / log2(3.0) + 0.5) / (1.0 + 1.0 / log2(3.0))], k=10, src="docs/source/evaluation.rst:17-23"),
]
