"""DCG and NDCG on the sm_100a ranking-metric kernel.

Drop-in for ``pytorchltr.evaluation.dcg`` (reference: pytorchltr/evaluation/dcg.py).
"""
from typing import Optional

import torch as _torch

from pytorchltr_b200 import _lib, _ops


def ndcg(scores: _torch.FloatTensor, relevance: _torch.LongTensor,
         n: _torch.LongTensor, k: Optional[int] = None,
         exp: Optional[bool] = True) -> _torch.FloatTensor:
    r"""Normalized DCG: ``dcg(scores) / dcg(relevance)`` with an ideal DCG of 0 replaced by
    1 (reference :8-38).

    Returns ``(B, L)`` (NDCG at every rank) when ``k`` is None, else ``(B,)`` NDCG@k.
    """
    return _ops.rank_metric(_lib.METRIC_NDCG, scores, relevance, n, k, bool(exp))


def dcg(scores: _torch.FloatTensor, relevance: _torch.LongTensor,
        n: _torch.LongTensor, k: Optional[int] = None,
        exp: Optional[bool] = True) -> _torch.FloatTensor:
    r"""Discounted cumulative gain :math:`\sum_i \text{gain}(y_{\pi_i}) / \log_2(1 + i)` with
    gain :math:`2^y - 1` (``exp=True``) or :math:`y` (reference :41-99).

    Returns ``(B, L)`` (DCG at every rank) when ``k`` is None, else ``(B,)`` DCG@k.  As in
    the reference (:85) the relevance of padded documents is not masked: pad it with zeros.
    """
    return _ops.rank_metric(_lib.METRIC_DCG, scores, relevance, n, k, bool(exp))
