"""Average relevant position on the sm_100a ranking-metric kernel.

Drop-in for ``pytorchltr.evaluation.arp`` (reference: pytorchltr/evaluation/arp.py).
"""
import torch as _torch

from pytorchltr_b200 import _lib, _ops


def arp(scores: _torch.FloatTensor, relevance: _torch.LongTensor,
        n: _torch.LongTensor) -> _torch.FloatTensor:
    r"""ARP :math:`\frac{1}{\sum_i y_i} \sum_i y_{\pi_i} \cdot i`; queries without relevant
    documents give 0 (reference :7-42).  Returns ``(B,)``."""
    return _ops.rank_metric(_lib.METRIC_ARP, scores, relevance, n, None, True)
