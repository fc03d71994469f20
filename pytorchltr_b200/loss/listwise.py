"""ListNet (top-1) loss on the sm_100a kernel.

The task names ListNet, but the reference snapshot has no listwise loss
(pytorchltr/loss/__init__.py:1-7): this module is specified by this repository
(Cao et al. 2007, top-one probability model) and its parity is UNPINNED:

    P = softmax(relevance) and Q = log_softmax(scores) over the valid documents
    (masking idiom of utils/tensor_operations.py:81-87), loss_b = -sum_i P_i Q_i,
    d loss_b / d s = softmax(s) - P, and n_b = 0 gives 0.
"""
import torch as _torch

from pytorchltr_b200 import _lib, _ops


class ListNetLoss(_torch.nn.Module):
    r"""ListNet top-1 cross entropy with the call signature of every other loss:
    ``forward(scores, relevance, n) -> FloatTensor(B)``."""

    def __init__(self):
        super().__init__()

    def forward(self, scores: _torch.FloatTensor, relevance: _torch.LongTensor,
                n: _torch.LongTensor, loss_sum=None) -> _torch.FloatTensor:
        return _ops.fused_loss(scores, relevance, n, _lib.FAMILY_LISTNET, 0, 1.0, loss_sum)
