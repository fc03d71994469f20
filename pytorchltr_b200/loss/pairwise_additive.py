"""Pairwise additive ranking losses on the fused sm_100a kernel.

Drop-in for ``pytorchltr.loss.pairwise_additive`` (reference:
pytorchltr/loss/pairwise_additive.py): same class names, constructor arguments and
``forward(scores, relevance, n) -> FloatTensor(B)``.  The reference materialises
``(B, L, L, 2)`` pair tensors (:68-69) and masks them (:75-81); here the pairs are
generated on chip by ``ltr_pairwise_additive`` and the gradient is produced by the
same launch.
"""
import torch as _torch

from pytorchltr_b200 import _lib, _ops


class _PairwiseAdditiveLoss(_torch.nn.Module):
    """Linearly decomposable additive pairwise losses (reference :5-90)."""

    _mode = None

    def __init__(self):
        super().__init__()

    def _sigma(self) -> float:
        return 1.0

    def forward(self, scores: _torch.FloatTensor, relevance: _torch.LongTensor,
                n: _torch.LongTensor, loss_sum=None) -> _torch.FloatTensor:
        """Computes the per-query loss for a padded batch.

        Args:
            scores: ``(B, L)`` or ``(B, L, 1)`` scores.
            relevance: ``(B, L)`` or ``(B, L, 1)`` integer relevance labels.
            n: ``(B,)`` number of documents per query; documents ``>= n`` are padding.
            loss_sum: extension (not in the reference): float32 CUDA scalar to which the kernel
                adds the sum of the per-query losses (see ``pytorchltr_b200.distributed``).
        """
        return _ops.fused_loss(scores, relevance, n, _lib.FAMILY_ADDITIVE, self._mode,
                               self._sigma(), loss_sum)


class PairwiseHingeLoss(_PairwiseAdditiveLoss):
    r"""RankSVM hinge loss :math:`\sum_{y_i > y_j} \max(0, 1 - (s_i - s_j))`
    (reference :93-113)."""
    _mode = _lib.ADD_HINGE


class PairwiseDCGHingeLoss(PairwiseHingeLoss):
    r"""DCG-modified hinge loss :math:`-1 / \ln(2 + \text{hinge})` (reference :116-133)."""
    _mode = _lib.ADD_DCG_HINGE


class PairwiseLogisticLoss(_PairwiseAdditiveLoss):
    r"""RankNet logistic loss :math:`\sum_{y_i > y_j} \log_2(1 + e^{-\sigma (s_i - s_j)})`
    (reference :136-163)."""
    _mode = _lib.ADD_LOGISTIC

    def __init__(self, sigma: float = 1.0):
        """
        Args:
            sigma: Steepness of the logistic curve.
        """
        super().__init__()
        self.sigma = sigma

    def _sigma(self) -> float:
        return float(self.sigma)
