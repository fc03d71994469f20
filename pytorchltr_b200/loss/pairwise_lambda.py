"""LambdaLoss family on the fused sm_100a kernel.

Drop-in for ``pytorchltr.loss.pairwise_lambda`` (reference:
pytorchltr/loss/pairwise_lambda.py): ``LambdaARPLoss1/2`` and ``LambdaNDCGLoss1/2``
with ``forward(scores, relevance, n) -> FloatTensor(B)``.  Ranking (:66), gathers
(:67-70), gains / max-DCG (:221-241), the delta table (:206-211), the O(L^2) pair
weights and the gradient all happen inside one ``ltr_lambda`` launch.

Difference from the reference, on purpose: ties in the ranking are broken
lowest-index-first instead of by a random permutation of the global torch RNG
(utils/tensor_operations.py:43-45), so a call consumes no RNG state and is
deterministic.
"""
import torch as _torch

from pytorchltr_b200 import _lib, _ops


class LambdaLoss(_torch.nn.Module):
    """LambdaLoss template (reference :6-92)."""

    _mode = None

    def __init__(self, sigma: float = 1.0):
        """
        Args:
            sigma: Steepness of the logistic curve.
        """
        super().__init__()
        self.sigma = sigma

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
        if self._mode is None:
            raise NotImplementedError
        return _ops.fused_loss(scores, relevance, n, _lib.FAMILY_LAMBDA, self._mode,
                               float(self.sigma), loss_sum)


class LambdaARPLoss1(LambdaLoss):
    r"""ARP loss 1: :math:`-\sum_{i,j} \log_2 \mathrm{sigmoid}(\sigma(s_i - s_j))^{y_i}`
    (reference :95-117)."""
    _mode = _lib.LAM_ARP1


class LambdaARPLoss2(LambdaLoss):
    r"""ARP loss 2: :math:`\sum_{y_i > y_j} |y_i - y_j| \log_2(1 + e^{-\sigma(s_i - s_j)})`
    (reference :120-140)."""
    _mode = _lib.LAM_ARP2


class LambdaNDCGLoss1(LambdaLoss):
    r"""NDCG loss 1: exponent :math:`G_{\pi_i} / D_i` per ordered pair (reference :143-173)."""
    _mode = _lib.LAM_NDCG1


class LambdaNDCGLoss2(LambdaLoss):
    r"""NDCG loss 2: exponent :math:`\delta_{ij} |G_{\pi_i} - G_{\pi_j}|` over pairs with
    :math:`y_i > y_j` (reference :176-218)."""
    _mode = _lib.LAM_NDCG2
