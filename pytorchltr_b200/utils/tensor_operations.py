"""Tensor helpers with the names and signatures of ``pytorchltr.utils.tensor_operations``.

``rank_by_score`` runs on the sm_100a per-query argsort kernel.  The other helpers
are shape plumbing that the fused kernels made unnecessary on the hot path (no pair
tensor is ever materialised there); they are kept as thin torch expressions so code
written against the reference keeps working.
"""
from typing import Optional

import torch as _torch

from pytorchltr_b200 import _ops


def mask_padded_values(xs: _torch.FloatTensor, n: _torch.LongTensor,
                       mask_value: float = -float('inf'),
                       mutate: bool = False):
    """Sets ``xs[b, j]`` to ``mask_value`` for ``j >= n[b]`` (reference :6-26).

    Args:
        xs: ``(B, L)`` values.
        n: ``(B,)`` list sizes.
        mask_value: value written over the padding (default ``-inf``).
        mutate: write into ``xs`` instead of a copy.
    """
    out = xs if mutate else xs.clone()
    positions = _torch.arange(out.shape[1], device=out.device)
    padded = positions.unsqueeze(0) >= n.to(out.device).reshape(-1, 1)
    out[padded] = mask_value
    return out


def tiebreak_argsort(x: _torch.FloatTensor, descending: bool = True,
                     generator: Optional[_torch.Generator] = None) -> _torch.LongTensor:
    """Per-row argsort with ties broken by a random column permutation (reference :29-45)."""
    kwargs = {} if generator is None else {"generator": generator}
    perm = _torch.randperm(x.shape[1], device=x.device, **kwargs)
    order = _torch.argsort(x.index_select(1, perm), dim=1, descending=descending)
    return perm[order]


def rank_by_score(scores: _torch.FloatTensor, n: _torch.LongTensor,
                  generator: Optional[_torch.Generator] = None) -> _torch.LongTensor:
    """Ranks documents by descending score with padded documents last (reference :48-64).

    Runs ``ltr_rank_by_score``.  By default ties are broken lowest-index-first and padded documents
    keep index order (deterministic, no RNG state consumed).  Passing a ``generator`` opts into the
    reference's random tie-break (one random column permutation per call drawn from it, :43-45):
    the columns are permuted before the kernel and the indices mapped back, so that a constant or
    degenerate scorer is neither rewarded nor penalised by the document order of the file.
    """
    if generator is None:
        return _ops.rank_by_score(scores, n)
    if scores.dim() == 3:
        scores = scores.reshape(scores.shape[0], scores.shape[1])
    L = scores.shape[1]
    perm = _torch.randperm(L, device=scores.device, generator=generator)
    masked = mask_padded_values(scores.to(_torch.float32), n, mask_value=float("-inf"))
    full = _torch.full_like(n, L)
    return perm[_ops.rank_by_score(masked.index_select(1, perm), full)]


def rank_by_plackettluce(scores: _torch.FloatTensor, n: _torch.LongTensor,
                         generator: Optional[_torch.Generator] = None) -> _torch.LongTensor:
    """Samples a ranking from the Plackett-Luce distribution of the scores, padded documents
    last (reference :67-91).

    The reference sorts ``log(-log(u)) - log_softmax(scores)`` in ascending order; the
    log-softmax normaliser is constant within a query, so that order is the descending order of
    ``scores - log(-log(u))`` (Gumbel-max).  The uniform draws come from torch's generator, the
    ranking itself from the ``ltr_rank_by_score`` kernel.
    """
    if scores.dim() == 3:
        scores = scores.reshape(scores.shape[0], scores.shape[1])
    kwargs = {} if generator is None else {"generator": generator}
    u = _torch.rand(scores.shape, device=scores.device, dtype=_torch.float32, **kwargs)
    u = u.clamp_(min=1e-38)
    perturbed = scores.to(_torch.float32) - _torch.log(-_torch.log(u))
    return _ops.rank_by_score(perturbed, n)


def batch_pairs(x: _torch.Tensor) -> _torch.Tensor:
    """``p[b, i, j, 0] = x[b, i]``, ``p[b, i, j, 1] = x[b, j]`` (reference :94-119).

    Provided for API compatibility only: the losses never build this ``(B, L, L, 2)``
    tensor, they generate pairs on chip.
    """
    if x.dim() == 3:
        x = x.reshape(x.shape[0], x.shape[1])
    L = x.shape[1]
    first = x.unsqueeze(2).expand(-1, L, L)
    second = x.unsqueeze(1).expand(-1, L, L)
    return _torch.stack((first, second), dim=3)
