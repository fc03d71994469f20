"""PBM click simulators with the reference's signatures (click_simulation/pbm.py:12-129).

Click probabilities and observation propensities come from one kernel (``ltr_pbm_probabilities``,
written straight in document order: no gathers, no second argsort); the Bernoulli draw uses torch's
generator, like the reference."""
from typing import Optional, Tuple

import torch

from pytorchltr_b200 import _lib, _ops


def pbm_probabilities(rankings, ys, n, relevance_probs, cutoff: Optional[int] = None,
                      eta: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (click probability, observation propensity), both ``(B, L)`` float32 in document order."""
    dev = _ops._device_for(rankings)
    rk = rankings.detach().to(dev, torch.int64).contiguous()
    if rk.dim() != 2:
        raise ValueError(f"rankings must be (B, L), got {tuple(rk.shape)}")
    B, L = rk.shape
    y = ys.detach()
    if y.dim() == 3:
        y = y.reshape(B, L)
    if tuple(y.shape) != (B, L):
        raise ValueError(f"ys {tuple(y.shape)} does not match rankings {(B, L)}")
    y = _ops._integer_labels(y).to(dev).contiguous()
    nn_ = n.detach()
    if nn_.dtype not in (torch.int64, torch.int32):
        nn_ = nn_.to(torch.int64)
    nn_ = nn_.to(dev).contiguous()
    if nn_.dim() != 1 or nn_.shape[0] != B:
        raise ValueError(f"n must have shape ({B},), got {tuple(nn_.shape)}")
    rp = relevance_probs.detach().to(dev, torch.float32).reshape(-1).contiguous()
    if cutoff is not None and cutoff < 1:
        raise ValueError("cutoff must be at least 1")
    # zero-filled: a row of `rankings` that is not a full permutation (a top-k list, duplicates) leaves
    # the documents it does not name at probability 0 instead of uninitialised memory
    cp = torch.zeros((B, L), dtype=torch.float32, device=dev)
    pr = torch.zeros((B, L), dtype=torch.float32, device=dev)
    if B > 0:
        with _lib.on_device(dev):
            rc = _lib.lib().ltr_pbm_probabilities(rk.data_ptr(), y.data_ptr(), y.element_size(), nn_.data_ptr(),
                                                  nn_.element_size(), rp.data_ptr(), rp.numel(),
                                                  0 if cutoff is None else int(cutoff), float(eta), B, L,
                                                  cp.data_ptr(), pr.data_ptr(),
                                                  _lib.raw_stream(dev))
        _lib.check(rc)
    return cp, pr


def simulate_pbm(rankings, ys, n, relevance_probs, cutoff: Optional[int] = None, eta: float = 1.0):
    """Clicks ``(B, L)`` int64 in {0, 1} and propensities ``(B, L)`` float32 (reference :12-63)."""
    cp, pr = pbm_probabilities(rankings, ys, n, relevance_probs, cutoff, eta)
    clicks = torch.bernoulli(cp).to(torch.int64)
    if not rankings.is_cuda:
        clicks, pr = clicks.cpu(), pr.cpu()
    return clicks, pr


def simulate_perfect(rankings, ys, n, cutoff: Optional[int] = None):
    """Perfect user model (reference :66-83)."""
    return simulate_pbm(rankings, ys, n, torch.tensor([0.0, 0.2, 0.4, 0.8, 1.0]), cutoff, 0.0)


def simulate_position(rankings, ys, n, cutoff: Optional[int] = None, eta: float = 1.0):
    """Binary position-biased user model (reference :86-105)."""
    return simulate_pbm(rankings, ys, n, torch.tensor([0.1, 0.1, 0.1, 1.0, 1.0]), cutoff, eta)


def simulate_nearrandom(rankings, ys, n, cutoff: Optional[int] = None, eta: float = 1.0):
    """Near-random user model (reference :108-127)."""
    return simulate_pbm(rankings, ys, n, torch.tensor([0.4, 0.45, 0.5, 0.55, 0.6]), cutoff, eta)
