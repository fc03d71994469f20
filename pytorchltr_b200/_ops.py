"""Torch <-> C-ABI glue: input normalisation, launches on the current stream, autograd.

PyTorch is plumbing here (device memory, streams, the autograd graph); every number
is produced by ``libltr_sm100.so``.  CPU tensors are accepted the way the reference
accepts them, but are staged through the GPU (H2D copy, kernel, D2H copy): there is
no CPU arithmetic path, and without a CUDA device the call raises.
"""
from typing import Optional, Tuple

import torch
from torch.autograd.function import once_differentiable

from pytorchltr_b200 import _lib


def _device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pytorchltr_b200 computes on CUDA only (sm_100a kernels, no CPU fallback) and "
            "no CUDA device is available")
    return torch.device("cuda", torch.cuda.current_device())


_REL_DTYPES = (torch.int64, torch.int32, torch.int16, torch.uint8)


def _integer_labels(relevance: torch.Tensor) -> torch.Tensor:
    """Relevance in a dtype the kernels read directly (int64 / int32 / int16 / uint8).  Other integer
    dtypes are widened to int64; floating-point labels must be integer valued (the kernels work on
    integer grades: a fractional label would be truncated silently otherwise)."""
    if relevance.dtype in _REL_DTYPES:
        return relevance
    if relevance.dtype.is_floating_point:
        if not bool((relevance == relevance.round()).all()):
            raise ValueError("relevance labels must be integers (got fractional floating-point labels)")
    return relevance.to(torch.int64)


def normalise(scores: torch.Tensor, relevance: Optional[torch.Tensor], n: torch.Tensor,
              device: torch.device):
    """Returns contiguous device tensors ``scores (B, L) f32``, ``relevance (B, L) i64|i32|i16|u8``
    and ``n (B,) i64|i32``.  Accepts ``(B, L)`` or ``(B, L, 1)`` like the reference
    (loss/pairwise_additive.py:60-65)."""
    if scores.dim() == 3:
        scores = scores.reshape(scores.shape[0], scores.shape[1])
    if scores.dim() != 2:
        raise ValueError(f"scores must be (B, L) or (B, L, 1), got {tuple(scores.shape)}")
    B, L = scores.shape
    if L < 1:
        raise ValueError("list size must be at least 1")
    if L > _lib.MAX_LIST_SIZE:
        raise ValueError(f"list size {L} exceeds LTR_MAX_LIST_SIZE={_lib.MAX_LIST_SIZE}")
    s = scores.detach()
    if s.dtype != torch.float32:
        s = s.to(torch.float32)
    s = s.to(device, non_blocking=True).contiguous()
    y = None
    if relevance is not None:
        if relevance.dim() == 3:
            relevance = relevance.reshape(relevance.shape[0], relevance.shape[1])
        if tuple(relevance.shape) != (B, L):
            raise ValueError(
                f"relevance {tuple(relevance.shape)} does not match scores {(B, L)}")
        y = _integer_labels(relevance.detach()).to(device, non_blocking=True).contiguous()
    if n.dim() != 1 or n.shape[0] != B:
        raise ValueError(f"n must have shape ({B},), got {tuple(n.shape)}")
    nn = n.detach()
    if nn.dtype not in (torch.int64, torch.int32):
        nn = nn.to(torch.int64)
    nn = nn.to(device, non_blocking=True).contiguous()
    return s, y, nn, B, L


def host_loss(scores: torch.Tensor, relevance: torch.Tensor, n: torch.Tensor, family: int, mode: int,
              sigma: float, want_grad: bool, dev: torch.device):
    """CPU caller: ONE call of the host-buffer entry point (ltr_loss_host_ex: H2D of scores /
    relevance / n -> ordering + fused kernel -> D2H of the loss, chunked and overlapped for large
    batches) and one stream synchronisation.  Returns the loss as a pinned host tensor and a device
    view of d loss / d scores inside the call's workspace (the backward pass scales it by the upstream
    gradient on the device and copies it to the host once)."""
    if scores.dim() == 3:
        scores = scores.reshape(scores.shape[0], scores.shape[1])
    if scores.dim() != 2:
        raise ValueError(f"scores must be (B, L) or (B, L, 1), got {tuple(scores.shape)}")
    B, L = scores.shape
    if L < 1:
        raise ValueError("list size must be at least 1")
    if L > _lib.MAX_LIST_SIZE:
        raise ValueError(f"list size {L} exceeds LTR_MAX_LIST_SIZE={_lib.MAX_LIST_SIZE}")
    if relevance.dim() == 3:
        relevance = relevance.reshape(relevance.shape[0], relevance.shape[1])
    if tuple(relevance.shape) != (B, L):
        raise ValueError(f"relevance {tuple(relevance.shape)} does not match scores {(B, L)}")
    if n.dim() != 1 or n.shape[0] != B:
        raise ValueError(f"n must have shape ({B},), got {tuple(n.shape)}")
    s = scores.detach().to(torch.float32).contiguous()
    y = _integer_labels(relevance.detach()).contiguous()
    nn = n.detach()
    if nn.dtype not in (torch.int64, torch.int32):
        nn = nn.to(torch.int64)
    nn = nn.contiguous()
    loss_h = torch.empty(B, dtype=torch.float32, device="cpu", pin_memory=True)
    grad_d = None
    if B > 0:
        lib = _lib.lib()
        ws_bytes = lib.ltr_host_workspace_bytes(B, L)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _lib.on_device(dev):
            rc = lib.ltr_loss_host_ex(family, mode, s.data_ptr(), y.data_ptr(), y.element_size(), nn.data_ptr(),
                                      nn.element_size(), B, L, float(sigma), loss_h.data_ptr(), None,
                                      1 if want_grad else 0, ws.data_ptr(), ws_bytes, _stream(dev))
            _lib.check(rc)
            torch.cuda.current_stream(dev).synchronize()
        if want_grad:
            off = lib.ltr_host_workspace_dscores_offset(B, L)
            grad_d = ws[off:off + 4 * B * L].view(torch.float32).view(B, L)
    elif want_grad:
        grad_d = torch.empty((0, L), dtype=torch.float32, device=dev)
    return loss_h, grad_d


def host_scaled_grad(g_scalar: float, grad_d: torch.Tensor) -> torch.Tensor:
    """``g_scalar * grad_d`` as a pinned HOST tensor: scaled on the device, copied once
    (ltr_scale_rows_host, chunked so that the copy overlaps the scaling), one synchronisation."""
    B, L = grad_d.shape
    out = torch.empty((B, L), dtype=torch.float32, device="cpu", pin_memory=True)
    if B == 0:
        return out
    dev = grad_d.device
    scratch = torch.empty_like(grad_d) if g_scalar != 1.0 else None
    with _lib.on_device(dev):
        rc = _lib.lib().ltr_scale_rows_host(float(g_scalar), grad_d.data_ptr(), out.data_ptr(), B, L,
                                            _ptr(scratch), _stream(dev))
        _lib.check(rc)
        torch.cuda.current_stream(dev).synchronize()
    return out


def to_host(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Device -> host copy for callers that passed CPU tensors: staged through pinned memory
    (torch's caching host allocator) so the copy runs at full PCIe rate, then one stream sync."""
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host


def _stream(device: torch.device) -> int:
    return _lib.raw_stream(device)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def launch_loss(family: int, mode: int, s: torch.Tensor, y: torch.Tensor, nn: torch.Tensor,
                sigma: float, want_grad: bool, want_ranking: bool = False,
                loss_sum: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """One fused launch: per-query loss (B,), d loss / d scores (B, L), ranking (B, L)."""
    B, L = s.shape
    dev = s.device
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    grad = torch.empty((B, L), dtype=torch.float32, device=dev) if want_grad else None
    ranking = torch.empty((B, L), dtype=torch.int64, device=dev) if want_ranking else None
    lib = _lib.lib()
    with _lib.on_device(dev):
        st = _stream(dev)
        if family != _lib.FAMILY_LISTNET:
            # scheduling workspace of the O(L^2) losses (longest queries first): scratch from the
            # caching allocator, stream-ordered like the outputs
            ws_bytes = lib.ltr_schedule_workspace_bytes(B)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if family == _lib.FAMILY_ADDITIVE:
            rc = lib.ltr_pairwise_additive_ws(mode, s.data_ptr(), y.data_ptr(), y.element_size(),
                                              nn.data_ptr(), nn.element_size(), B, L, float(sigma),
                                              loss.data_ptr(), _ptr(grad), _ptr(loss_sum),
                                              ws.data_ptr(), ws_bytes, st)
        elif family == _lib.FAMILY_LAMBDA:
            rc = lib.ltr_lambda_ws(mode, s.data_ptr(), y.data_ptr(), y.element_size(),
                                   nn.data_ptr(), nn.element_size(), B, L, float(sigma),
                                   loss.data_ptr(), _ptr(grad), _ptr(ranking), _ptr(loss_sum),
                                   ws.data_ptr(), ws_bytes, st)
        elif family == _lib.FAMILY_LISTNET:
            rc = lib.ltr_listnet(s.data_ptr(), y.data_ptr(), y.element_size(), nn.data_ptr(),
                                 nn.element_size(), B, L, loss.data_ptr(), _ptr(grad),
                                 _ptr(loss_sum), st)
        else:
            raise ValueError(f"unknown loss family {family}")
    _lib.check(rc)
    return loss, grad, ranking


def scale_rows(g: torch.Tensor, dscores: torch.Tensor) -> torch.Tensor:
    """``g[:, None] * dscores`` on the device (ltr_scale_rows)."""
    B, L = dscores.shape
    g = g.detach()
    # `loss.sum().backward()` hands over a broadcast (stride-0) gradient: read it in place
    # instead of materialising B copies of the same number
    if B > 1 and g.stride(0) == 0:
        g_stride = 0
        if not g.is_cuda:
            # CPU caller: the scalar is already on the host -- ship it as a fill, not as a copy
            g = torch.full((1,), float(g[0]), dtype=torch.float32, device=dscores.device)
        else:
            g = g.to(device=dscores.device, dtype=torch.float32)
    else:
        g = g.to(device=dscores.device, dtype=torch.float32).contiguous()
        g_stride = 1
    out = torch.empty_like(dscores)
    if B == 0:
        return out
    with _lib.on_device(dscores.device):
        rc = _lib.lib().ltr_scale_rows(g.data_ptr(), g_stride, dscores.data_ptr(), out.data_ptr(), B, L,
                                       _stream(dscores.device))
    _lib.check(rc)
    return out


class _FusedLoss(torch.autograd.Function):
    """``forward`` computes the per-query loss and the unscaled d loss_b / d scores in ONE
    kernel and saves the latter; ``backward`` is the row scale ``g[b] * saved[b, :]``.

    Nothing flows to relevance / n; the ranking, gains and discounts are constants of the
    backward pass, exactly as in the reference (the gather indices carry no gradient).
    Double backward is not supported (``once_differentiable``); the reference supports it
    only implicitly through composite autograd.
    """

    @staticmethod
    def forward(ctx, scores, relevance, n, family, mode, sigma, loss_sum=None):
        dev = _device_for(scores)
        want_grad = bool(ctx.needs_input_grad[0])
        ctx.scores_shape = scores.shape
        ctx.scores_dtype = scores.dtype
        ctx.scores_device = scores.device
        ctx.host_caller = False
        if not scores.is_cuda and not relevance.is_cuda and not n.is_cuda:
            # CPU caller: host-buffer entry point, results come back as (pinned) CPU tensors
            if loss_sum is not None:
                raise ValueError("loss_sum needs CUDA inputs")
            loss_h, grad_d = host_loss(scores, relevance, n, family, mode, sigma, want_grad, dev)
            if want_grad:
                ctx.host_caller = True
                ctx.save_for_backward(grad_d)
            if loss_h.dtype != scores.dtype and scores.dtype.is_floating_point:
                loss_h = loss_h.to(scores.dtype)
            return loss_h
        s, y, nn, B, L = normalise(scores, relevance, n, dev)
        if B == 0:
            loss = torch.empty(0, dtype=torch.float32, device=dev)
            grad = torch.empty((0, L), dtype=torch.float32, device=dev) if want_grad else None
        else:
            loss, grad, _ = launch_loss(family, mode, s, y, nn, sigma, want_grad, loss_sum=loss_sum)
        if want_grad:
            ctx.save_for_backward(grad)
        out = loss
        if out.dtype != scores.dtype and scores.dtype.is_floating_point:
            out = out.to(scores.dtype)
        if out.device != scores.device:
            out = to_host(out, scores)    # synchronising D2H copy: CPU callers get CPU results
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (saved,) = ctx.saved_tensors
        if ctx.host_caller and not g.is_cuda:
            B = saved.shape[0]
            # `loss.sum().backward()` / `loss.mean().backward()` of a CPU caller: the upstream gradient is
            # one broadcast scalar -- it travels as a kernel argument, the product comes back in one copy
            if B <= 1 or g.stride(0) == 0:
                g0 = float(g.reshape(-1)[0]) if B > 0 else 1.0
                out = host_scaled_grad(g0, saved)
                if out.dtype != ctx.scores_dtype:
                    out = out.to(ctx.scores_dtype)
                return out.reshape(ctx.scores_shape), None, None, None, None, None, None
        out = scale_rows(g, saved)
        if out.dtype != ctx.scores_dtype:
            out = out.to(ctx.scores_dtype)
        if out.device != ctx.scores_device:
            out = to_host(out, g)
        return out.reshape(ctx.scores_shape), None, None, None, None, None, None


def fused_loss(scores, relevance, n, family: int, mode: int, sigma: float = 1.0,
               loss_sum: Optional[torch.Tensor] = None):
    """``loss_sum`` (extension, CUDA callers only): a float32 device scalar to which the kernel
    adds ``sum_b loss_b`` in its epilogue (atomically; the caller zeroes it) -- the input of the
    scalar all-reduce behind a global mean (``pytorchltr_b200.distributed``)."""
    if loss_sum is not None:
        if not (scores.is_cuda and loss_sum.is_cuda and loss_sum.dtype == torch.float32 and loss_sum.numel() >= 1):
            raise ValueError("loss_sum must be a float32 CUDA tensor and needs CUDA inputs")
        loss_sum = loss_sum.detach()
    return _FusedLoss.apply(scores, relevance, n, family, mode, float(sigma), loss_sum)


def rank_metric(metric: int, scores, relevance, n, k: Optional[int], exp: bool):
    dev = _device_for(scores)
    s, y, nn, B, L = normalise(scores, relevance, n, dev)
    if k is not None:
        # dcg[:, :k][:, -1] (evaluation/dcg.py:97-98) with Python slice semantics
        kk = k if k >= 0 else L + k
        kk = min(kk, L)
        if kk <= 0:
            raise IndexError("index -1 is out of bounds for dimension 1 with size 0")
    else:
        kk = 0
    all_k = metric != _lib.METRIC_ARP and k is None
    out = torch.empty((B, L) if all_k else (B,), dtype=torch.float32, device=dev)
    if B > 0:
        with _lib.on_device(dev):
            rc = _lib.lib().ltr_rank_metrics(metric, s.data_ptr(), y.data_ptr(), y.element_size(),
                                             nn.data_ptr(), nn.element_size(), B, L,
                                             1 if metric == _lib.METRIC_ARP else kk,
                                             1 if exp else 0, out.data_ptr(), L if all_k else 1,
                                             _stream(dev))
        _lib.check(rc)
    if out.device != scores.device:
        out = to_host(out, scores)
    return out


def rank_by_score(scores, n):
    dev = _device_for(scores)
    s, _, nn, B, L = normalise(scores, None, n, dev)
    out = torch.empty((B, L), dtype=torch.int64, device=dev)
    if B > 0:
        with _lib.on_device(dev):
            rc = _lib.lib().ltr_rank_by_score(s.data_ptr(), nn.data_ptr(), nn.element_size(), B, L,
                                              out.data_ptr(), _stream(dev))
        _lib.check(rc)
    if out.device != scores.device:
        out = to_host(out, scores)
    return out
