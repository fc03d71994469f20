"""Fused linear scorer + ListNet: the caller side of the loss path.

The reference's training step (examples/01-basic-usage.py:44,72, getting-started.rst:42-51) is

    model = torch.nn.Linear(F, 1)
    loss = loss_fn(model(xs), ys, n).mean(); loss.backward()

which reads the ``(B, L, F)`` feature tensor twice (forward and backward GEMV).
``LinearListNet`` computes scores, the ListNet loss, d loss / d scores and the per-query weight /
bias gradients in ONE pass over the features (``ltr_linear_listnet``); the backward pass for any
upstream gradient is a small weighted column sum (``ltr_linear_listnet_backward``).  The parameters live in a
``torch.nn.Linear(F, 1)`` so that a ``state_dict`` is interchangeable with the reference's model.
"""
import torch
from torch.autograd.function import once_differentiable

from pytorchltr_b200 import _lib, _ops


class _LinearListNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, weight, bias, relevance, n):
        if ctx.needs_input_grad[0]:
            # the fused kernel never forms d loss / d features (B x L x F): returning None would silently
            # starve an upstream encoder of its gradient
            raise NotImplementedError(
                "linear_listnet is differentiable with respect to weight and bias only; `features` requires "
                "grad (it comes from a trainable module): use torch.nn.Linear + ListNetLoss for that model")
        if not features.is_cuda:
            raise RuntimeError("linear_listnet computes on CUDA only (sm_100a kernels, no CPU fallback)")
        if features.dim() != 3:
            raise ValueError(f"features must be (B, L, F), got {tuple(features.shape)}")
        B, L, F = features.shape
        dev = features.device
        x = features.detach()
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        x = x.contiguous()
        w = weight.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if w.numel() != F:
            raise ValueError(f"weight must have {F} elements, got {tuple(weight.shape)}")
        b = None if bias is None else bias.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if relevance.dim() == 3:
            relevance = relevance.reshape(B, L)
        if tuple(relevance.shape) != (B, L):
            raise ValueError(f"relevance {tuple(relevance.shape)} does not match features {(B, L, F)}")
        y = _ops._integer_labels(relevance.detach()).to(dev).contiguous()
        if n.dim() != 1 or n.shape[0] != B:
            raise ValueError(f"n must have shape ({B},), got {tuple(n.shape)}")
        nn_ = n.detach()
        if nn_.dtype not in (torch.int64, torch.int32):
            nn_ = nn_.to(torch.int64)
        nn_ = nn_.to(dev).contiguous()

        lib = _lib.lib()
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        dscores = torch.empty((B, L), dtype=torch.float32, device=dev)
        qgrad = torch.empty((B, F + 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.ltr_linear_listnet(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(),
                                        y.data_ptr(), y.element_size(), nn_.data_ptr(), nn_.element_size(),
                                        B, L, F, None, loss.data_ptr(), dscores.data_ptr(), qgrad.data_ptr(),
                                        None, st)
            fused = rc == 0
            if rc == -2:
                # LTR_EUNSUPPORTED: not even the tiled kernel's O(L + F) shared memory fits (F beyond ~50k):
                # scorer by the library GEMV, then the ListNet kernel
                s = torch.mv(x.reshape(B * L, F), w)
                if b is not None:
                    s = s + b
                loss, dscores, _ = _ops.launch_loss(_lib.FAMILY_LISTNET, 0, s.reshape(B, L), y, nn_, 1.0, True)
            else:
                _lib.check(rc)
        ctx.fused = fused
        ctx.has_bias = bias is not None
        ctx.weight_shape = weight.shape
        ctx.bias_shape = None if bias is None else bias.shape
        if fused:
            ctx.save_for_backward(qgrad)
        else:
            ctx.save_for_backward(x, dscores)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.detach()
        if ctx.fused:
            (qgrad,) = ctx.saved_tensors
            B, F = qgrad.shape[0], qgrad.shape[1] - 1
            dev = qgrad.device
            g = g.to(device=dev, dtype=torch.float32)
            g_stride = 1
            if B > 1 and g.stride(0) == 0:
                g_stride = 0                     # `.sum()`: a broadcast scalar, read in place
            else:
                g = g.contiguous()
            lib = _lib.lib()
            gw = torch.empty(F, dtype=torch.float32, device=dev)
            gb = torch.empty(1, dtype=torch.float32, device=dev)
            ws_bytes = lib.ltr_linear_listnet_workspace_bytes(F)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                rc = lib.ltr_linear_listnet_backward(qgrad.data_ptr(), g.data_ptr(), g_stride, B, F, gw.data_ptr(),
                                                     gb.data_ptr(), ws.data_ptr(), ws_bytes,
                                                     torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc)
        else:
            x, dscores = ctx.saved_tensors
            B, L, F = x.shape
            gd = _ops.scale_rows(g.to(device=x.device, dtype=torch.float32), dscores).reshape(B * L)
            gw = torch.mv(x.reshape(B * L, F).t(), gd)
            gb = gd.sum().reshape(1)
        gw = gw.reshape(ctx.weight_shape)
        gb = gb.reshape(ctx.bias_shape) if ctx.has_bias else None
        return None, gw, gb, None, None


def linear_listnet(features, weight, bias, relevance, n):
    """Per-query ListNet loss of the linear ranker ``features @ weight + bias``; differentiable with
    respect to ``weight`` and ``bias`` (not ``features``)."""
    return _LinearListNet.apply(features, weight, bias, relevance, n)


class LinearListNet(torch.nn.Module):
    """``ListNetLoss()(torch.nn.Linear(F, 1)(xs), ys, n)`` in one pass over ``xs``.

    ``forward(xs, relevance, n) -> FloatTensor (B,)``; ``self.linear`` is a ``torch.nn.Linear(F, 1)``
    (same parameter names and shapes as the reference's model), ``score(xs)`` returns its output.
    """

    def __init__(self, in_features: int, bias: bool = True):
        super().__init__()
        self.linear = torch.nn.Linear(in_features, 1, bias=bias)

    def score(self, xs):
        return self.linear(xs)

    def forward(self, xs, relevance, n):
        return linear_listnet(xs, self.linear.weight, self.linear.bias, relevance, n)
