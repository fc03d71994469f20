"""Fused scorers: the caller side of the loss path -- linear scorer + ListNet in one
pass, and the documented MLP ranker on the tensor cores (``MLPRanker``, end of the file).

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
        with _lib.on_device(dev):
            st = _lib.raw_stream(dev)
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
            with _lib.on_device(dev):
                rc = lib.ltr_linear_listnet_backward(qgrad.data_ptr(), g.data_ptr(), g_stride, B, F, gw.data_ptr(),
                                                     gb.data_ptr(), ws.data_ptr(), ws_bytes,
                                                     _lib.raw_stream(dev))
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


# ---- MLP scorer: the reference's documented model on the tensor cores --------------------
_LTR_EUNSUPPORTED = -2
_warned_fallback = set()


def _warn_fallback(what):
    if what not in _warned_fallback:
        _warned_fallback.add(what)
        import warnings
        warnings.warn(f"pytorchltr_b200.fused.MLPRanker: {what}; this shape runs on plain torch modules "
                      "(see include/ltr_sm100.h, ltr_mlp_scores / ltr_mlp_backward, for the kernel's limits)")


def _mlp_torch_forward(x2, w1, b1, w2, b2, w3, b3):
    h1 = torch.relu(torch.nn.functional.linear(x2, w1, b1))
    h2 = torch.relu(torch.nn.functional.linear(h1, w2, b2))
    return torch.nn.functional.linear(h2, w3.reshape(1, -1), b3).reshape(-1), h1, h2


class _MlpScores(torch.autograd.Function):
    """scores = l3(relu(l2(relu(l1(x))))) over the flat (rows, F) feature block: ``ltr_mlp_scores`` forward
    (features read once, layer 1 on tcgen05 with TF32 operands), ``ltr_mlp_backward`` for the parameter
    gradients (second pass over the features, dW1 on tcgen05).  Differentiable with respect to the six
    parameters, not the features.  With an ``exchange`` (``distributed.PeerExchange``) the gradients come back
    summed over the ranks of the node (``ltr_mlp_backward_allreduce``: the exchange runs inside the final
    reduction of the backward pass, over NVLink peer memory)."""

    @staticmethod
    def forward(ctx, features, w1, b1, w2, b2, w3, b3, exchange=None):
        ctx.exchange = exchange
        if ctx.needs_input_grad[0]:
            raise NotImplementedError(
                "MLPRanker is differentiable with respect to its parameters only; `features` requires grad "
                "(it comes from a trainable module): use plain torch.nn.Linear layers for that model")
        if not features.is_cuda:
            raise RuntimeError("MLPRanker computes on CUDA only (sm_100a kernels, no CPU fallback)")
        F = features.shape[-1]
        dev = features.device
        x2 = features.detach()
        if x2.dtype != torch.float32:
            x2 = x2.to(torch.float32)
        x2 = x2.reshape(-1, F).contiguous()
        rows = x2.shape[0]
        # TMA needs feature rows of a multiple of 16 bytes: other widths (MQ2007's 46, Yahoo's 699) get zero
        # columns appended -- one extra copy of the features per call; a dataset kept on the device can be
        # stored padded once instead (zero features times zero weight columns change nothing).  When a copy is
        # made anyway it goes to a multiple of 8 floats: rows that start on 32-byte sectors stream ~20 % faster
        # (700 -> 704 features: 5.3 -> 6.4 TB/s, tools/mlp_width_sweep.py)
        Fp = F if F % 4 == 0 else (F + 7) // 8 * 8
        if Fp != F:
            x2 = torch.nn.functional.pad(x2, (0, Fp - F))

        def prep(t, shape):
            if t is None:
                return None
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"parameter of shape {tuple(t.shape)}, expected {shape}")
            return t
        H1, H2 = w1.shape[0], w2.shape[0]
        p = [prep(w1, (H1, F)), prep(b1, (H1,)), prep(w2, (H2, H1)), prep(b2, (H2,)), prep(w3, (1, H2)),
             prep(b3, (1,))]
        if Fp != F:
            p[0] = torch.nn.functional.pad(p[0], (0, Fp - F))
        ctx.F_in = F
        F = Fp
        ptr = [None if t is None else t.data_ptr() for t in p]
        lib = _lib.lib()
        scores = torch.empty(rows, dtype=torch.float32, device=dev)
        # a training step keeps [H1 | Z2] per document for the backward pass (as autograd would keep the
        # activations of the torch layers); an inference call writes the scores only
        hz = None
        pitch = lib.ltr_mlp_hz_pitch(H1, H2)
        if pitch and any(ctx.needs_input_grad[1:]):
            hz = torch.empty((rows, pitch), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            rc = lib.ltr_mlp_scores(x2.data_ptr(), rows, F, ptr[0], ptr[1], H1, ptr[2], ptr[3], H2, ptr[4], ptr[5],
                                    scores.data_ptr(), None if hz is None else hz.data_ptr(),
                                    _lib.raw_stream(dev))
        if rc == _LTR_EUNSUPPORTED:
            _warn_fallback(f"features={F}, hidden=({H1}, {H2}) is outside the scorer kernel's limits")
            scores = _mlp_torch_forward(x2, *p)[0]
        else:
            _lib.check(rc)
        if rc == _LTR_EUNSUPPORTED:
            hz = None
        ctx.save_for_backward(x2, *[t for t in p if t is not None], *([] if hz is None else [hz]))
        ctx.has_hz = hz is not None
        ctx.has = [t is not None for t in p]
        ctx.dims = (rows, F, H1, H2)
        ctx.param_shapes = [None if t is None else t.shape for t in (w1, b1, w2, b2, w3, b3)]
        return scores.reshape(features.shape[:-1] + (1,))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        x2 = saved.pop(0)
        p = [saved.pop(0) if h else None for h in ctx.has]
        hz = saved.pop(0) if ctx.has_hz else None
        rows, F, H1, H2 = ctx.dims
        dev = x2.device
        ds = g.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        ptr = [None if t is None else t.data_ptr() for t in p]
        lib = _lib.lib()
        n = lib.ltr_mlp_grad_len(F, H1, H2)
        grads = torch.empty(n, dtype=torch.float32, device=dev)
        ws_bytes = lib.ltr_mlp_workspace_bytes(F, H1, H2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ex = ctx.exchange

        def call(hz_ptr):
            st = _lib.raw_stream(dev)
            head = (x2.data_ptr(), rows, F, ptr[0], ptr[1], H1, ptr[2], ptr[3], H2, ptr[4], ptr[5], hz_ptr,
                    ds.data_ptr(), grads.data_ptr(), ws.data_ptr(), ws_bytes)
            if ex is None:
                return lib.ltr_mlp_backward(*head, st)
            return lib.ltr_mlp_backward_allreduce(*head, ex.handle, st)      # an unsupported shape exchanges nothing
        with _lib.on_device(dev):
            rc = call(None if hz is None else hz.data_ptr())
            if rc == _LTR_EUNSUPPORTED and hz is not None:      # no kept-activation kernel for this shape: recompute
                rc = call(None)
        if rc == _LTR_EUNSUPPORTED:
            _warn_fallback(f"features={F}, hidden=({H1}, {H2}) is outside the backward kernel's limits")
            _, h1, h2 = _mlp_torch_forward(x2, *p)
            d = ds.reshape(-1, 1)
            dz2 = d * p[4].reshape(1, -1) * (h2 > 0)
            dz1 = (dz2 @ p[2]) * (h1 > 0)
            out = [dz1.t() @ x2, dz1.sum(0), dz2.t() @ h1, dz2.sum(0), (d * h2).sum(0).reshape(1, -1), d.sum().reshape(1)]
            if ex is not None:
                flat = torch.cat([t.reshape(-1) for t in out]).contiguous()
                ex.all_reduce_vec_(flat)
                out = [c.reshape(t.shape) for c, t in zip(flat.split([t.numel() for t in out]), out)]
        else:
            _lib.check(rc)
            o = [0, H1 * F, H1 * F + H1, H1 * F + H1 + H2 * H1, H1 * F + H1 + H2 * H1 + H2,
                 H1 * F + H1 + H2 * H1 + 2 * H2, n]
            out = [grads[o[k]:o[k + 1]] for k in range(6)]
        if ctx.F_in != F:                              # drop the gradient of the padding columns of W1
            out[0] = out[0].reshape(H1, F)[:, :ctx.F_in]
        res = [None]
        for k in range(6):
            res.append(out[k].reshape(ctx.param_shapes[k]) if ctx.has[k] else None)
        res.append(None)
        return tuple(res)


def mlp_scores(features, w1, b1, w2, b2, w3, b3, exchange=None):
    """``(..., F) -> (..., 1)`` scores of the ReLU MLP with torch.nn.Linear-shaped parameters
    ``w1 (H1, F), b1 (H1,), w2 (H2, H1), b2 (H2,), w3 (1, H2), b3 (1,)`` (biases may be None).
    ``exchange``: a ``distributed.PeerExchange``; the parameter gradients are then summed over its ranks."""
    return _MlpScores.apply(features, w1, b1, w2, b2, w3, b3, exchange)


class MLPRanker(torch.nn.Module):
    """The reference's documented scoring function (docs/source/getting-started.rst:42-51)::

        class Model(torch.nn.Module):   # l1 = Linear(F, 50), l2 = Linear(50, 10), l3 = Linear(10, 1), ReLU between

    with the same attribute names (``l1``, ``l2``, ``l3``: a ``state_dict`` moves either way), scored by the
    tcgen05 kernels: ``forward(xs: (B, L, F)) -> (B, L, 1)``, to be fed to any loss or metric of this package
    exactly as ``loss_fn(model(xs), ys, n)`` in the reference's training loop (``:83-138``).
    Layer 1 multiplies TF32 operands (what torch does under ``torch.backends.cuda.matmul.allow_tf32``).

    Data-parallel training over the GPUs of one node (one process per GPU, each on its shard of the queries): pass
    ``exchange=distributed.PeerExchange()``; every rank's ``backward`` then returns the parameter gradients SUMMED
    over the ranks (the all-reduce DistributedDataParallel would issue, fused into the backward pass's final
    reduction over NVLink peer memory).  With ``distributed.sharded_mean_loss`` -- the loss divided by the global
    query count -- that sum is the gradient of the global mean, identical bits on every rank, so the replicas'
    optimizers stay in step without a wrapper.
    """

    def __init__(self, in_features: int, hidden=(50, 10), exchange=None):
        super().__init__()
        h1, h2 = hidden
        self.exchange = exchange
        self.l1 = torch.nn.Linear(in_features, h1)
        self.l2 = torch.nn.Linear(h1, h2)
        self.l3 = torch.nn.Linear(h2, 1)

    def forward(self, x):
        return mlp_scores(x, self.l1.weight, self.l1.bias, self.l2.weight, self.l2.bias, self.l3.weight, self.l3.bias,
                          self.exchange)
