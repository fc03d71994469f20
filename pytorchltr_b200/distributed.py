"""Query-sharded data parallelism for the loss / metric path.

Every query (row ``b`` of the padded batch) is independent -- the reference reduces within
a row only (loss/pairwise_additive.py:45, loss/pairwise_lambda.py:48) -- so a batch is split
into contiguous blocks of queries, one per rank, with NO data-path collective: each rank
runs the fused kernel on its own shard and the gradient w.r.t. its scores stays local.
The only exchange is the 2-element ``[sum, count]`` all-reduce behind a global mean
(NCCL over NVLink / NVSwitch for CUDA tensors, gloo for the CPU tests).

The reference has no distributed code at all; this module is new surface.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from pytorchltr_b200 import _lib


def shard_bounds(num_queries: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` block of queries owned by ``rank``; blocks differ by at most
    one query and cover ``range(num_queries)`` exactly."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, extra = divmod(num_queries, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(scores: torch.Tensor, relevance: torch.Tensor, n: torch.Tensor,
                rank: Optional[int] = None, world_size: Optional[int] = None):
    """This rank's block of queries of a replicated ``(scores, relevance, n)`` batch."""
    if rank is None:
        rank = dist.get_rank()
    if world_size is None:
        world_size = dist.get_world_size()
    lo, hi = shard_bounds(scores.shape[0], rank, world_size)
    return scores[lo:hi], relevance[lo:hi], n[lo:hi]


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def global_sum_count(per_query: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduced ``[sum over all queries of all ranks, number of queries]`` (float64 on
    CPU / float32 on CUDA); one 2-element collective."""
    dtype = torch.float32 if per_query.is_cuda else torch.float64
    # built from device-side fills only (no host->device copy): CUDA-graph capturable
    buf = torch.empty(2, dtype=dtype, device=per_query.device)
    buf[:1].copy_(per_query.detach().sum().to(dtype).reshape(1))
    buf[1:].fill_(float(per_query.numel()))
    if _world(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def global_mean(per_query: torch.Tensor, group=None) -> torch.Tensor:
    """Mean of a per-query quantity (a loss or a metric) over the queries of ALL ranks."""
    buf = global_sum_count(per_query, group)
    return (buf[0] / buf[1].clamp(min=1.0)).to(per_query.dtype)


class PeerScalarExchange:
    """Scalar all-reduce over NVLink peer memory for the ranks of ONE node (``ltr_p2p_*`` in
    ``include/ltr_sm100.h``): every rank owns a mailbox in device memory, exported by CUDA IPC; a step's
    exchange is one single-CTA kernel (remote 64-bit stores + a spin on the own mailbox) instead of a
    library collective -- a few microseconds, CUDA-graph capturable.  Construct it once, collectively
    (the IPC handles travel through ``torch.distributed.all_gather_object``), after the process group
    exists and the CUDA device of the rank is current; pass it to :func:`sharded_mean_loss`.
    """

    def __init__(self, group=None):
        import ctypes

        from pytorchltr_b200 import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerScalarExchange needs an initialised process group")
        self._lib = _lib.lib()
        self._check = _lib.check
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._handle = ctypes.c_void_p()
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(self._lib.ltr_p2p_create(self.rank, self.world, ctypes.byref(self._handle), mine))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(mine), group=group)
        blob = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(gathered))
        rc = self._lib.ltr_p2p_connect(self._handle, blob)
        # every rank learns whether every rank connected (a rank that failed must not leave the others
        # spinning on a mailbox nobody writes)
        ok = [None] * self.world
        dist.all_gather_object(ok, rc == 0, group=group)
        if not all(ok):
            self.close()
            raise RuntimeError("CUDA IPC peer mapping failed on rank(s) %s" % [i for i, v in enumerate(ok) if not v])

    def all_reduce_(self, values: torch.Tensor) -> torch.Tensor:
        """In place: ``values`` (float32 CUDA, 1..4 elements) becomes the sum over all ranks."""
        if not (values.is_cuda and values.dtype == torch.float32 and values.is_contiguous() and 1 <= values.numel() <= 4):
            raise ValueError("values must be a contiguous float32 CUDA tensor with 1 to 4 elements")
        with _lib.on_device(values.device):
            self._check(self._lib.ltr_p2p_allreduce_sum(self._handle, values.data_ptr(), values.numel(),
                                                        _lib.raw_stream(values.device)))
        return values

    def all_reduce_vec_(self, values: torch.Tensor) -> torch.Tensor:
        """In place: ``values`` (float32 CUDA, contiguous, any length) becomes the sum over all ranks
        (``ltr_p2p_allreduce_vec``: 65536 floats per launch, the same mailbox protocol)."""
        if not (values.is_cuda and values.dtype == torch.float32 and values.is_contiguous() and values.numel() >= 1):
            raise ValueError("values must be a non-empty contiguous float32 CUDA tensor")
        with _lib.on_device(values.device):
            self._check(self._lib.ltr_p2p_allreduce_vec(self._handle, values.data_ptr(), values.numel(),
                                                        _lib.raw_stream(values.device)))
        return values

    @property
    def handle(self):
        """The ``ltr_p2p *`` of this exchange (for entry points that fuse it: ``ltr_mlp_backward_allreduce``)."""
        return self._handle

    def timed_out(self) -> bool:
        """True if an exchange gave up waiting for a peer (synchronises the device)."""
        return self._lib.ltr_p2p_error(self._handle) == 1

    def close(self):
        if self._handle:
            self._lib.ltr_p2p_destroy(self._handle)
            self._handle = None


PeerExchange = PeerScalarExchange      # scalars (loss sums) and vectors (parameter gradients) share one mailbox


class _MeanOfShards(torch.autograd.Function):
    """``total / count`` as a function of the local per-query losses: ``total`` already holds the
    all-reduced sum; the backward pass hands every local query the broadcast gradient ``g / count`` (a
    stride-0 view, which the loss's backward reads in place).  ``count`` is the all-reduced query count
    (a device scalar) or, when the caller knows it, a Python number (no kernel for it)."""

    @staticmethod
    def forward(ctx, per_query, total, count):
        if isinstance(count, torch.Tensor):
            count = count.clamp(min=1.0)
            ctx.save_for_backward(count)
            ctx.inv_count = None
            out = total / count
        else:
            ctx.inv_count = 1.0 / max(float(count), 1.0)
            out = total * ctx.inv_count
        ctx.num_queries = per_query.shape[0]
        ctx.out_dtype = per_query.dtype
        return out.reshape(()).to(per_query.dtype)

    @staticmethod
    def backward(ctx, g):
        if ctx.inv_count is None:
            (count,) = ctx.saved_tensors
            gq = g / count.to(g.dtype)
        else:
            gq = g * ctx.inv_count
        return gq.to(ctx.out_dtype).reshape(1).expand(ctx.num_queries), None, None


def _accepts_loss_sum(loss_fn) -> bool:
    fwd = getattr(loss_fn, "forward", None)
    code = getattr(fwd, "__code__", None)
    return code is not None and "loss_sum" in code.co_varnames[:code.co_argcount]


def sharded_mean_loss(loss_fn: Callable, scores: torch.Tensor, relevance: torch.Tensor,
                      n: torch.Tensor, group=None, global_count: Optional[int] = None,
                      exchange: Optional[PeerScalarExchange] = None) -> torch.Tensor:
    """Global mean loss over the query shards of all ranks.

    ``loss_fn(scores, relevance, n)`` runs on the local shard only.  The returned scalar
    equals the mean over every query of every rank; its backward gives each rank
    ``d mean / d scores_local = grad_local / B_global`` (other ranks' terms do not depend
    on the local scores), so no gradient collective is needed for the scores.

    With CUDA inputs and a loss module of this package the local sum is produced by the loss
    kernel's own epilogue (``loss_sum``: one atomic add per query), so the ``[sum, count]``
    all-reduce follows the kernel with no reduction launch in between; the whole step (kernel,
    collective, backward) is CUDA-graph capturable.  ``global_count`` (the number of queries over
    all ranks, when the caller knows it -- e.g. a fixed global batch size) shrinks the collective to
    the sum alone and removes the count arithmetic from the step.  ``exchange`` (a
    :class:`PeerScalarExchange`, ranks of one node) carries the scalars over NVLink peer memory instead
    of the process group's all-reduce.  The float32 atomics make the
    last bits of the reported mean depend on the summation order; gradients do not depend on it.

    Parameter gradients of a model in front of the loss must be SUM-reduced across ranks
    (the 1 / B_global factor is already in the scalar's backward).  Under
    ``torch.nn.parallel.DistributedDataParallel``, which AVERAGES gradients, multiply the
    returned loss by the world size to compensate.
    """
    if scores.is_cuda and _accepts_loss_sum(loss_fn):
        if global_count is not None:
            buf = torch.zeros(1, dtype=torch.float32, device=scores.device)
        else:
            buf = torch.zeros(2, dtype=torch.float32, device=scores.device)
            buf[1:].fill_(float(scores.shape[0]))
        per_query = loss_fn(scores, relevance, n, loss_sum=buf[:1])
        if _world(group) > 1:
            if exchange is not None:
                exchange.all_reduce_(buf)
            else:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    else:
        per_query = loss_fn(scores, relevance, n)
        buf = global_sum_count(per_query, group)
    count = buf[1] if global_count is None else int(global_count)
    return _MeanOfShards.apply(per_query, buf[0], count)
