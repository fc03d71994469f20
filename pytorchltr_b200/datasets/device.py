"""A ranking dataset held in GPU memory as one ragged block, collated by ``ltr_collate``.

The reference collates on the CPU, one Python iteration per sample
(``SVMRankDataset.collate_fn``, datasets/svmrank/svmrank.py:135-205) and the padded batch is then
copied to the device.  ``DeviceRankingDataset`` keeps ``features (N, F)``, ``relevance (N,)`` and the
query ``offsets (Q + 1,)`` on the device; ``collate(indices, max_list_size)`` returns the same
``(features, relevance, n, qid, sparse)`` batch the reference's collate_fn produces, already on the
device, for the three list samplers of ``datasets/list_sampler.py``:

* ``ListSampler(max_list_size)`` (default): first ``list_size`` documents, zero padding (``ltr_collate``);
* ``UniformSampler``: a uniformly random subset in random order -- the random permutation of a query's
  documents is the ``ltr_rank_by_score`` ranking of uniform random keys drawn from torch's generator;
* ``BalancedRelevanceSampler``: round-robin over the relevance grades in random grade order, random
  order inside a grade -- three stable rankings of random keys on the same kernel.

As in the reference the sampler is only consulted for queries longer than the batch's ``list_size``.
Sparse datasets (CSR features over the documents) collate into a torch sparse ``(B, list_size, F)``
tensor like the reference's ``sparse=True`` branch (svmrank.py:163-177, 198-203).
"""
from typing import Iterable, Optional, Sequence

import torch

from pytorchltr_b200 import _lib, _ops


class ListSampler:
    """First ``max_list_size`` documents (datasets/list_sampler.py:5-16)."""

    def __init__(self, max_list_size: Optional[int] = None):
        self._max_list_size = max_list_size

    def max_list_size(self, count: int) -> int:
        return count if self._max_list_size is None else min(self._max_list_size, count)

    def select(self, relevance: torch.Tensor, offsets: torch.Tensor, qidx: torch.Tensor,
               counts: torch.Tensor, cmax: int) -> Optional[torch.Tensor]:
        """``(B, cmax)`` int64 device tensor: row b lists document indices inside query b in sampling
        order (only its first ``list_size`` entries are used, and only if the query is longer than
        ``list_size``); None = take the documents in order."""
        return None


class UniformSampler(ListSampler):
    """Uniformly random documents in random order (datasets/list_sampler.py:19-27): the ranking of
    uniform random keys is a uniform random permutation."""

    def __init__(self, max_list_size: Optional[int] = None, generator: Optional[torch.Generator] = None):
        super().__init__(max_list_size)
        self.generator = generator

    def _rand(self, shape, device):
        if self.generator is not None and self.generator.device.type != device.type:
            # a CPU generator (the reference's convention): draw on the host, ship the keys
            return torch.rand(shape, generator=self.generator).to(device)
        kw = {} if self.generator is None else {"generator": self.generator}
        return torch.rand(shape, device=device, **kw)

    def select(self, relevance, offsets, qidx, counts, cmax):
        u = self._rand((counts.shape[0], cmax), counts.device)
        return _ops.rank_by_score(u, counts)


class BalancedRelevanceSampler(UniformSampler):
    """Round-robin over the relevance grades in random grade order, random order inside a grade
    (datasets/list_sampler.py:30-61).  Three rankings on the device: by a random key per document; then,
    stably, by a random key per (query, grade) -- documents are now grouped by grade in random grade order,
    shuffled inside a grade; then, stably, by the position inside the grade."""

    def select(self, relevance, offsets, qidx, counts, cmax):
        dev = counts.device
        B = counts.shape[0]
        pos = torch.arange(cmax, device=dev).unsqueeze(0)
        valid = pos < counts.unsqueeze(1)
        doc = (offsets[qidx].unsqueeze(1) + pos).clamp(max=relevance.shape[0] - 1)
        grade = torch.where(valid, relevance[doc], torch.zeros((), dtype=relevance.dtype, device=dev))
        p1 = _ops.rank_by_score(self._rand((B, cmax), dev), counts)               # shuffle
        g1 = grade.gather(1, p1)
        # one random key per (query, distinct grade): dense grade ids through a per-row sort
        gs, _ = g1.sort(dim=1)
        dense = torch.searchsorted(gs.contiguous(), g1.contiguous())           # first position of the grade
        gkey = self._rand((B, cmax), dev).gather(1, dense)
        p2 = _ops.rank_by_score(gkey, counts)                                     # stable: keeps the shuffle
        s2 = p1.gather(1, p2)
        g2 = g1.gather(1, p2)
        start = torch.ones_like(g2, dtype=torch.bool)
        start[:, 1:] = g2[:, 1:] != g2[:, :-1]
        first = torch.where(start, pos.expand(B, cmax), torch.zeros((), dtype=pos.dtype, device=dev))
        first = torch.cummax(first, dim=1).values
        inside = (pos - first).to(torch.float32)                                  # position inside the grade
        p3 = _ops.rank_by_score(-inside, counts)                                  # ascending, stable
        return s2.gather(1, p3)


class RankingBatch:
    """Same fields as ``pytorchltr.datasets.svmrank.SVMRankBatch`` (svmrank.py:30-40)."""

    def __init__(self, features, relevance, n, qid, sparse=False):
        self.features = features
        self.relevance = relevance
        self.n = n
        self.qid = qid
        self.sparse = sparse


class DeviceRankingDataset:
    def __init__(self, features: torch.Tensor, relevance: torch.Tensor, offsets: torch.Tensor,
                 qids: Optional[torch.Tensor] = None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceRankingDataset needs a CUDA device (sm_100a kernels, no CPU fallback)")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if features.dim() != 2:
            raise ValueError(f"features must be (N, F), got {tuple(features.shape)}")
        if relevance.dim() != 1 or relevance.shape[0] != features.shape[0]:
            raise ValueError("relevance must be (N,)")
        offsets_host = offsets.detach().to("cpu", torch.int64).contiguous()
        if offsets_host.dim() != 1 or offsets_host.numel() < 1 or int(offsets_host[-1]) != features.shape[0] \
                or int(offsets_host[0]) != 0 or bool((offsets_host[1:] < offsets_host[:-1]).any()):
            raise ValueError("offsets must be a non-decreasing (Q + 1,) vector from 0 to N")
        self.device = dev
        self.features = features.detach().to(dev, torch.float32).contiguous()
        self.relevance = relevance.detach().to(dev, torch.int64).contiguous()
        self.offsets = offsets_host.to(dev)
        self._counts_host = (offsets_host[1:] - offsets_host[:-1])     # list sizes, host copy: sizes the outputs
        q = self._counts_host.numel()
        self.qids = (torch.arange(q, dtype=torch.int64) if qids is None
                     else qids.detach().to("cpu", torch.int64).contiguous())
        if self.qids.numel() != q:
            raise ValueError("qids must have one entry per query")
        self._qids_dev = self.qids.to(dev)

    @classmethod
    def from_items(cls, items: Iterable, device=None) -> "DeviceRankingDataset":
        """From an iterable of reference-style items (``.features (n, F)``, ``.relevance (n,)``,
        ``.qid``), e.g. every ``SVMRankDataset[i]`` (dense)."""
        feats, rels, qids, offs = [], [], [], [0]
        for it in items:
            if getattr(it, "sparse", False):
                raise ValueError("sparse items are not supported")
            feats.append(torch.as_tensor(it.features, dtype=torch.float32))
            rels.append(torch.as_tensor(it.relevance, dtype=torch.int64))
            qids.append(int(it.qid))
            offs.append(offs[-1] + feats[-1].shape[0])
        return cls(torch.cat(feats, 0), torch.cat(rels, 0), torch.tensor(offs, dtype=torch.int64),
                   torch.tensor(qids, dtype=torch.int64), device=device)

    def __len__(self) -> int:
        return self._counts_host.numel()

    @classmethod
    def from_csr(cls, indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, num_features: int,
                 relevance: torch.Tensor, offsets: torch.Tensor, qids: Optional[torch.Tensor] = None,
                 device=None) -> "DeviceRankingDataset":
        """Sparse dataset: CSR features over the N documents (``indptr (N + 1,)``, ``indices (nnz,)``,
        ``values (nnz,)``), as ``SVMRankDataset(sparse=True)`` holds them (scipy CSR, svmrank.py:90-101)."""
        N = relevance.shape[0]
        if indptr.dim() != 1 or indptr.numel() != N + 1:
            raise ValueError("indptr must be (N + 1,)")
        self = cls(torch.zeros((N, 1)), relevance, offsets, qids, device=device)
        dev = self.device
        self.features = None
        self.num_features = int(num_features)
        self.indptr = indptr.detach().to(dev, torch.int64).contiguous()
        self.indices = indices.detach().to(dev, torch.int64).contiguous()
        self.values = values.detach().to(dev, torch.float32).contiguous()
        self.sparse = True
        return self

    def collate(self, indices: Sequence[int], max_list_size: Optional[int] = None,
                list_sampler: Optional[ListSampler] = None) -> RankingBatch:
        """The batch ``collate_fn(list_sampler)([dataset[i] for i in indices])`` of the reference, built on
        the device.  ``list_sampler`` defaults to ``ListSampler(max_list_size)``."""
        if list_sampler is None:
            list_sampler = ListSampler(max_list_size)
        elif max_list_size is not None:
            raise ValueError("pass max_list_size to the sampler")
        idx_host = torch.as_tensor(indices, dtype=torch.int64, device="cpu").reshape(-1)
        B = idx_host.numel()
        if B == 0:
            raise ValueError("empty batch")
        if bool(((idx_host < 0) | (idx_host >= len(self))).any()):
            raise IndexError("query index out of range")
        counts_host = self._counts_host[idx_host]
        cmax = int(counts_host.max())
        L = max(list_sampler.max_list_size(int(c)) for c in counts_host.tolist())    # svmrank.py:144-145
        dev = self.device
        idx = idx_host.to(dev, non_blocking=True)
        rel = torch.empty((B, L), dtype=torch.int64, device=dev)
        n = torch.empty(B, dtype=torch.int64, device=dev)
        sparse = getattr(self, "sparse", False)
        F = self.num_features if sparse else self.features.shape[1]
        if L == 0:
            n.zero_()
            feats = (torch.sparse_coo_tensor(torch.zeros((3, 0), dtype=torch.int64, device=dev),
                                             torch.zeros(0, device=dev), (B, 0, F)) if sparse
                     else torch.empty((B, 0, F), dtype=torch.float32, device=dev))
            return RankingBatch(feats, rel, n, self._qids_dev[idx], sparse)
        sel = None
        if cmax > L:
            sel = list_sampler.select(self.relevance, self.offsets, idx, counts_host.to(dev), cmax)
            if sel is not None:
                sel = sel.contiguous()
        lib = _lib.lib()
        st = _lib.raw_stream(dev)
        sel_ptr, sel_ld = (None, 0) if sel is None else (sel.data_ptr(), sel.shape[1])
        if not sparse:
            feats = torch.empty((B, L, F), dtype=torch.float32, device=dev)
            with _lib.on_device(dev):
                if sel is None:
                    rc = lib.ltr_collate(self.features.data_ptr(), self.relevance.data_ptr(), self.offsets.data_ptr(),
                                         idx.data_ptr(), B, L, F, feats.data_ptr(), rel.data_ptr(), n.data_ptr(),
                                         None, st)
                else:
                    rc = lib.ltr_collate_sampled(self.features.data_ptr(), self.relevance.data_ptr(),
                                                 self.offsets.data_ptr(), idx.data_ptr(), sel_ptr, sel_ld, B, L, F,
                                                 feats.data_ptr(), rel.data_ptr(), n.data_ptr(), None, st)
            _lib.check(rc)
            return RankingBatch(feats, rel, n, self._qids_dev[idx], False)
        # sparse: non-zeros of the selected documents -> output offsets (device scan; its total sizes the
        # output, one host read) -> COO triplets + values written by the kernel
        counts = counts_host.to(dev)
        pos = torch.arange(L, device=dev).unsqueeze(0)
        nb = counts.clamp(max=L).unsqueeze(1)
        if sel is None:
            local = pos.expand(B, L)
        else:
            local = torch.where(counts.unsqueeze(1) > L, sel[:, :L], pos.expand(B, L))
        doc = (self.offsets[idx].unsqueeze(1) + local).clamp(max=self.relevance.shape[0] - 1)
        row_nnz = torch.where(pos < nb, self.indptr[doc + 1] - self.indptr[doc],
                              torch.zeros((), dtype=torch.int64, device=dev)).reshape(-1)
        out_ptr = torch.zeros(B * L + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=out_ptr[1:])
        nnz = int(out_ptr[-1])
        coo = torch.empty((3, nnz), dtype=torch.int64, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            rc = lib.ltr_collate_sparse(self.indptr.data_ptr(), self.indices.data_ptr(), self.values.data_ptr(),
                                        self.relevance.data_ptr(), self.offsets.data_ptr(), idx.data_ptr(), sel_ptr,
                                        sel_ld, out_ptr.data_ptr(), B, L, nnz, coo.data_ptr(), val.data_ptr(),
                                        rel.data_ptr(), n.data_ptr(), st)
        _lib.check(rc)
        feats = torch.sparse_coo_tensor(coo, val, (B, L, F))
        return RankingBatch(feats, rel, n, self._qids_dev[idx], True)
