"""A ranking dataset held in GPU memory as one ragged block, collated by ``ltr_collate``.

The reference collates on the CPU, one Python iteration per sample
(``SVMRankDataset.collate_fn``, datasets/svmrank/svmrank.py:135-205) and the padded batch is then
copied to the device.  ``DeviceRankingDataset`` keeps ``features (N, F)``, ``relevance (N,)`` and the
query ``offsets (Q + 1,)`` on the device; ``collate(indices, max_list_size)`` returns the same
``(features, relevance, n, qid, sparse)`` batch the reference's collate_fn produces with its
default ``ListSampler(max_list_size)`` (first ``list_size`` documents, zero padding), already on the
device.  Random list samplers are not supported.
"""
from typing import Iterable, Optional, Sequence

import torch

from pytorchltr_b200 import _lib


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

    def collate(self, indices: Sequence[int], max_list_size: Optional[int] = None) -> RankingBatch:
        """The batch ``collate_fn(ListSampler(max_list_size))([dataset[i] for i in indices])`` of the
        reference, built on the device."""
        idx_host = torch.as_tensor(indices, dtype=torch.int64, device="cpu").reshape(-1)
        B = idx_host.numel()
        F = self.features.shape[1]
        if B == 0:
            raise ValueError("empty batch")
        if bool(((idx_host < 0) | (idx_host >= len(self))).any()):
            raise IndexError("query index out of range")
        counts = self._counts_host[idx_host]
        if max_list_size is not None:
            counts = counts.clamp(max=int(max_list_size))
        L = max(int(counts.max()), 1) if int(counts.max()) > 0 else 0
        dev = self.device
        idx = idx_host.to(dev, non_blocking=True)
        feats = torch.empty((B, L, F), dtype=torch.float32, device=dev)
        rel = torch.empty((B, L), dtype=torch.int64, device=dev)
        n = torch.empty(B, dtype=torch.int64, device=dev)
        if L > 0:
            with torch.cuda.device(dev):
                rc = _lib.lib().ltr_collate(self.features.data_ptr(), self.relevance.data_ptr(),
                                            self.offsets.data_ptr(), idx.data_ptr(), B, L, F, feats.data_ptr(),
                                            rel.data_ptr(), n.data_ptr(), None,
                                            torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc)
        else:
            n.zero_()
        return RankingBatch(feats, rel, n, self._qids_dev[idx], False)
