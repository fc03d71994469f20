"""pytorchltr_b200: the pytorchltr loss / metric hot path as hand-written sm_100a CUDA.

Drop-in for ``pytorchltr.loss.*``, ``pytorchltr.evaluation.{ndcg, dcg, arp}`` and
``pytorchltr.utils.rank_by_score`` with the reference's own call signature
``(scores, relevance, n)`` over padded ``(B, L)`` batches.  Host code is Python /
PyTorch (device memory, streams, autograd plumbing); all arithmetic runs in
``csrc/libltr_sm100.so`` (C ABI in ``include/ltr_sm100.h``).  There is no CPU
compute path: CPU tensors are staged through the GPU, and a missing extension or
missing GPU raises.

Beyond the drop-in path: ``fused`` (linear scorer + ListNet in one pass over the features),
``datasets`` (device-resident ragged dataset, GPU collation) and ``click_simulation`` (the
reference's PBM click simulators).
"""
from pytorchltr_b200 import click_simulation, datasets, evaluation, fused, loss, utils  # noqa: F401

__version__ = "0.1.0"
