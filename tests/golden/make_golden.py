"""Generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference there and
does not travel to the GPU box):

    python tests/golden/make_golden.py

For seeded random padded batches it records, per loss, the reference's output
when called with float64 scores ("ref64": the arbiter),
the gradient its own CPU autograd produces, and its float32 output as shipped
("ref32").  For the metrics it records dcg / ndcg / arp and rank_by_score.
Scores are tie-free continuous draws (the reference breaks ties randomly);
padded relevance is zero-filled exactly as the reference's collate_fn does
(datasets/svmrank/svmrank.py:149-150) because evaluation/dcg.py:85 does not mask it.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LTR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from pytorchltr.evaluation import arp, dcg, ndcg  # noqa: E402
from pytorchltr.loss import (LambdaARPLoss1, LambdaARPLoss2,  # noqa: E402
                             LambdaNDCGLoss1, LambdaNDCGLoss2,
                             PairwiseDCGHingeLoss, PairwiseHingeLoss,
                             PairwiseLogisticLoss)
from pytorchltr.utils import rank_by_score  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

LOSSES = {
    "hinge": lambda sigma: PairwiseHingeLoss(),
    "dcg_hinge": lambda sigma: PairwiseDCGHingeLoss(),
    "logistic": lambda sigma: PairwiseLogisticLoss(sigma),
    "arp1": lambda sigma: LambdaARPLoss1(sigma),
    "arp2": lambda sigma: LambdaARPLoss2(sigma),
    "ndcg1": lambda sigma: LambdaNDCGLoss1(sigma),
    "ndcg2": lambda sigma: LambdaNDCGLoss2(sigma),
}

# (name, B, L, sigma, grades, special n values placed first)
CASES = [
    ("b6_l5", 6, 5, 1.0, 3, [5, 4, 0, 1, 2, 5]),
    ("b8_l37", 8, 37, 1.0, 5, [37, 0, 1, 2, 36]),
    ("b4_l128", 4, 128, 1.0, 5, [128, 64]),
    ("b3_l200", 3, 200, 0.7, 5, [200, 101]),
    ("b2_l300", 2, 300, 2.0, 4, [300]),
]


def make_batch(seed, B, L, grades, special_n):
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn(B, L, generator=g, dtype=torch.float32)
    n = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int64)
    for i, v in enumerate(special_n[:B]):
        n[i] = v
    rel = torch.randint(0, grades, (B, L), generator=g, dtype=torch.int64)
    rel[torch.arange(L)[None, :] >= n[:, None]] = 0
    return scores, rel, n


def main():
    for idx, (name, B, L, sigma, grades, special_n) in enumerate(CASES):
        scores, rel, n = make_batch(4242 + idx, B, L, grades, special_n)
        out = {"scores": scores.numpy(), "relevance": rel.numpy(), "n": n.numpy(),
               "sigma": np.float64(sigma)}
        for lname, ctor in LOSSES.items():
            fn = ctor(sigma)
            s64 = scores.double().requires_grad_(True)
            torch.manual_seed(0)
            l64 = fn(s64, rel, n)
            l64.sum().backward()
            out[f"{lname}_loss64"] = l64.detach().numpy()
            out[f"{lname}_grad64"] = s64.grad.numpy()
            s32 = scores.clone().requires_grad_(True)
            l32 = fn(s32, rel, n)
            l32.sum().backward()
            out[f"{lname}_loss32"] = l32.detach().numpy()
            out[f"{lname}_grad32"] = s32.grad.numpy()
        torch.manual_seed(0)
        out["ranking"] = rank_by_score(scores, n).numpy()
        for exp in (True, False):
            tag = "exp" if exp else "lin"
            out[f"dcg_all_{tag}"] = dcg(scores, rel, n, k=None, exp=exp).numpy()
            out[f"ndcg_all_{tag}"] = ndcg(scores, rel, n, k=None, exp=exp).numpy()
            for k in (1, 3, 10, 1000):
                out[f"dcg_k{k}_{tag}"] = dcg(scores, rel, n, k=k, exp=exp).numpy()
                out[f"ndcg_k{k}_{tag}"] = ndcg(scores, rel, n, k=k, exp=exp).numpy()
        out["arp"] = arp(scores, rel, n).numpy()
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
