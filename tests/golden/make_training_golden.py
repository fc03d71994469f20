"""Generates tests/golden/train_linear.npz by running the UNMODIFIED reference on CPU.

The reference's own end-to-end check is a training loop (tests/test_integration.py:18-63,
examples/01-basic-usage.py:68-75): `loss_fn(model(xs), ys, n).mean().backward(); optimizer.step()`.
This script runs that loop -- a `torch.nn.Linear(F, 1)` ranker, full-batch SGD, 12 steps -- with
every reference loss on one seeded synthetic batch and records the loss trajectory, the final
parameters and the reference's ndcg@10 / arp of the trained model.  The GPU test
(tests/test_gpu_training_parity.py) replays the same loop with pytorchltr_b200 and must land on
the same numbers.  Run in the build container only:

    python tests/golden/make_training_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LTR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from pytorchltr.evaluation import arp, ndcg  # noqa: E402
from pytorchltr.loss import (LambdaARPLoss1, LambdaARPLoss2,  # noqa: E402
                             LambdaNDCGLoss1, LambdaNDCGLoss2,
                             PairwiseDCGHingeLoss, PairwiseHingeLoss,
                             PairwiseLogisticLoss)

HERE = os.path.dirname(os.path.abspath(__file__))
LOSSES = {
    "hinge": PairwiseHingeLoss, "dcg_hinge": PairwiseDCGHingeLoss, "logistic": PairwiseLogisticLoss,
    "arp1": LambdaARPLoss1, "arp2": LambdaARPLoss2, "ndcg1": LambdaNDCGLoss1, "ndcg2": LambdaNDCGLoss2,
}
# per-loss learning rates: the losses differ by orders of magnitude in scale
LR = {"hinge": 2e-4, "dcg_hinge": 0.5, "logistic": 5e-4, "arp1": 2e-4, "arp2": 2e-4, "ndcg1": 0.05, "ndcg2": 0.5}
STEPS = 12


def main():
    g = torch.Generator().manual_seed(42)
    B, L, F = 48, 40, 8
    xs = torch.randn(B, L, F, generator=g)
    n = torch.randint(3, L + 1, (B,), generator=g)
    n[0], n[1] = L, 1
    true_w = torch.randn(F, generator=g)
    noise = torch.randn(B, L, generator=g)
    ys = torch.clamp(((xs @ true_w) * 0.8 + noise * 0.5 + 1.5).round(), 0, 4).long()
    mask = torch.arange(L)[None, :] >= n[:, None]
    ys[mask] = 0                      # collate_fn zero-pads (svmrank.py:149-150)
    xs[mask] = 0.0
    w0 = torch.randn(1, F, generator=g) * 0.1
    b0 = torch.zeros(1)
    out = {"xs": xs.numpy(), "ys": ys.numpy(), "n": n.numpy(), "w0": w0.numpy(), "b0": b0.numpy(),
           "steps": np.array(STEPS)}
    for name, cls in LOSSES.items():
        model = torch.nn.Linear(F, 1)
        with torch.no_grad():
            model.weight.copy_(w0)
            model.bias.copy_(b0)
        opt = torch.optim.SGD(model.parameters(), lr=LR[name])
        loss_fn = cls()
        traj = []
        for _ in range(STEPS):
            opt.zero_grad()
            loss = loss_fn(model(xs), ys, n).mean()
            loss.backward()
            opt.step()
            traj.append(float(loss))
        with torch.no_grad():
            scores = model(xs)
            out[f"{name}_ndcg10"] = ndcg(scores, ys, n, k=10).numpy()
            out[f"{name}_arp"] = arp(scores, ys, n).numpy()
        out[f"{name}_loss"] = np.array(traj)
        out[f"{name}_w"] = model.weight.detach().numpy().copy()
        out[f"{name}_b"] = model.bias.detach().numpy().copy()
        out[f"{name}_lr"] = np.array(LR[name])
        print(name, "loss", traj[0], "->", traj[-1], "ndcg@10", float(out[f"{name}_ndcg10"].mean()))
    np.savez_compressed(os.path.join(HERE, "train_linear.npz"), **out)


if __name__ == "__main__":
    main()
