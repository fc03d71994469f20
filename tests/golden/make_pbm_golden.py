"""Generates tests/golden/pbm_ref.npz by running the UNMODIFIED reference simulate_pbm
(pytorchltr/click_simulation/pbm.py:12-63) on seeded rankings.  The propensities it returns are
deterministic; with 0 / 1 relevance probabilities and eta = 0 so are the clicks.  Run in the build
container only:  python tests/golden/make_pbm_golden.py"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LTR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from pytorchltr.click_simulation.pbm import simulate_pbm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(11)
    out = {}
    for name, B, L in (("small", 4, 7), ("mid", 6, 40), ("long", 3, 300)):
        rankings = torch.stack([torch.randperm(L, generator=g) for _ in range(B)])
        ys = torch.randint(0, 5, (B, L), generator=g)
        n = torch.randint(1, L + 1, (B,), generator=g)
        n[0] = L
        out[f"{name}_rankings"] = rankings.numpy()
        out[f"{name}_ys"] = ys.numpy()
        out[f"{name}_n"] = n.numpy()
        probs = torch.tensor([0.05, 0.3, 0.5, 0.7, 0.95])
        for cutoff in (None, 3, 10):
            for eta in (0.0, 1.0, 2.0):
                _, props = simulate_pbm(rankings, ys, n, probs, cutoff, eta)
                out[f"{name}_props_c{cutoff}_e{eta}"] = props.numpy()
        hard = torch.tensor([0.0, 1.0, 0.0, 1.0, 1.0])
        for cutoff in (None, 5):
            clicks, _ = simulate_pbm(rankings, ys, n, hard, cutoff, 0.0)
            out[f"{name}_clicks_c{cutoff}"] = clicks.numpy()
    np.savez_compressed(os.path.join(HERE, "pbm_ref.npz"), **out)
    print("wrote pbm_ref.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
