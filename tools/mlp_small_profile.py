"""cProfile of the eager small-batch step (host overhead): python tools/mlp_small_profile.py"""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
import pytorchltr_b200.loss as L  # noqa: E402
from pytorchltr_b200.fused import MLPRanker  # noqa: E402

B, Lq, F = 16, 200, 136
xs = torch.randn(B, Lq, F, device="cuda")
ys = torch.randint(0, 5, (B, Lq), device="cuda")
n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
model = MLPRanker(F).cuda()
loss_fn = L.PairwiseHingeLoss()


def step():
    for p in model.parameters():
        p.grad = None
    loss_fn(model(xs), ys, n).mean().backward()


for _ in range(50):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumtime").print_stats(40)
