import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P
lib = _lib.lib()
torch.set_printoptions(precision=4, linewidth=200)
H1, H2, F = 50, 10, 136
o = [0, H1 * F, H1 * F + H1, H1 * F + H1 + H2 * H1, H1 * F + H1 + H2 * H1 + H2, H1 * F + H1 + H2 * H1 + 2 * H2]
for rows in (256, 384, 512, 1000):
    args = P.make(rows, F, H1, H2, seed=1)
    ds = torch.randn(rows, device="cuda")
    out = P.run_bwd(lib, *args, ds).double()
    dw1, dw1_t, rest = P.ref_grads(*args, ds)
    ref = torch.cat([dw1.reshape(-1), rest])
    print("rows", rows, "db2 out", out[o[3]:o[4]].cpu().numpy().round(4), "\n      db2 ref", ref[o[3]:o[4]].cpu().numpy().round(4))
    # per-tile reference of db2 to see which tiles are missing / doubled
    d = torch.float64
    x, w1, b1, w2, b2, w3, b3 = args
    z1 = x.to(d) @ w1.to(d).t() + b1.to(d); h1 = torch.relu(z1); z2 = h1 @ w2.to(d).t() + b2.to(d)
    dz2 = ds.to(d).reshape(-1, 1) * w3.to(d) * (z2 > 0)
    per_tile = torch.stack([dz2[t * 128:(t + 1) * 128].sum(0) for t in range((rows + 127) // 128)])
    diff = out[o[3]:o[4]] - ref[o[3]:o[4]]
    # least squares: which combination of per-tile sums explains the difference
    sol = torch.linalg.lstsq(per_tile.t(), diff.reshape(-1, 1)).solution.reshape(-1)
    print("      diff explained by per-tile coefficients", sol.cpu().numpy().round(3))
