import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P
lib = _lib.lib()
F, H1, H2 = 136, 50, 10
for rows in (128 * 148 * 3, 40000):
    args = P.make(rows, F, H1, H2, seed=5, exact=False)
    ds = torch.randn(rows, device="cuda")
    full = P.run_bwd(lib, *args, ds).double()
    acc = torch.zeros_like(full)
    x = args[0]
    step = 128 * 100          # <= 148 tiles per launch: one tile per CTA
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        acc += P.run_bwd(lib, x[r0:r1].contiguous(), *args[1:], ds[r0:r1].contiguous()).double()
    err = (full - acc).abs()
    print(f"rows={rows}: multi-tile vs sum of single-tile launches: max abs diff {err.max().item():.3e} "
          f"(max |grad| {acc.abs().max().item():.2f}), worst index {int(err.argmax())}")
