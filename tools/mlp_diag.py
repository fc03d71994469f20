import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P
lib = _lib.lib()
for (rows, F) in [(128, 32), (128, 136), (256, 32), (256, 136), (1000, 136), (128 * 148 * 3, 136)]:
    H1, H2 = 50, 10
    args = P.make(rows, F, H1, H2, seed=1)
    ds = torch.randn(rows, device="cuda")
    out = P.run_bwd(lib, *args, ds).double()
    dw1, dw1_t, rest = P.ref_grads(*args, ds)
    o = [0, H1 * F, H1 * F + H1, H1 * F + H1 + H2 * H1, H1 * F + H1 + H2 * H1 + H2, H1 * F + H1 + H2 * H1 + 2 * H2]
    ref = torch.cat([dw1.reshape(-1), rest])
    names = ["dW1", "db1", "dW2", "db2", "dW3", "db3"]
    msg = []
    for k, nm in enumerate(names):
        a = out[o[k]:(o[k + 1] if k + 1 < len(o) else None)]
        b = ref[o[k]:(o[k + 1] if k + 1 < len(o) else None)]
        msg.append(f"{nm} {(a - b).abs().max().item() / (b.abs().max().item() + 1e-30):.1e}")
    print(f"rows={rows} F={F}: " + "  ".join(msg))
