import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P
lib = _lib.lib()
res = {}
for (rows, F) in [(128, 32), (128, 136), (256, 32), (128, 64), (128, 8)]:
    args = P.make(rows, F, 50, 10, seed=1)
    ds = torch.randn(rows, device="cuda")
    out = P.run_bwd(lib, *args, ds)
    dw1, dw1_t, rest = P.ref_grads(*args, ds)
    res[f"out_{rows}_{F}"] = out[:50 * F].reshape(50, F).cpu().numpy()
    res[f"ref_{rows}_{F}"] = dw1_t.cpu().numpy()
np.savez("gpurun_out/mlp_diag.npz", **res)
print("saved")
