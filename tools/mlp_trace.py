"""Event trace of one CTA of the MLP backward kernel (LTR_MLP_TRACE=1): where a tile's time goes."""
import ctypes
import os
import sys

os.environ["LTR_MLP_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P

lib = _lib.lib()
rows = 8192 * 200
args = P.make(rows, 136, 50, 10, exact=False)
ds = torch.randn(rows, device="cuda")
for _ in range(2):
    P.run_bwd(lib, *args, ds)
buf = (ctypes.c_longlong * (24 * 32))()
lib.ltr_mlp_trace_read.argtypes = [ctypes.c_void_p]
rc = lib.ltr_mlp_trace_read(buf)
assert rc == 0, rc
a = np.array(buf[:], dtype=np.int64).reshape(24, 32)
names = {0: "Kreq0", 1: "KreqN", 2: "MNreq0", 3: "MNreqN",  # TMA requests: K-major chunks, MN-major document groups
         4: "m1.buf", 5: "m1.k0", 6: "m1.iss", 7: "h1.rdy", 8: "dz2.rdy",
         9: "a2.rdy", 10: "mn.all", 11: "m2.iss", 16: "Z1", 17: "H1st", 18: "Z2", 19: "dZ2st", 20: "A2free", 21: "dW2done",
         22: "dH", 23: "dZ1st"}
t0 = a[2, 16]
print("cycles relative to tile 2's Z1-ready; one row per tile")
print("tile " + " ".join(f"{names[k]:>8s}" for k in sorted(names)))
for it in range(2, 10):
    print(f"{it:4d} " + " ".join(f"{(a[it, k] - t0) if a[it, k] else 0:8d}" for k in sorted(names)))
