"""Small MLP scorer run for compute-sanitizer: forward (with and without kept activations), both backward kernels."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib
import mlp_probe as P
lib = _lib.lib()
for (rows, F, H1, H2) in [(128 * 3 + 5, 136, 50, 10), (300, 24, 20, 5), (128 * 150 + 9, 136, 50, 10)]:
    args = P.make(rows, F, H1, H2, seed=2, exact=False)
    ds = torch.randn(rows, device="cuda")
    hz = P.make_hz(lib, args[0], args[1], args[3])
    P.run_fwd(lib, *args)
    P.run_fwd(lib, *args, hz=hz)
    a = P.run_bwd(lib, *args, ds)
    b = P.run_bwd(lib, *args, ds, hz=hz)
    print(rows, F, H1, H2, "ok", float((a - b).norm() / a.norm()))
# wide rows: streamed-W1 forward, column-slab backward from kept activations (no recomputing backward there)
for (rows, F, H1, H2) in [(128 * 5 + 77, 700, 50, 10), (1000, 292, 20, 5), (128 * 40 + 1, 1024, 40, 10),
                          (128 * 160 + 3, 320, 50, 10)]:
    args = P.make(rows, F, H1, H2, seed=3, exact=False)
    ds = torch.randn(rows, device="cuda")
    hz = P.make_hz(lib, args[0], args[1], args[3])
    s0 = P.run_fwd(lib, *args)
    s1 = P.run_fwd(lib, *args, hz=hz)
    b = P.run_bwd(lib, *args, ds, hz=hz)
    print(rows, F, H1, H2, "ok", bool((s0 == s1).all()), bool(torch.isfinite(b).all()))
