"""The c5 list size (512) at the per-rank batch of the 8-GPU split (8192 queries) against the full batch (65536):
time per query with 2 / 4 (default) / 8 warps per query (LTR_RING_WARPS).  One process per setting."""
import os, sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
tag = os.environ.get("LTR_RING_WARPS", "default")
fn = L.LambdaNDCGLoss2()
for B in (65536, 8192):
    Lq = 512
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for _ in range(3):
        fn(s, y, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if B > 10000 else 30
    e0.record()
    for _ in range(reps):
        fn(s, y, n)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"warps={tag:8s} B={B:6d} L={Lq}: {us:8.1f} us  {us * 1e3 / B:7.2f} ns/query", flush=True)
