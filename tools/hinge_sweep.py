"""PairwiseHingeLoss / PairwiseDCGHingeLoss forward (+ saved gradient) at several list sizes: sorted O(n log n) kernel
against the O(n^2) pair kernels (LTR_HINGE=pairs).  Run once per setting: python tools/hinge_sweep.py"""
import os
import sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
for B, Lq in ((8192, 136), (8192, 160), (8192, 200), (8192, 256), (4096, 320), (4096, 400), (2048, 512), (1024, 1024)):
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for name in ("PairwiseHingeLoss", "PairwiseDCGHingeLoss"):
        fn = getattr(L, name)()
        for _ in range(3):
            fn(s, y, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn(s, y, n)
        e1.record()
        torch.cuda.synchronize()
        print(f"LTR_HINGE={os.environ.get('LTR_HINGE', 'default'):8s} {name:22s} B={B:5d} L={Lq:4d}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us")
