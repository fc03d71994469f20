"""LambdaNDCGLoss2 / PairwiseLogisticLoss forward (+ saved gradient) across list sizes around the kernel boundaries:
ns per valid pair shows where a dispatch threshold is misplaced.  python tools/lambda_sweep.py"""
import sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
for Lq in (32, 64, 96, 128, 132, 160, 200, 256, 384, 512, 516, 768, 1024, 1028, 1536, 2048):
    B = max(256, min(65536, (1 << 31) // (Lq * Lq * 4)))
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    pairs = float((n.double() * (n.double() - 1) / 2).sum())
    for name in ("LambdaNDCGLoss2", "PairwiseLogisticLoss", "LambdaARPLoss2"):
        fn = getattr(L, name)()
        for _ in range(3):
            fn(s, y, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(s, y, n)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"{name:22s} B={B:6d} L={Lq:5d}: {us:9.1f} us  {us * 1e3 / pairs * 1e3:7.3f} ps/pair  {us * 1e3 / B:8.1f} ns/query")
