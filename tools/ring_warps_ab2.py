import os, sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
tag = os.environ.get("LTR_RING_WARPS", "default")
for B, Lq in ((8192, 384), (4096, 516), (4096, 640), (2048, 768), (2048, 1024)):
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for name in ("LambdaNDCGLoss2", "PairwiseLogisticLoss"):
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
        print(f"warps={tag:8s} {name:22s} B={B} L={Lq}: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us")
