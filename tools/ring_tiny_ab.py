import os, sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
tag = os.path.basename(os.environ.get("LTR_SM100_LIB", "default"))
for B, Lq in ((8192, 288), (8192, 320), (8192, 384), (4096, 448), (4096, 512)):
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
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
        print(f"{tag:16s} {name:22s} B={B} L={Lq}: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us")
