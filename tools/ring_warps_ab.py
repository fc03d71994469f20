import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
import oracle
torch.manual_seed(0)
tag = os.environ.get("LTR_RING_WARPS", "default")
for Lq in (132, 160, 200, 256):
    B = 16384
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for name in ("LambdaNDCGLoss2", "PairwiseLogisticLoss", "PairwiseHingeLoss"):
        fn = getattr(L, name)()
        for _ in range(3):
            out = fn(s, y, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(s, y, n)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        chk = ""
        if name == "LambdaNDCGLoss2":
            ref, _ = oracle.lambda_loss("ndcg2", s[:64].cpu().numpy(), y[:64].cpu().numpy(), n[:64].cpu().numpy())
            chk = f" parity {np.abs(out[:64].cpu().double().numpy() - ref).max() / np.abs(ref).max():.1e}"
        print(f"warps={tag:8s} {name:22s} B={B} L={Lq}: {us:8.1f} us{chk}")
