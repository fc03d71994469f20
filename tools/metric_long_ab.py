import os, sys
import torch
sys.path.insert(0, ".")
from pytorchltr_b200.evaluation import ndcg, arp
torch.manual_seed(0)
tag = os.path.basename(os.environ.get("LTR_SM100_LIB", "default"))
for Lq in (260, 384, 512, 640, 1024):
    B = 32768
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for name, fn in (("ndcg", lambda: ndcg(s, y, n)), ("arp", lambda: arp(s, y, n))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"{tag:14s} {name:5s} B={B} L={Lq:5d}: {us:8.1f} us  {us * 1e3 / B:7.2f} ns/query")
