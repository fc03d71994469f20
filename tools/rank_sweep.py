import sys
import torch
sys.path.insert(0, ".")
from pytorchltr_b200.utils import rank_by_score
from pytorchltr_b200.evaluation import arp
torch.manual_seed(0)
for Lq in (128, 256, 260, 512, 520, 640, 768, 900, 1024):
    B = 32768
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    for name, fn in (("rank_by_score", lambda: rank_by_score(s, n)), ("arp", lambda: arp(s, y, n))):
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
        print(f"{name:14s} B={B} L={Lq:5d}: {us:8.1f} us  {us * 1e3 / B:7.2f} ns/query")
