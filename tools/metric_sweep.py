"""ndcg@10 / ndcg / arp / ListNet across list sizes: ns per query and GB/s show kernel-boundary steps."""
import sys
import torch
sys.path.insert(0, ".")
from pytorchltr_b200.evaluation import ndcg, arp
import pytorchltr_b200.loss as L
torch.manual_seed(0)
for Lq in (32, 64, 128, 200, 256, 260, 384, 512, 1024, 1028, 2048):
    B = max(512, min(65536, (1 << 26) // Lq))
    s = torch.randn(B, Lq, device="cuda")
    y = torch.randint(0, 5, (B, Lq), device="cuda")
    n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
    fns = {"ndcg@10": lambda: ndcg(s, y, n, k=10), "ndcg": lambda: ndcg(s, y, n), "arp": lambda: arp(s, y, n),
           "listnet": lambda: L.ListNetLoss()(s, y, n)}
    for name, fn in fns.items():
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
        gb = B * (12 * Lq + 12) / 1e9
        print(f"{name:8s} B={B:6d} L={Lq:5d}: {us:8.1f} us  {us * 1e3 / B:7.2f} ns/query  {gb / us * 1e6:7.0f} GB/s")
