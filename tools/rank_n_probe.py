import sys
import torch
sys.path.insert(0, ".")
from pytorchltr_b200.utils import rank_by_score
torch.manual_seed(0)
B = 32768
for Lq, nval in ((1024, 1024), (1024, 768), (1024, 520), (1024, 300), (640, 640), (640, 330), (520, 520), (520, 270)):
    s = torch.randn(B, Lq, device="cuda")
    n = torch.full((B,), nval, device="cuda")
    for kind in ("randn", "grid"):
        ss = s if kind == "randn" else (torch.arange(Lq, device="cuda").float().flip(0)[None, :].expand(B, Lq).contiguous())
        for _ in range(3):
            rank_by_score(ss, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rank_by_score(ss, n)
        e1.record()
        torch.cuda.synchronize()
        print(f"L={Lq:5d} n={nval:5d} {kind:6s}: {e0.elapsed_time(e1) / 10 * 1e3 * 1e3 / B:7.2f} ns/query")
