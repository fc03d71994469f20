import sys
import torch
sys.path.insert(0, ".")
import pytorchltr_b200.loss as L
torch.manual_seed(0)
B, Lq = 16384, 200
s = torch.randn(B, Lq, device="cuda")
y = torch.randint(0, 5, (B, Lq), device="cuda")
n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
fn = L.LambdaNDCGLoss2()
for _ in range(3):
    fn(s, y, n)
torch.cuda.synchronize()
