import sys, numpy as np, torch
sys.path.insert(0, '.')
from pytorchltr_b200 import _lib, _ops
import oracle
B, L = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
s = rng.standard_normal((B, L)).astype(np.float32)
n = rng.integers(L // 2, L + 1, B)
y = rng.integers(0, 5, (B, L)); y[np.arange(L)[None, :] >= n[:, None]] = 0
dev = torch.device('cuda')
st, yt, nt = (torch.as_tensor(a).to(dev) for a in (s, y, n))
loss, grad, _ = _ops.launch_loss(_lib.FAMILY_LAMBDA, _lib.LAM_NDCG2, st, yt, nt, 1.0, True)
torch.cuda.synchronize()
rl, rg = oracle.lambda_loss('ndcg2', s, y, n)
print('max loss err', np.abs(loss.cpu().numpy() - rl).max(), 'max grad err', np.abs(grad.cpu().numpy() - rg).max())
