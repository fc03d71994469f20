"""Times ltr_collate (kernel only, indices resident on the device) at the MSLR-WEB30K shape."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from pytorchltr_b200 import _lib
from pytorchltr_b200.datasets import DeviceRankingDataset

Q, F = 8192, 136
rng = np.random.default_rng(0)
counts = rng.integers(100, 201, size=Q)
offs = np.concatenate([[0], np.cumsum(counts)])
ds = DeviceRankingDataset(torch.randn(int(offs[-1]), F), torch.randint(0, 5, (int(offs[-1]),)), torch.from_numpy(offs))
dev = ds.device
idx = torch.from_numpy(rng.permutation(Q)).to(dev)
L = int(counts.max())
feats = torch.empty(Q, L, F, device=dev)
rel = torch.empty(Q, L, dtype=torch.int64, device=dev)
n = torch.empty(Q, dtype=torch.int64, device=dev)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.ltr_collate(ds.features.data_ptr(), ds.relevance.data_ptr(), ds.offsets.data_ptr(),
                               idx.data_ptr(), Q, L, F, feats.data_ptr(), rel.data_ptr(), n.data_ptr(), None, st))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byt = sum(4 * F * (c + L) + 8 * (c + L) + 24 for c in counts)
print("ltr_collate (8192 queries, F=136, L=%d): %.1f us, %.0f GB/s algorithmic (read + write)" % (L, ms * 1e3, byt / ms / 1e6))
