import sys, numpy as np, torch
sys.path.insert(0, '.')
from bench import make_batch_numpy
from pytorchltr_b200 import _lib
lib = _lib.lib()
dev = torch.device('cuda')
L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fixed_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for B in (148, 592, 1184, 2368, 4096, 4144, 8192, 16384, 65536):
    s, y, n = make_batch_numpy(1, B, L)
    if fixed_n:
        n[:] = fixed_n
        y[np.arange(L)[None, :] >= n[:, None]] = 0
    ds, dy, dn = (torch.from_numpy(a).to(dev) for a in (s, y, n))
    loss = torch.empty(B, device=dev); grad = torch.empty(B, L, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        rc = lib.ltr_lambda(mode, ds.data_ptr(), dy.data_ptr(), 8, dn.data_ptr(), 8, B, L, 1.0, loss.data_ptr(), grad.data_ptr(), None, None, st)
        assert rc == 0
    for _ in range(5): run()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(20): run()
    st = torch.cuda.current_stream().cuda_stream
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    pairs = float((n * (n - 1) / 2).sum())
    print(f"B={B:6d} L={L} kernel {us:9.2f} us  {us / B * 1e3:8.2f} ns/query  {pairs / us / 1e6:8.3f} Tpairs/s (L2-warm)")
