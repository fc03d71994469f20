import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bench import make_batch_numpy
import pytorchltr_b200.loss as L
from pytorchltr_b200 import _lib
B, Ln = 4096, 128
s, y, n = make_batch_numpy(1, B, Ln)
hs, hy, hn = torch.from_numpy(s).pin_memory(), torch.from_numpy(y).pin_memory(), torch.from_numpy(n).pin_memory()
dev = torch.device('cuda')
fn = L.LambdaNDCGLoss2()
def t(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
print('H2D 3 tensors   us', t(lambda: (hs.to(dev, non_blocking=True), hy.to(dev, non_blocking=True), hn.to(dev, non_blocking=True))))
ds, dy, dn = hs.to(dev), hy.to(dev), hn.to(dev)
g = torch.empty(B, Ln, device=dev)
hg = torch.empty(B, Ln, pin_memory=True)
print('D2H 2MB pinned  us', t(lambda: hg.copy_(g, non_blocking=True)))
print('D2H 2MB .cpu()  us', t(lambda: g.cpu()))
def full():
    hs.grad = None
    x = hs.requires_grad_(True)
    out = fn(x, hy, hn); out.sum().backward()
print('full public API us', t(full))
def fwd_only():
    with torch.no_grad(): fn(hs, hy, hn)
print('fwd only no grad us', t(fwd_only))
# C-ABI host entry
import ctypes
lib = _lib.lib()
nbytes = lib.ltr_host_workspace_bytes(B, Ln)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
hl = torch.empty(B).pin_memory()
st = torch.cuda.current_stream().cuda_stream
def cabi():
    rc = lib.ltr_loss_host(1, 3, hs.data_ptr(), hy.data_ptr(), hn.data_ptr(), B, Ln, 1.0, hl.data_ptr(), hg.data_ptr(), ws.data_ptr(), ctypes.c_size_t(nbytes), st)
    assert rc == 0
    torch.cuda.current_stream().synchronize()
print('ltr_loss_host + sync us', t(cabi))
