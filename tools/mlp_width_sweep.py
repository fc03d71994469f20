"""Timing of the MLP scorer kernels over the feature width (MQ2007 46 -> 48, MSLR 136, Istella 220, Yahoo 699 -> 700):
inference forward, forward keeping activations, backward from them; GB/s of features and of all bytes moved.
python tools/mlp_width_sweep.py [F ...]"""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib  # noqa: E402
from mlp_probe import make, make_hz  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.lib()
    widths = [int(a) for a in sys.argv[1:]] or [48, 136, 220, 288, 320, 512, 700, 1024]
    H1, H2 = 50, 10
    st = torch.cuda.current_stream().cuda_stream
    for F in widths:
        rows = min(8192 * 200, (1 << 30) // (F * 4) // 128 * 128)      # at most 1 GiB of features (> L2)
        args = make(rows, F, H1, H2, exact=False)
        x, w1, b1, w2, b2, w3, b3 = args
        ds = torch.randn(rows, device="cuda")
        hz = make_hz(lib, x, w1, w2)
        out = torch.empty(rows, device="cuda")
        n = lib.ltr_mlp_grad_len(F, H1, H2)
        grads = torch.empty(n, device="cuda")
        wsb = lib.ltr_mlp_workspace_bytes(F, H1, H2)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")

        def fwd(hzp):
            return lib.ltr_mlp_scores(x.data_ptr(), rows, F, w1.data_ptr(), b1.data_ptr(), H1, w2.data_ptr(),
                                      b2.data_ptr(), H2, w3.data_ptr(), b3.data_ptr(), out.data_ptr(), hzp, st)

        def bwd():
            return lib.ltr_mlp_backward(x.data_ptr(), rows, F, w1.data_ptr(), b1.data_ptr(), H1, w2.data_ptr(),
                                        b2.data_ptr(), H2, w3.data_ptr(), b3.data_ptr(), hz.data_ptr(),
                                        ds.data_ptr(), grads.data_ptr(), ws.data_ptr(), wsb, st)
        xb = rows * F * 4
        hb = hz.numel() * 4
        t_inf = timed(lambda: fwd(None))
        t_keep = timed(lambda: fwd(hz.data_ptr()))
        t_bwd = timed(bwd)
        lin = torch.nn.Sequential(torch.nn.Linear(F, H1), torch.nn.ReLU(), torch.nn.Linear(H1, H2), torch.nn.ReLU(),
                                  torch.nn.Linear(H2, 1)).cuda()
        torch.backends.cuda.matmul.allow_tf32 = False
        for _ in range(2):
            lin(x).backward(ds.reshape(-1, 1))
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            lin(x).backward(ds.reshape(-1, 1))
        torch.cuda.synchronize()
        t_torch = (time.perf_counter() - t) / 3 * 1e3
        print(f"F={F:5d} rows={rows:8d}: inference {t_inf * 1e3:7.1f} us ({xb / t_inf / 1e6:5.0f} GB/s)  "
              f"keep {t_keep * 1e3:7.1f} us ({(xb + hb) / t_keep / 1e6:5.0f} GB/s)  "
              f"backward {t_bwd * 1e3:7.1f} us ({xb / t_bwd / 1e6:5.0f} GB/s of features)  "
              f"fwd+bwd {(t_keep + t_bwd) * 1e3:7.1f} us vs torch fp32 modules {t_torch * 1e3:8.1f} us", flush=True)
        del x, hz, args, lin
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
