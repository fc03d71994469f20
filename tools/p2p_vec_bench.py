"""Data-parallel MLP ranker step under torchrun (one rank per GPU): the parameter-gradient exchange fused into the
backward pass's final reduction (ltr_mlp_backward_allreduce, NVLink peer memory) against the same backward pass
followed by an NCCL all-reduce, and the bare vector exchange against NCCL at the gradient's length.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/p2p_vec_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from pytorchltr_b200 import _lib  # noqa: E402
from pytorchltr_b200.distributed import PeerExchange  # noqa: E402
from mlp_probe import make, make_hz, run_fwd  # noqa: E402


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    lib = _lib.lib()
    ex = PeerExchange()
    st = torch.cuda.current_stream().cuda_stream
    F, H1, H2 = 136, 50, 10
    rows = 8192 * 200 // world // 128 * 128                # the c4mlp batch split over the ranks
    args = make(rows, F, H1, H2, seed=rank, exact=False)
    x, w1, b1, w2, b2, w3, b3 = args
    ds = torch.randn(rows, device=dev)
    hz = make_hz(lib, x, w1, w2)
    run_fwd(lib, *args, hz=hz)
    n = lib.ltr_mlp_grad_len(F, H1, H2)
    grads = torch.empty(n, device=dev)
    wsb = lib.ltr_mlp_workspace_bytes(F, H1, H2)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    head = (x.data_ptr(), rows, F, w1.data_ptr(), b1.data_ptr(), H1, w2.data_ptr(), b2.data_ptr(), H2, w3.data_ptr(),
            b3.data_ptr(), hz.data_ptr(), ds.data_ptr(), grads.data_ptr(), ws.data_ptr(), wsb)

    def local():
        _lib.check(lib.ltr_mlp_backward(*head, st))

    def nccl():
        _lib.check(lib.ltr_mlp_backward(*head, st))
        dist.all_reduce(grads)

    def fused():
        _lib.check(lib.ltr_mlp_backward_allreduce(*head, ex.handle, st))
    nccl()
    ref = grads.clone()
    fused()
    # two ranks: a + b in either order, the same bits; more ranks: NCCL adds in another order (rounding only)
    err = ((grads - ref).abs().max() / ref.abs().max()).item()
    same = bool(torch.equal(grads, ref)) if world == 2 else err < 1e-5
    chk = [grads.clone() for _ in range(world)]
    dist.all_gather(chk, grads)
    identical = all(bool(torch.equal(c, chk[0])) for c in chk)        # every rank holds the same bits
    t_local, t_nccl, t_fused = timed(local), timed(nccl), timed(fused)
    v = torch.randn(n, device=dev)
    t_vec = timed(lambda: ex.all_reduce_vec_(v))
    t_ncv = timed(lambda: dist.all_reduce(v))
    bad = ex.timed_out()
    if rank == 0:
        print(f"N={world} rows/rank={rows} grad floats={n}: backward alone {t_local:.1f} us | + NCCL all-reduce "
              f"{t_nccl:.1f} us (+{t_nccl - t_local:.1f}) | exchange fused into the reduction {t_fused:.1f} us "
              f"(+{t_fused - t_local:.1f}) | bare vector exchange {t_vec:.1f} us vs NCCL {t_ncv:.1f} us | "
              f"same sums as NCCL: {same} (max err / max |g| = {err:.1e}) | bit-identical on every rank: {identical} | "
              f"timed out: {bad}", flush=True)
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
