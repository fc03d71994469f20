"""The getting-started training step at the batch sizes the reference's documentation uses (batch_size = 16 lists of
up to 200 documents, 136 features): MLPRanker + PairwiseHingeLoss forward + backward, eager and as a CUDA graph,
against the same model from torch.nn.Linear layers with the reference-style loss replaced by ours (so that only the
scorer differs) -- latency, not bandwidth.  python tools/mlp_small_batch.py [B ...]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import pytorchltr_b200.loss as L  # noqa: E402
from pytorchltr_b200.fused import MLPRanker  # noqa: E402


def wall(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [16, 64, 256, 1024]
    Lq, F = 200, 136
    loss_fn = L.PairwiseHingeLoss()
    for B in Bs:
        torch.manual_seed(0)
        xs = torch.randn(B, Lq, F, device="cuda")
        ys = torch.randint(0, 5, (B, Lq), device="cuda")
        n = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
        fused = MLPRanker(F).cuda()
        plain = torch.nn.Sequential(torch.nn.Linear(F, 50), torch.nn.ReLU(), torch.nn.Linear(50, 10), torch.nn.ReLU(),
                                    torch.nn.Linear(10, 1)).cuda()

        def step(model):
            for p in model.parameters():
                p.grad = None
            loss_fn(model(xs), ys, n).mean().backward()
        res = {}
        for name, model in (("fused", fused), ("plain", plain)):
            res[name + "_eager"] = wall(lambda: step(model))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(model)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(model)
            res[name + "_graph"] = wall(g.replay)
        print(f"B={B:5d} (L={Lq}, F={F}): MLPRanker step eager {res['fused_eager']:7.1f} us, graph {res['fused_graph']:7.1f} us"
              f" | torch.nn.Linear layers eager {res['plain_eager']:7.1f} us, graph {res['plain_graph']:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
