#!/usr/bin/env python
"""Times the UNMODIFIED reference (rjagerman/pytorchltr, installed into git-ignored baseline/_ref
by `pip install --no-deps --target baseline/_ref`) on the loss / metric hot path:

  (a) on the host CPU, all cores (torch.set_num_threads(os.cpu_count())), chunked over B;
  (b) on the B200 as it would run today: the same Python, eager ATen kernels, chunked so that the
      ~78 L^2 bytes per query of pair tensors fit in HBM.

Step = `loss_fn(scores, relevance, n).mean().backward()` (examples/01-basic-usage.py:72-74) or one
metric call.  Same synthetic inputs as bench.py (make_batch_numpy).  Writes one JSON document to
stdout; the committed copy is profiles/reference_timing.json, which bench.py quotes as
`cpu_baseline.python_reference` / `gpu_eager_reference` when baseline/_ref is absent at run time.

    python tools/time_reference.py [--configs c2,ns,c3,c5,c4m,c4a] [--budget-s 20] > profiles/reference_timing.json
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import bench  # noqa: E402  (make_batch_numpy, CONFIGS)


def reference_callable(cfg):
    import pytorchltr.evaluation as ref_eval
    import pytorchltr.loss as ref_loss
    if "metric" in cfg:
        if cfg["metric"] == "arp":
            return lambda s, y, n: ref_eval.arp(s, y, n), False
        fn = getattr(ref_eval, cfg["metric"])
        k = cfg.get("k")
        return lambda s, y, n: fn(s, y, n, k=k), False
    if not hasattr(ref_loss, cfg["loss"]):
        return None, True
    return getattr(ref_loss, cfg["loss"])(), True


def time_arm(cfg, device, chunk, budget_s):
    import torch
    fn, is_loss = reference_callable(cfg)
    if fn is None:
        return {"unavailable": f"{cfg['loss']} does not exist in the reference (SURVEY.md F2)"}
    L = cfg["L"]
    s_np, y_np, n_np = bench.make_batch_numpy(1234, chunk, L, cfg.get("skew", False))
    s = torch.from_numpy(s_np).to(device).requires_grad_(is_loss)
    y, n = torch.from_numpy(y_np).to(device), torch.from_numpy(n_np).to(device)

    def step():
        if is_loss:
            s.grad = None
            fn(s, y, n).mean().backward()
        else:
            fn(s, y, n)
        if device.type == "cuda":
            torch.cuda.synchronize()

    step()                                   # warm-up (allocator, kernels)
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 30):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    out = {"queries_per_s": chunk / med, "best_queries_per_s": chunk / min(times), "chunk": chunk,
           "median_s_per_chunk": med, "chunks_timed": len(times)}
    if device.type == "cuda":
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
        torch.cuda.reset_peak_memory_stats()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,ns,c3,c5,c4m,c4a")
    ap.add_argument("--budget-s", type=float, default=15.0)
    ap.add_argument("--no-gpu", action="store_true")
    args = ap.parse_args()
    import torch
    import pytorchltr
    torch.set_num_threads(os.cpu_count() or 1)
    doc = {"reference": "rjagerman/pytorchltr 0.2.1 (unmodified, baseline/_ref)", "module": pytorchltr.__file__,
           "torch": torch.__version__, "cpu_cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
           "step": "loss_fn(scores, relevance, n).mean().backward() on a chunk of the config's batch "
                   "(queries are independent, so per-query rates extrapolate exactly); metrics: one call",
           "gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None, "configs": {}}
    for name in args.configs.split(","):
        cfg = bench.CONFIGS[name]
        L = cfg["L"]
        pair = "metric" not in cfg and cfg["loss"] != "ListNetLoss"
        # ~82 MB per query at L = 1024 for the O(L^2) losses; metrics are O(L)
        per_q = 82e6 * (L / 1024.0) ** 2 if pair else 64.0 * L
        cpu_chunk = int(max(8, min(cfg["B"], 6e9 // per_q)))
        gpu_chunk = int(max(8, min(cfg["B"], 60e9 // per_q)))
        entry = {"workload": cfg["workload"], "L": L}
        entry["cpu"] = time_arm(cfg, torch.device("cpu"), cpu_chunk, args.budget_s)
        if torch.cuda.is_available() and not args.no_gpu:
            try:
                entry["gpu_eager"] = time_arm(cfg, torch.device("cuda", 0), gpu_chunk, args.budget_s / 3)
            except Exception as e:  # pragma: no cover  (e.g. out of memory on a smaller part)
                entry["gpu_eager"] = {"unavailable": repr(e)[:300]}
        doc["configs"][name] = entry
        sys.stderr.write(f"[time_reference] {name}: {json.dumps(entry)}\n")
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
