#!/usr/bin/env python
"""Benchmark of the loss hot path: `loss = fn(scores, relevance, n); loss.sum().backward()`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one forward+backward pass of the hot path over one padded batch of synthetic
queries (BASELINE.json metric "loss fwd+bwd queries/sec at (B, L)").  The default workload
is BASELINE.json configs[1]: LambdaNDCGLoss2, B=4096 queries x L=128 documents per GPU, fp32,
n ~ U[L/2, L], 5 relevance grades.  Successive steps use different batches of a resident
pool that is larger than the 126 MB L2, so no step finds its inputs in cache.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (the fused loss+gradient kernel) vs the measured HBM peak
  issue_roofline  the same kernel vs the FP32-issue / MUFU pair-rate ceiling, which is the
                binding limit of every O(L^2) loss
  cpu_baseline  the CPU oracle (a C port of the reference algorithm; the reference itself is
                Python and cannot travel to the GPU box) on this host's cores, bounded sample
  e2e           same metric through the public API with pinned HOST tensors: H2D of the inputs
                and D2H of loss + gradient inside the timed region
  clocks        SM clock / throttle reasons sampled during the timed region
`--impl reference` times the CPU oracle port with all host threads on the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024

# name -> (loss, B per GPU, L, relevance distribution)
CONFIGS = {
    "c2": dict(loss="LambdaNDCGLoss2", B=4096, L=128, workload="LambdaNDCGLoss2 synthetic (B=4096, L=128) fp32"),
    "c3": dict(loss="PairwiseDCGHingeLoss", B=1024, L=1024,
               workload="PairwiseDCGHingeLoss synthetic (B=1024, L=1024) fp32"),
    "c4": dict(loss="ListNetLoss", B=8192, L=200, skew=True,
               workload="ListNet synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
    # config 4 at eight times the batch: shows the ListNet kernel's bandwidth once the launch ramp
    # (a few microseconds) no longer dominates a 7 us kernel
    "c4x": dict(loss="ListNetLoss", B=65536, L=200, skew=True,
                workload="ListNet synthetic MSLR-WEB30K-shaped (B=65536, L=200) fp32"),
    "c5": dict(loss="LambdaNDCGLoss2", B=65536, L=512, strong=True,
               workload="LambdaNDCGLoss2 synthetic (B=65536, L=512) query-sharded"),
    "ns": dict(loss="LambdaNDCGLoss2", B=4096, L=1024,
               workload="LambdaNDCGLoss2 synthetic (B=4096, L=1024) fp32 (north-star point)"),
    # config 4 with its 136 features: linear scorer + ListNet fused into one pass over the features
    # (SURVEY.md 8(f) N1).  Single GPU per rank (weak scaling); 891 MB of features per batch.
    "c4f": dict(fused="linear_listnet", loss="ListNetLoss", B=8192, L=200, F=136, skew=True,
                workload="Linear(136,1) scorer + ListNet fused, MSLR-WEB30K-shaped (B=8192, L=200, F=136) fp32"),
    # forward-only ranking metrics (second half of config 4): 12 L + 12 algorithmic bytes / query
    "c4m": dict(metric="ndcg", k=10, B=8192, L=200, skew=True,
                workload="ndcg@10 synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
    "c4a": dict(metric="arp", k=None, B=8192, L=200, skew=True,
                workload="arp synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
    "c4d": dict(metric="dcg", k=None, B=8192, L=200, skew=True,
                workload="dcg at every rank synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
}
ORACLE_MODE = {"LambdaNDCGLoss2": ("lambda", "ndcg2"), "LambdaNDCGLoss1": ("lambda", "ndcg1"),
               "LambdaARPLoss1": ("lambda", "arp1"), "LambdaARPLoss2": ("lambda", "arp2"),
               "PairwiseHingeLoss": ("additive", "hinge"), "PairwiseDCGHingeLoss": ("additive", "dcg_hinge"),
               "PairwiseLogisticLoss": ("additive", "logistic"), "ListNetLoss": ("listnet", None)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_batch_numpy(seed, B, L, skew=False):
    """Synthetic padded batch: scores randn, n ~ U[L/2, L], relevance uniform
    over 5 grades (or MSLR-like skew), padded relevance zeroed, padded scores left random."""
    import numpy as np
    rng = np.random.default_rng(seed)
    scores = rng.standard_normal((B, L), dtype=np.float32)
    n = rng.integers(L // 2, L + 1, size=B, dtype=np.int64)
    if skew:
        rel = rng.choice(5, size=(B, L), p=[.52, .32, .13, .02, .01]).astype(np.int64)
    else:
        rel = rng.integers(0, 5, size=(B, L), dtype=np.int64)
    rel[np.arange(L)[None, :] >= n[:, None]] = 0
    return scores, rel, n


def metric_name(cfg):
    return "loss fwd+bwd queries/sec" if "loss" in cfg else "ranking metric queries/sec"


def valid_pairs(n):
    import numpy as np
    n = np.asarray(n, dtype=np.float64)
    return float((n * (n - 1) / 2).sum())


def oracle_step(oracle, family, mode, cfg, s, y, n):
    """One pass of the CPU oracle port over a batch (loss + gradient, or metric)."""
    if family == "lambda":
        return oracle.lambda_loss(mode, s, y, n)
    if family == "additive":
        return oracle.pairwise_additive(mode, s, y, n, f32="hinge" in mode)
    if family == "listnet":
        return oracle.listnet(s, y, n)
    if mode == "arp":
        return oracle.arp(s, y, n), None
    return oracle.dcg(s, y, n, k=cfg.get("k"), normalized=mode == "ndcg"), None


# --------------------------------------------------------------------------- reference arm
def run_reference(args, cfg, rank, world):
    """CPU oracle port of the reference algorithm, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    oracle.build()
    oracle.set_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1: use every host core
    threads = oracle.max_threads()
    family, mode = ORACLE_MODE[cfg["loss"]] if "loss" in cfg else ("metric", cfg["metric"])
    L = cfg["L"]
    # bounded sample: a slice of the workload's batch sized for ~1 s per step
    probe_B = 64
    s, y, n = make_batch_numpy(1234, probe_B, L, cfg.get("skew", False))
    fused_F = cfg.get("F") if cfg.get("fused") else None
    rng = np.random.default_rng(7)
    fused_w = rng.standard_normal(fused_F or 1).astype(np.float32) * 0.1

    def step(s, y, n):
        if fused_F:
            # scorer + ListNet + weight gradient (numpy float64 port; features drawn once per size)
            key = len(n)
            if key not in step.feat:
                step.feat[key] = rng.standard_normal((key, L, fused_F), dtype=np.float32)
            return oracle.linear_listnet(step.feat[key], fused_w, None, y, n)
        return oracle_step(oracle, family, mode, cfg, s, y, n)

    step.feat = {}

    step(s, y, n)
    t0 = time.perf_counter()
    step(s, y, n)
    per_q = (time.perf_counter() - t0) / probe_B
    sample_B = int(max(threads * 8, min(cfg["B"], 1.0 / max(per_q, 1e-9))))
    s, y, n = make_batch_numpy(1235, sample_B, L, cfg.get("skew", False))
    for _ in range(min(args.warmup, 3)):
        step(s, y, n)
    steps = max(1, min(args.steps, 30))
    t0 = time.perf_counter()
    for _ in range(steps):
        step(s, y, n)
    dt = time.perf_counter() - t0
    qps = sample_B * steps / dt
    sample = (f"{sample_B} of the {cfg['B']} queries of one batch per step, {steps} steps; " +
              ("oracle.linear_listnet (numpy float64 port of Linear + ListNet + weight gradient, BLAS threads)"
               if fused_F else "oracle/ltr_oracle.c (C port of the reference algorithm, OpenMP over queries)"))
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "loss": cfg.get("loss", cfg.get("metric")),
                   "B_per_step": sample_B, "L": L},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- fused scorer + loss
def run_fused(args, cfg, rank, local_rank, world):
    """Config c4f: `loss = LinearListNet(F)(xs, ys, n); loss.mean().backward()` -- scores, ListNet,
    d loss / d scores and the weight / bias gradients in one pass over the (B, L, F) features."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from pytorchltr_b200 import _lib
    from pytorchltr_b200.fused import LinearListNet
    from pytorchltr_b200.loss import ListNetLoss

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L, F = cfg["B"], cfg["L"], cfg["F"]
    _, y_np, n_np = make_batch_numpy(1234 + 1000 * rank, B, L, True)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    xs = torch.randn(B, L, F, device=dev, generator=gen)            # 891 MB: larger than L2 by itself
    ys, ns = torch.from_numpy(y_np).to(dev), torch.from_numpy(n_np).to(dev)
    torch.manual_seed(5)
    model = LinearListNet(F).to(dev)
    plain = torch.nn.Linear(F, 1).to(dev)
    plain.load_state_dict(model.linear.state_dict())
    plain_loss = ListNetLoss()
    red = torch.zeros(2, device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        out = model(xs, ys, ns)
        out.mean().backward()
        return out

    def unfused_step():
        plain.zero_grad(set_to_none=True)
        out = plain_loss(plain(xs), ys, ns)
        out.mean().backward()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # CUDA graph of the whole step (forward + backward): the eager step is launch-bound
    launch = "eager"
    run = step
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            model.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            run = graph.replay
            launch = "cuda_graph"
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); falling back to eager\n")
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(run, args.steps, max(args.warmup, 3))
    sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)
    unfused_ms = timed(unfused_step, min(args.steps, 20), 3) / min(args.steps, 20)

    # kernel only: ltr_linear_listnet (fused kernel + the partial-gradient reduction)
    lib = _lib.lib()
    w = model.linear.weight.detach().reshape(-1).contiguous()
    bias = model.linear.bias.detach().contiguous()
    loss_buf, dsc = torch.empty(B, device=dev), torch.empty(B, L, device=dev)
    qg = torch.empty(B, F + 1, device=dev)
    dw, db = torch.empty(F, device=dev), torch.empty(1, device=dev)
    ws = torch.empty(lib.ltr_linear_listnet_workspace_bytes(F), dtype=torch.uint8, device=dev)
    gmean = torch.full((B,), 1.0 / B, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def kernel_only(nq=B):
        _lib.check(lib.ltr_linear_listnet(xs.data_ptr(), w.data_ptr(), bias.data_ptr(), ys.data_ptr(), 8,
                                          ns.data_ptr(), 8, nq, L, F, None, loss_buf.data_ptr(), dsc.data_ptr(),
                                          qg.data_ptr(), None, st))

    def backward_only(nq=B, g=None, stride=1):
        g = gmean if g is None else g
        _lib.check(lib.ltr_linear_listnet_backward(qg.data_ptr(), g.data_ptr(), stride, nq, F, dw.data_ptr(),
                                                   db.data_ptr(), ws.data_ptr(), ws.numel(), st))

    k_reps = 20
    kernel_ms = timed(kernel_only, k_reps, 3) / k_reps
    backward_ms = timed(backward_only, k_reps, 3) / k_reps
    alg_bytes = B * (4 * L * F + 12 * L + 12 + 4 * (F + 1))   # + the per-query gradient row
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6548.2, "fallback: B200_PROFILING.md measured copy bandwidth"
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = (json.load(open(tpath)).get("c4f") or {}).get("dram_bytes_per_launch")

    # parity spot check against the float64 oracle (not timed)
    parity = None
    cpu_baseline = None
    e2e = None
    if rank == 0:
        import oracle
        idx = np.arange(0, B, B // 16)
        out = step()
        torch.cuda.synchronize()
        xn = xs[idx].cpu().numpy()
        _, rl, _, _, _, _ = oracle.linear_listnet(xn, w.cpu().numpy(), bias.cpu().numpy(), y_np[idx], n_np[idx])
        got = out.detach().cpu().double().numpy()[idx]
        err = float(np.abs(got - rl).max() / max(1e-30, np.abs(rl).max()))
        _, _, _, dw_ref, _, gscale = oracle.linear_listnet(xs[:256].cpu().numpy(), w.cpu().numpy(),
                                                           bias.cpu().numpy(), y_np[:256], n_np[:256])
        # weight gradient of the first 256 queries through the C ABI (upstream gradient of ones)
        kernel_only(256)
        backward_only(256, torch.ones(1, device=dev), 0)
        torch.cuda.synchronize()
        gerr = float((np.abs(dw.cpu().double().numpy() - dw_ref) / gscale).max())
        parity = {"max_loss_err_rel": err, "max_dweight_err_rel_to_abs_sum": gerr, "queries_checked": 16 + 256,
                  "ok": bool(err <= 1e-5 and gerr <= 1e-5)}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample_B = 512
            xn = xs[:sample_B].cpu().numpy()
            wn, bn = w.cpu().numpy(), bias.cpu().numpy()
            oracle.linear_listnet(xn[:32], wn, bn, y_np[:32], n_np[:32])
            t0 = time.perf_counter()
            oracle.linear_listnet(xn, wn, bn, y_np[:sample_B], n_np[:sample_B])
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": sample_B / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                            "sample": f"{sample_B} queries of the same workload in one call ({dt:.2f} s wall), "
                                      "oracle.linear_listnet (numpy float64, BLAS threads)"}
    if not args.no_e2e:
        # end to end: features, relevance and n come from pinned HOST memory every step
        hB = B // 8
        hx = torch.empty(hB, L, F, pin_memory=True).normal_()
        hy, hn = torch.from_numpy(y_np[:hB]).pin_memory(), torch.from_numpy(n_np[:hB]).pin_memory()

        def e2e_step():
            model.zero_grad(set_to_none=True)
            out = model(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True), hn.to(dev, non_blocking=True))
            out.mean().backward()
            return out.cpu(), model.linear.weight.grad.cpu()

        for _ in range(2):
            e2e_step()
        barrier()
        e_steps = 5
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": hB * world * e_steps / dt, "unit": "queries/s",
               "h2d_bytes_per_step": hB * (4 * L * F + 8 * L + 8), "d2h_bytes_per_step": hB * 4 + 4 * F,
               "steps": e_steps, "ms_per_step": dt / e_steps * 1e3,
               "path": f"LinearListNet on {hB} queries per step from pinned host memory: H2D of features / "
                       "relevance / n, fused kernel, D2H of the loss and the weight gradient"}
    if rank == 0:
        line = {
            "metric": "loss fwd+bwd queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "loss": "LinearListNet", "B_per_gpu": B, "L": L, "F": F,
                       "global_batch": B * world, "parallelism": f"query-sharded dp{world}", "launch": launch,
                       "l2_policy": "inputs larger than L2: one 891 MB feature batch per step",
                       "n_distribution": "n ~ U[L/2, L]",
                       "unfused_ms_per_step": unfused_ms,
                       "unfused_path": "ListNetLoss()(torch.nn.Linear(F, 1)(xs), ys, n).mean().backward(): "
                                       "two passes over the features (cuBLAS) + ltr_listnet"},
            "roofline": {"bound": "hbm", "kernel": "linear_listnet_kernel (ltr_linear_listnet)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel_ms": kernel_ms, "backward_ms": backward_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "issue_roofline": None, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "clocks": sampler.summary("timed region"),
            "gpu_launches": 3 * args.steps,
            "gpu_launches_note": "per step: fused kernel (forward) + weighted column sum and its reduction "
                                 "(backward) of libltr_sm100.so (plus torch's mean kernels)",
            "parity_spot_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every ~20 ms (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, torch_device_index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()

    def summary(self, window):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "window": "unavailable: " + getattr(self, "err", "no samples")}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "window": window}


# --------------------------------------------------------------------------- our arm
def run_ours(args, cfg, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import pytorchltr_b200.evaluation as ltr_eval
    import pytorchltr_b200.loss as ltr_loss
    from pytorchltr_b200 import _lib, _ops
    from pytorchltr_b200.distributed import shard_bounds

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = cfg["L"]
    if cfg.get("strong"):
        lo, hi = shard_bounds(cfg["B"], rank, world)
        B = hi - lo
        scaling = "strong"
    else:
        B = cfg["B"]
        scaling = "weak"
    is_metric = "metric" in cfg
    if is_metric:
        family, mode = "metric", cfg["metric"]
        if mode == "arp":
            loss_fn = ltr_eval.arp
        else:
            _fn, _k = getattr(ltr_eval, mode), cfg.get("k")
            loss_fn = lambda s, y, n: _fn(s, y, n, k=_k)  # noqa: E731
    else:
        loss_fn = getattr(ltr_loss, cfg["loss"])()
        family, mode = ORACLE_MODE[cfg["loss"]]

    # ---- resident pool of distinct batches, larger than L2 ---------------------------------
    in_bytes = B * L * 12 + B * 8
    pool_n = max(4, -(-2 * L2_BYTES // in_bytes))
    pool_n = min(pool_n, 64)
    pool, pairs_per_batch = [], []
    for i in range(pool_n):
        s, y, n = make_batch_numpy(1234 + 1000 * rank + i, B, L, cfg.get("skew", False))
        pairs_per_batch.append(valid_pairs(n))
        pool.append((torch.from_numpy(s).to(dev).requires_grad_(not is_metric), torch.from_numpy(y).to(dev),
                     torch.from_numpy(n).to(dev)))
    pool_bytes = pool_n * in_bytes
    torch.cuda.synchronize()

    # N > 1: the path's only exchange is the 2-element [sum loss, #queries] all-reduce behind the
    # global mean.  Its input is produced inside the (graph-captured) step; the collective itself
    # is enqueued after the step, asynchronously, so it overlaps the next step's kernels.
    red = [torch.tensor([0.0, float(B)], device=dev) for _ in range(pool_n)] if world > 1 else None
    pending = []

    metric_out = [None] * pool_n

    def step(i):
        s, y, n = pool[i % pool_n]
        if is_metric:
            out = loss_fn(s, y, n)
            metric_out[i % pool_n] = out
        else:
            s.grad = None
            out = loss_fn(s, y, n)
        if world > 1:
            r = red[i % pool_n]
            torch.sum(out.detach(), dim=0, keepdim=True, out=r[:1])   # r[1] (the count) is constant
        if not is_metric:
            out.sum().backward()
        return out

    def exchange(i):
        if world > 1:
            pending.append(dist.all_reduce(red[i % pool_n], async_op=True))
            if len(pending) > 8:
                pending.pop(0).wait()

    # ---- CUDA graphs: one per pool entry (launch-bound otherwise: ~25 us of GPU work/step) ---
    graphs = None
    launch = "eager"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs = []
            for i in range(pool_n):
                g = torch.cuda.CUDAGraph()
                pool[i][0].grad = None
                with torch.cuda.graph(g):
                    step(i)
                graphs.append(g)
            launch = "cuda_graph"
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); falling back to eager\n")
            graphs = None
            torch.cuda.synchronize()

    def run_step(i, collective=True):
        if graphs is not None:
            graphs[i % pool_n].replay()
        else:
            step(i)
        if collective:
            exchange(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        run_step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        run_step(args.warmup + i)
    while pending:
        pending.pop(0).wait()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clock_window = "timed region"
    if len(sampler.samples) < 5:
        # the timed region was shorter than a few NVML periods: keep the same load running
        t_end = time.perf_counter() + 0.5
        i = 0
        while time.perf_counter() < t_end:
            run_step(i, collective=False)   # time-bounded: ranks may differ in step count
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        while pending:
            pending.pop(0).wait()
        torch.cuda.synchronize()
        clock_window = "timed region + 0.5 s of the same load"
    sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_queries = B * world if not cfg.get("strong") else cfg["B"]
    value = total_queries * args.steps / (ms * 1e-3)

    # parity spot check of the timed path (not timed): last batch vs the oracle on a sample
    parity = None
    if rank == 0:
        import oracle
        s, y, n = pool[0]
        idx = np.arange(0, B, max(1, B // 16))
        run_step(0, collective=False)   # rank 0 only: must not enqueue an unmatched collective
        torch.cuda.synchronize()
        sn, yn, nn = s.detach().cpu().numpy()[idx], y.cpu().numpy()[idx], n.cpu().numpy()[idx]
        # (hinge: float32 restatement -- pairs on the kink flip between f32 and f64 rounding)
        rl, rg = oracle_step(oracle, family, mode, cfg, sn, yn, nn)
        if is_metric:
            got = metric_out[0].detach().cpu().double().numpy()[idx]
            err = float(np.abs(got - rl).max())
            parity = {"max_abs_err": err, "queries_checked": int(len(idx)), "ok": err <= 1e-5}
        else:
            got = s.grad.detach().cpu().double().numpy()[idx]
            gerr = float((np.abs(got - rg) / (np.abs(rg).max(axis=1, keepdims=True) + 1e-30)).max())
            parity = {"max_grad_err_rel_to_rowmax": gerr, "queries_checked": int(len(idx)), "ok": gerr <= 1e-5}

    # ---- dominant kernel alone: the fused loss + gradient kernel -----------------------------
    lib_family = {"lambda": _lib.FAMILY_LAMBDA, "additive": _lib.FAMILY_ADDITIVE,
                  "listnet": _lib.FAMILY_LISTNET, "metric": -1}[family]
    lib_mode = {"ndcg2": _lib.LAM_NDCG2, "ndcg1": _lib.LAM_NDCG1, "arp1": _lib.LAM_ARP1,
                "arp2": _lib.LAM_ARP2, "hinge": _lib.ADD_HINGE, "dcg_hinge": _lib.ADD_DCG_HINGE,
                "logistic": _lib.ADD_LOGISTIC, None: 0, "dcg": _lib.METRIC_DCG,
                "ndcg": _lib.METRIC_NDCG, "arp": _lib.METRIC_ARP}[mode]
    metric_k = 1 if mode == "arp" else (cfg.get("k") or 0) if is_metric else 0
    metric_ld = L if (is_metric and mode != "arp" and not cfg.get("k")) else 1
    metric_buf = torch.empty(B * metric_ld, device=dev) if is_metric else None
    lib = _lib.lib()
    loss_buf = torch.empty(B, device=dev)
    grad_buf = torch.empty(B, L, device=dev)
    sched_ws = torch.empty(lib.ltr_schedule_workspace_bytes(B), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def kernel_only(i):
        nonlocal st
        s, y, n = pool[i % pool_n]
        # the _ws entry points are what the public modules call: the timed launch includes the
        # small query-ordering kernel that precedes the fused kernel
        if lib_family == _lib.FAMILY_LAMBDA:
            rc = lib.ltr_lambda_ws(lib_mode, s.data_ptr(), y.data_ptr(), 8, n.data_ptr(), 8, B, L, 1.0,
                                   loss_buf.data_ptr(), grad_buf.data_ptr(), None, None,
                                   sched_ws.data_ptr(), sched_ws.numel(), st)
        elif lib_family == _lib.FAMILY_ADDITIVE:
            rc = lib.ltr_pairwise_additive_ws(lib_mode, s.data_ptr(), y.data_ptr(), 8, n.data_ptr(), 8, B, L,
                                              1.0, loss_buf.data_ptr(), grad_buf.data_ptr(), None,
                                              sched_ws.data_ptr(), sched_ws.numel(), st)
        elif lib_family == _lib.FAMILY_LISTNET:
            rc = lib.ltr_listnet(s.data_ptr(), y.data_ptr(), 8, n.data_ptr(), 8, B, L,
                                 loss_buf.data_ptr(), grad_buf.data_ptr(), None, st)
        else:
            rc = lib.ltr_rank_metrics(lib_mode, s.data_ptr(), y.data_ptr(), 8, n.data_ptr(), 8, B, L,
                                      min(metric_k, L), 1, metric_buf.data_ptr(), metric_ld, st)
        _lib.check(rc)

    for i in range(8):
        kernel_only(i)
    torch.cuda.synchronize()
    reps = max(1, min(20, 2000 // pool_n))
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kgraph = None
    if not args.no_graph:
        # one graph holding a launch per pool entry: device-side back-to-back, no host launch gaps
        try:
            kgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(kgraph):
                st = torch.cuda.current_stream().cuda_stream
                for i in range(pool_n):
                    kernel_only(i)
            st = torch.cuda.current_stream().cuda_stream
            kgraph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] kernel-only graph capture failed ({e!r}); timing eager launches\n")
            kgraph = None
            st = torch.cuda.current_stream().cuda_stream
    k0.record()
    for r in range(reps):
        if kgraph is not None:
            kgraph.replay()
        else:
            for i in range(pool_n):
                kernel_only(i)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / (reps * pool_n)
    alg_bytes = B * (12 * L + 8 + 4 * metric_ld) if is_metric else B * (16 * L + 16)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.config, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "ltr_rank_metrics kernel" if is_metric else
                "fused loss+gradient kernel (ltr_lambda / ltr_pairwise_additive / ltr_listnet)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src}
    issue = None
    if family not in ("listnet", "metric"):
        mean_pairs = sum(pairs_per_batch) / len(pairs_per_batch)
        sm_hz = (sampler.max_mhz or 1965) * 1e6
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # sigmoid losses: the tile kernels spend 2 MUFU (rcp, lg2) per unordered pair at 16 MUFU
        # lanes / clk / SM (a naive evaluation needs 3: ex2, rcp, lg2); hinge: FP32 issue, ~6
        # lane-ops per pair at 128 lanes / clk / SM
        per_pair_clk = (6.0 / 128.0) if "Hinge" in cfg["loss"] else (2.0 / 16.0)
        peak_pairs = sms * sm_hz / per_pair_clk
        ach_pairs = mean_pairs / (kernel_ms * 1e-3)
        issue = {"bound": "mufu" if "Hinge" not in cfg["loss"] else "fp32_issue",
                 "achieved": ach_pairs, "peak": peak_pairs, "unit": "unordered pairs/s",
                 "frac": ach_pairs / peak_pairs, "pairs_per_launch": mean_pairs,
                 "note": "binding roofline of the O(L^2) losses (SURVEY.md F4); peak derived from "
                         "SM count x max SM clock x pipe width"}
        if "Hinge" in cfg["loss"] and 128 < L <= 4096 and os.environ.get("LTR_HINGE") != "pairs" \
                and os.environ.get("LTR_KERNEL") not in ("tiles", "generic"):
            # the hinge losses no longer enumerate pairs at these sizes (sort + scans, O(n log n)):
            # "achieved" is the pair rate an O(L^2) kernel would need to match it, not an issue rate
            issue.update({"bound": "latency (O(n log n) sort + scans; pairs are not enumerated)",
                          "frac": None, "equivalent_pair_rate_vs_fp32_issue_peak": ach_pairs / peak_pairs})

    # ---- e2e: public API, pinned host tensors in, host loss + gradient out -------------------
    e2e = None
    if not args.no_e2e:
        host_n = 4
        host_pool = []
        for i in range(host_n):
            s, y, n = make_batch_numpy(99 + 1000 * rank + i, B, L, cfg.get("skew", False))
            host_pool.append((torch.from_numpy(s).pin_memory().requires_grad_(not is_metric),
                              torch.from_numpy(y).pin_memory(), torch.from_numpy(n).pin_memory()))

        def e2e_step(i):
            s, y, n = host_pool[i % host_n]
            if is_metric:
                return loss_fn(s, y, n)     # H2D scores/relevance/n, kernel, D2H metric
            s.grad = None
            out = loss_fn(s, y, n)          # H2D scores/relevance/n, kernel, D2H loss
            out.sum().backward()            # H2D g, scale kernel, D2H gradient
            return out

        e_steps = max(5, min(args.steps, 100))
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_queries * e_steps / dt, "unit": "queries/s",
               "h2d_bytes_per_step": B * L * 12 + B * 8 + (0 if is_metric else B * 4),
               "d2h_bytes_per_step": B * 4 * metric_ld if is_metric else B * 4 + B * L * 4,
               "steps": e_steps, "ms_per_step": dt / e_steps * 1e3,
               "path": "loss_fn(pinned CPU tensors).sum().backward(): results returned as CPU tensors"}

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.set_threads(os.cpu_count() or 1)
        threads = oracle.max_threads()
        s, y, n = make_batch_numpy(1234, min(B, 64), L, cfg.get("skew", False))

        def ostep(s, y, n):
            oracle_step(oracle, family, mode, cfg, s, y, n)

        ostep(s, y, n)
        t0 = time.perf_counter()
        ostep(s, y, n)
        per_q = (time.perf_counter() - t0) / len(n)
        sample_B = int(max(threads * 8, min(4 * B, 5.0 / max(per_q, 1e-9))))
        s, y, n = make_batch_numpy(1235, sample_B, L, cfg.get("skew", False))
        t0 = time.perf_counter()
        ostep(s, y, n)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sample_B / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                        "sample": f"{sample_B} queries of the same synthetic workload in one call "
                                  f"({dt:.1f} s wall), oracle/ltr_oracle.c with OpenMP over queries"}

    if rank == 0:
        # fused kernel (+ backward row-scale kernel; + the query-ordering kernel of the O(L^2) losses)
        launches_per_step = 1 if is_metric else (2 if family == "listnet" else 3)
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "loss": cfg.get("loss", cfg.get("metric")), "B_per_gpu": B, "L": L,
                       "global_batch": total_queries, "parallelism": f"query-sharded dp{world}",
                       "launch": launch,
                       "l2_policy": f"inputs larger than L2: {pool_n} distinct resident batches "
                                    f"({pool_bytes / 2**20:.0f} MiB) visited round-robin",
                       "n_distribution": "n ~ U[L/2, L]", "sigma": 1.0},
            "roofline": roofline, "issue_roofline": issue, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "clocks": sampler.summary(clock_window),
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_note": "per step: 1 fused kernel (+ 1 row-scale kernel in the backward pass, + 1 "
                                 "query-ordering kernel before the O(L^2) losses when queries queue) of "
                                 "libltr_sm100.so (plus torch's sum / ones_like fill)",
            "parity_spot_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if world != args.gpus and rank == 0:
        sys.stderr.write(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun "
                         f"for N > 1; running {world} rank(s)\n")
    if cfg.get("fused"):
        run_fused(args, cfg, rank, local_rank, world)
        return
    run_ours(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
