#!/usr/bin/env python
"""Benchmark of the loss hot path: `loss_fn(scores, relevance, n).mean().backward()`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one forward+backward pass of the hot path over one padded batch of synthetic
queries (BASELINE.json metric "loss fwd+bwd queries/sec at (B, L)"), written the way the
reference's own training loop writes it (examples/01-basic-usage.py:72-74).

Default workload = BASELINE.json configs[4], the largest single-GPU configuration and the split
`north_star` names for 1/2/4/8 GPUs: LambdaNDCGLoss2, B=65536 queries x L=512 documents, fp32,
n ~ U[L/2, L], 5 relevance grades, STRONG scaling: the 65536 queries are sharded over the ranks
(contiguous blocks, no data-path collective) and the step of a multi-rank run is
`pytorchltr_b200.distributed.sharded_mean_loss(...).backward()`, whose 2-element [sum, count]
all-reduce (fed by the loss kernel's own epilogue) is captured into the step's CUDA graph.
Successive steps use different batches of a resident pool that is larger than the 126 MB L2.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline        dominant kernel (the fused loss+gradient kernel) vs the measured HBM peak
  issue_roofline  the same kernel vs the MEASURED MUFU pair-rate ceiling (profiles/issue_peaks.json,
                  tools/issue_peak.cu), the binding limit of every O(L^2) loss
  windows         best / median ms per step over repeated K-step windows of the same run
  cpu_baseline    the CPU oracle (C port of the reference algorithm, OpenMP) on this host's cores,
                  bounded sample; plus the unmodified Python reference on CPU and eager on the B200
                  (live when git-ignored baseline/_ref is present, else profiles/reference_timing.json)
  e2e             same metric through the public API with pinned HOST tensors (reference dtypes:
                  float32 scores, int64 relevance, int64 n): H2D of the inputs and D2H of loss +
                  gradient inside the timed region; e2e_compact = same with uint8 relevance /
                  int32 n (an extension of this package)
  configs         (N = 1 only) the other BASELINE.json configurations measured in the same run:
                  ns (4096, 1024) north-star point, c2, c3, c4, c4m, c4a, c4d -- ms_per_step,
                  kernel_ms, HBM frac, issue frac, parity spot check
  clocks          SM clock / throttle reasons sampled during the timed region
`--impl reference` times the reference's CPU implementation of the same workload on the host cores:
the unmodified Python reference from baseline/_ref when present ("kind": "reference"), else the C
oracle port ("kind": "port"); each step is a bounded sample of the workload's batch.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# name -> workload
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
    # config 4 with the reference's documented model (docs/source/getting-started.rst:42-51, 136 -> 50 -> 10 -> 1
    # ReLU MLP) and its documented loss (:73-75): scorer forward and backward on tcgen05 (ltr_mlp_scores /
    # ltr_mlp_backward), the loss kernel between them
    "c4mlp": dict(fused="mlp", loss="PairwiseHingeLoss", B=8192, L=200, F=136, skew=True,
                  workload="MLP(136-50-10-1) ranker + PairwiseHingeLoss, MSLR-WEB30K-shaped (B=8192, L=200, F=136) "
                           "tf32 layer 1 / fp32"),
    # forward-only ranking metrics (second half of config 4): 12 L + 12 algorithmic bytes / query
    "c4m": dict(metric="ndcg", k=10, B=8192, L=200, skew=True,
                workload="ndcg@10 synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
    "c4a": dict(metric="arp", k=None, B=8192, L=200, skew=True,
                workload="arp synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
    "c4d": dict(metric="dcg", k=None, B=8192, L=200, skew=True,
                workload="dcg at every rank synthetic MSLR-WEB30K-shaped (B=8192, L=200) fp32"),
}
DEFAULT_SUB = "ns,c2,c3,c4,c4m,c4a,c4d,c4mlp"
ORACLE_MODE = {"LambdaNDCGLoss2": ("lambda", "ndcg2"), "LambdaNDCGLoss1": ("lambda", "ndcg1"),
               "LambdaARPLoss1": ("lambda", "arp1"), "LambdaARPLoss2": ("lambda", "arp2"),
               "PairwiseHingeLoss": ("additive", "hinge"), "PairwiseDCGHingeLoss": ("additive", "dcg_hinge"),
               "PairwiseLogisticLoss": ("additive", "logistic"), "ListNetLoss": ("listnet", None)}
# MUFU instructions the shipped kernels spend per unordered pair of a sigmoid-weighted loss
# (lg2 of every pair + one rcp per TWO pairs)
MUFU_PER_PAIR = 1.5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--sub", default=DEFAULT_SUB,
                    help="comma list of other configs measured in the same run at N = 1 ('' = none)")
    ap.add_argument("--no-sub", action="store_true")
    ap.add_argument("--windows", type=int, default=5, help="extra K-step windows for best / median")
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


def config_block(cfg, world):
    """Names the workload; identical in both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": cfg["workload"], "loss": cfg.get("loss", cfg.get("metric")), "B": cfg["B"], "L": cfg["L"],
            "n_distribution": "n ~ U[L/2, L]", "sigma": 1.0,
            "parallelism": f"query-sharded dp{world}"}


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


def family_mode(cfg):
    return ORACLE_MODE[cfg["loss"]] if "loss" in cfg else ("metric", cfg["metric"])


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def mufu_peak():
    """Measured MUFU lane-operations per second of one B200 (tools/issue_peak.cu)."""
    path = os.path.join(ROOT, "profiles", "issue_peaks.json")
    if os.path.exists(path):
        try:
            st = json.load(open(path))["streams"]
            rate = max(float(st[k]["per_s"]) for k in ("mufu_lg2", "mufu_ex2", "mufu_rcp_lg2") if k in st)
            return rate, "profiles/issue_peaks.json (tools/issue_peak.cu: measured MUFU lane-ops/s)"
        except Exception:  # pragma: no cover
            pass
    return 148 * 1.965e9 * 16, "derived: 148 SMs x 1965 MHz x 16 MUFU lanes/clk (no measurement found)"


# --------------------------------------------------------------------------- the real reference
def import_reference():
    """The unmodified reference installed into git-ignored baseline/_ref (pip --target), or None."""
    if not os.path.isdir(os.path.join(REF_DIR, "pytorchltr")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import pytorchltr  # noqa: F401
        import pytorchltr.evaluation
        import pytorchltr.loss
        return pytorchltr
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"[bench] baseline/_ref present but not importable: {e!r}\n")
        return None


def reference_callable(ref, cfg):
    """(callable(scores, relevance, n), is_loss) of the reference for this workload, or (None, _)."""
    if "metric" in cfg:
        if cfg["metric"] == "arp":
            return (lambda s, y, n: ref.evaluation.arp(s, y, n)), False
        fn, k = getattr(ref.evaluation, cfg["metric"]), cfg.get("k")
        return (lambda s, y, n: fn(s, y, n, k=k)), False
    if not hasattr(ref.loss, cfg["loss"]):
        return None, True                     # ListNet: absent from the reference
    return getattr(ref.loss, cfg["loss"])(), True


def time_python_reference(ref, cfg, device, chunk, n_chunks, warm=1):
    """Seconds per chunk (median) of `fn(s, y, n).mean().backward()` of the unmodified reference."""
    import torch
    fn, is_loss = reference_callable(ref, cfg)
    if fn is None:
        return None
    s_np, y_np, n_np = make_batch_numpy(1235, chunk, cfg["L"], cfg.get("skew", False))
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

    for _ in range(warm):
        step()
    times = []
    for _ in range(n_chunks):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times


def committed_reference_timing(name):
    path = os.path.join(ROOT, "profiles", "reference_timing.json")
    if not os.path.exists(path):
        return None
    try:
        doc = json.load(open(path))
        e = doc["configs"].get(name)
        if e is None:
            return None
        return {"source": "profiles/reference_timing.json (tools/time_reference.py on a gpurun B200 box; "
                          "baseline/_ref absent in this run)",
                "cpu_cores": doc.get("cpu_cores"),
                "cpu_queries_per_s": (e.get("cpu") or {}).get("queries_per_s"),
                "cpu_chunk": (e.get("cpu") or {}).get("chunk"),
                "gpu_eager_queries_per_s": (e.get("gpu_eager") or {}).get("queries_per_s"),
                "gpu_eager_chunk": (e.get("gpu_eager") or {}).get("chunk")}
    except Exception:  # pragma: no cover
        return None


def pair_chunk(cfg, budget_bytes):
    """Queries per call of the reference that keep its ~78 L^2 B/query of pair tensors in budget."""
    L = cfg["L"]
    pair = "metric" not in cfg and cfg["loss"] != "ListNetLoss"
    per_q = 82e6 * (L / 1024.0) ** 2 if pair else 64.0 * L
    return int(max(8, min(cfg["B"], budget_bytes // per_q)))


# --------------------------------------------------------------------------- reference arm
def run_reference(args, name, cfg, rank, world):
    """The reference's own CPU implementation of the workload on the host cores, all threads:
    the unmodified Python reference when baseline/_ref is present, else the C oracle port.
    EXACTLY --steps timed steps after --warmup; a step is a bounded sample of the batch."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    oracle.set_threads(cores)   # torchrun exports OMP_NUM_THREADS=1: use every host core
    family, mode = family_mode(cfg)
    L = cfg["L"]
    fused_F = cfg.get("F") if cfg.get("fused") else None
    total_budget_s = 90.0
    n_steps = args.steps + args.warmup

    # ---- the C port (always measured: continuity with round 1, and the fallback arm) --------
    rng = np.random.default_rng(7)
    fused_w = rng.standard_normal(fused_F or 1).astype(np.float32) * 0.1
    feat_cache = {}

    mlp_p = [rng.standard_normal(sh).astype(np.float32) * 0.1
             for sh in ((50, fused_F or 1), (50,), (10, 50), (10,), (1, 10), (1,))]

    def port_step(s, y, n):
        if fused_F:
            key = len(n)
            if key not in feat_cache:
                feat_cache[key] = rng.standard_normal((key, L, fused_F), dtype=np.float32)
            if cfg.get("fused") == "mlp":      # numpy restatement of the model around the C port of the loss
                x2 = feat_cache[key].reshape(-1, fused_F)
                sc = oracle.mlp_scores(x2, *mlp_p).reshape(key, L).astype(np.float32)
                out = oracle_step(oracle, family, mode, cfg, sc, y, n)
                oracle.mlp_grads(x2, *mlp_p, np.asarray(out[1]).reshape(-1) / key)
                return out
            return oracle.linear_listnet(feat_cache[key], fused_w, None, y, n)
        return oracle_step(oracle, family, mode, cfg, s, y, n)

    s, y, n = make_batch_numpy(1234, 64, L, cfg.get("skew", False))
    port_step(s, y, n)
    t0 = time.perf_counter()
    port_step(s, y, n)
    port_per_q = (time.perf_counter() - t0) / 64

    ref = None if fused_F else import_reference()
    fn = reference_callable(ref, cfg)[0] if ref is not None else None
    if fn is not None:
        import torch
        torch.set_num_threads(cores)
        kind = "reference"
        probe = pair_chunk(cfg, 1e9)
        t = time_python_reference(ref, cfg, torch.device("cpu"), probe, 1, warm=1)
        per_q = t[0] / probe
        sample_B = int(max(8, min(cfg["B"], pair_chunk(cfg, 8e9), (total_budget_s / n_steps) / max(per_q, 1e-9))))
        s_np, y_np, n_np = make_batch_numpy(1235, sample_B, L, cfg.get("skew", False))
        ts = torch.from_numpy(s_np).requires_grad_("loss" in cfg)
        ty, tn = torch.from_numpy(y_np), torch.from_numpy(n_np)
        is_loss = "loss" in cfg

        def step():
            if is_loss:
                ts.grad = None
                fn(ts, ty, tn).mean().backward()
            else:
                fn(ts, ty, tn)

        threads = torch.get_num_threads()
        what = (f"unmodified rjagerman/pytorchltr from baseline/_ref on CPU, torch {torch.__version__}, "
                f"{threads} threads: loss_fn(scores, relevance, n).mean().backward()")
        dtype = "f32"
    else:
        kind = "port"
        sample_B = int(max(cores * 8, min(cfg["B"], (total_budget_s / n_steps) / max(port_per_q, 1e-9))))
        s, y, n = make_batch_numpy(1235, sample_B, L, cfg.get("skew", False))

        def step():
            port_step(s, y, n)

        threads = oracle.max_threads()
        what = ("oracle.linear_listnet (numpy float64 port of Linear + ListNet + weight gradient, BLAS threads)"
                if fused_F else "oracle/ltr_oracle.c (C port of the reference algorithm, OpenMP over queries); "
                                "baseline/_ref (the Python reference) is absent in this run")
        dtype = "f64"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    qps = sample_B * args.steps / dt
    sample = f"{sample_B} of the {cfg['B']} queries of one batch per step, {args.steps} steps; {what}"
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if cfg.get("strong") else "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": config_block(cfg, world),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind, "sample": sample,
                         "c_port_queries_per_s": 1.0 / port_per_q,
                         "c_port_note": "oracle/ltr_oracle.c with OpenMP on a 64-query probe (never builds the "
                                        "(B, L, L, 2) pair tensors: a stronger baseline than the reference itself)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    peak, peak_src = hbm_peak()
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
        cb = config_block(cfg, world)
        cb.update({"F": F, "loss": "LinearListNet"})
        line = {
            "metric": "loss fwd+bwd queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cb,
            "details": {"B_per_gpu": B, "global_batch": B * world, "launch": launch,
                        "l2_policy": "inputs larger than L2: one 891 MB feature batch per step",
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


# --------------------------------------------------------------------------- one workload on the device
def measure_config(args, name, cfg, rank, world, dev, steps, warmup, windows, sampler=None):
    """Device-resident measurement of one workload: the step through the public API (CUDA graph per
    resident batch), K timed steps + extra windows, the dominant kernel alone, a parity spot check.
    Returns a dict; `ms` is this rank's time for the K steps (the caller takes the max over ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import pytorchltr_b200.evaluation as ltr_eval
    import pytorchltr_b200.loss as ltr_loss
    from pytorchltr_b200 import _lib
    from pytorchltr_b200.distributed import global_sum_count, shard_bounds, sharded_mean_loss

    L = cfg["L"]
    if cfg.get("strong"):
        lo, hi = shard_bounds(cfg["B"], rank, world)
        B = hi - lo
    else:
        B = cfg["B"]
    is_metric = "metric" in cfg
    family, mode = family_mode(cfg)
    if is_metric:
        if mode == "arp":
            loss_fn = ltr_eval.arp
        else:
            _fn, _k = getattr(ltr_eval, mode), cfg.get("k")
            loss_fn = lambda s, y, n: _fn(s, y, n, k=_k)  # noqa: E731
    else:
        loss_fn = getattr(ltr_loss, cfg["loss"])()

    # ---- resident pool of distinct batches, larger than L2 ---------------------------------
    in_bytes = B * L * 12 + B * 8
    pool_n = min(64, max(4, -(-2 * L2_BYTES // in_bytes)))
    pool, pairs_per_batch = [], []
    for i in range(pool_n):
        s, y, n = make_batch_numpy(1234 + 1000 * rank + i, B, L, cfg.get("skew", False))
        pairs_per_batch.append(valid_pairs(n))
        pool.append((torch.from_numpy(s).to(dev).requires_grad_(not is_metric), torch.from_numpy(y).to(dev),
                     torch.from_numpy(n).to(dev)))
    pool_bytes = pool_n * in_bytes
    torch.cuda.synchronize()
    outs = [None] * pool_n
    global_queries = cfg["B"] if cfg.get("strong") else B * world
    # the step's scalar exchange: NVLink peer-memory mailboxes (ltr_p2p_*), NCCL if the mapping fails
    exchange, exchange_kind = None, "none"
    if world > 1 and not is_metric:
        exchange_kind = "nccl all_reduce"
        if os.environ.get("LTR_EXCHANGE", "p2p") == "p2p":
            try:
                from pytorchltr_b200.distributed import PeerScalarExchange
                exchange = PeerScalarExchange()
                exchange_kind = "ltr_p2p_allreduce_sum (NVLink peer-memory mailboxes, one single-CTA kernel)"
            except Exception as e:  # pragma: no cover
                sys.stderr.write(f"[bench] peer-memory exchange unavailable ({e!r}); using NCCL\n")

    def step(i):
        s, y, n = pool[i % pool_n]
        if is_metric:
            out = loss_fn(s, y, n)
            if world > 1:
                global_sum_count(out)      # [sum metric, #queries] all-reduce behind the global mean
        else:
            s.grad = None
            if world > 1:
                # the path's only exchange: the loss kernel's epilogue leaves the local sum on the
                # device, the 2-element all-reduce follows inside the same captured step
                out = sharded_mean_loss(loss_fn, s, y, n, global_count=global_queries, exchange=exchange)
                out.backward()
            else:
                out = loss_fn(s, y, n)
                out.mean().backward()
        outs[i % pool_n] = out
        return out

    # ---- CUDA graphs: one per pool entry --------------------------------------------------
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
                # thread_local: NCCL's watchdog thread may touch the device while a collective is captured
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    step(i)
                graphs.append(g)
            launch = "cuda_graph (collective captured)" if world > 1 else "cuda_graph"
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); falling back to eager\n")
            graphs = None
            torch.cuda.synchronize()

    def run_step(i):
        if graphs is not None:
            graphs[i % pool_n].replay()
        else:
            step(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def window(first, count):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if world > 1:
            # One more untimed step behind the host barrier: its collective lines the ranks' device
            # timelines up, so that the first timed step does not absorb the host-side start skew of the
            # slowest rank (milliseconds, against a step of half a millisecond at N = 8).
            run_step(first - 1)
        ev0.record()
        for i in range(count):
            run_step(first + i)
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1)

    for i in range(warmup):
        run_step(i)
    barrier()
    if sampler is not None:
        sampler.start()
    ms = window(warmup, steps)                       # the contract's K timed steps
    extra = [window(warmup + (w + 1) * steps, steps) / steps for w in range(max(0, windows))]
    clock_window = "timed region + %d more windows of the same load" % len(extra)
    if sampler is not None:
        # keep the same load running until NVML has a few samples; the decision and the step count are
        # made collective, because every replay carries a collective at N > 1
        need = torch.tensor([1.0 if len(sampler.samples) < 5 else 0.0, ms / steps], device=dev)
        if world > 1:
            dist.all_reduce(need, op=dist.ReduceOp.MAX)
        if float(need[0]) > 0:
            n_more = max(1, int(0.5 / max(float(need[1]) * 1e-3, 1e-6)))
            for i in range(n_more):
                run_step(i)
                if i % 64 == 63:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            clock_window += " + ~0.5 s more"
        sampler.stop()
    all_ms = [ms / steps] + extra
    win = {"n": len(all_ms), "steps_each": steps, "best_ms_per_step": min(all_ms),
           "median_ms_per_step": statistics.median(all_ms), "max_ms_per_step": max(all_ms),
           "scope": "this rank's CUDA events (rank 0)"}

    # ---- parity spot check of the timed path (not timed): batch 0 vs the oracle on a sample ---
    parity = None
    run_step(0)                                      # every rank: the captured collective must match
    torch.cuda.synchronize()
    if rank == 0:
        import oracle
        s, y, n = pool[0]
        idx = np.arange(0, B, max(1, B // 16))
        sn, yn, nn = s.detach().cpu().numpy()[idx], y.cpu().numpy()[idx], n.cpu().numpy()[idx]
        # (hinge: float32 restatement -- pairs on the kink flip between f32 and f64 rounding)
        rl, rg = oracle_step(oracle, family, mode, cfg, sn, yn, nn)
        if is_metric:
            got = outs[0].detach().cpu().double().numpy()[idx]
            err = float(np.abs(got - rl).max())
            parity = {"max_abs_err": err, "queries_checked": int(len(idx)), "ok": err <= 1e-5}
        else:
            # the step is the MEAN over all queries of all ranks: undo the 1 / B_global factor
            got = s.grad.detach().cpu().double().numpy()[idx] * float(cfg["B"] if cfg.get("strong") else B * world)
            gerr = float((np.abs(got - rg) / (np.abs(rg).max(axis=1, keepdims=True) + 1e-30)).max())
            parity = {"max_grad_err_rel_to_rowmax": gerr, "queries_checked": int(len(idx)), "ok": gerr <= 1e-5}
            if world > 1:
                parity["global_mean_loss"] = float(outs[0].detach().cpu())

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

    for i in range(min(8, pool_n)):
        kernel_only(i)
    torch.cuda.synchronize()
    approx_ms = max(ms / steps, 1e-3)
    reps = max(1, min(20, int(200.0 / (approx_ms * pool_n)) or 1))
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
    peak, peak_src = hbm_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name, {}).get("dram_bytes_per_launch")
        if traffic is not None and world > 1 and cfg.get("strong"):
            traffic = traffic / world            # captured at N = 1: per-launch bytes scale with the shard
    roofline = {"bound": "hbm", "kernel": "ltr_rank_metrics kernel" if is_metric else
                "fused loss+gradient kernel (ltr_lambda / ltr_pairwise_additive / ltr_listnet)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src}
    issue = None
    if family not in ("listnet", "metric"):
        mean_pairs = sum(pairs_per_batch) / len(pairs_per_batch)
        ach_pairs = mean_pairs / (kernel_ms * 1e-3)
        if "Hinge" in cfg["loss"]:
            # hinge: FP32 issue, ~6 lane-ops per pair at 128 lanes / clk / SM (measured ~118, issue_peaks.json)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_pairs = sms * 1.965e9 * 128.0 / 6.0
            issue = {"bound": "fp32_issue", "achieved": ach_pairs, "peak": peak_pairs, "unit": "unordered pairs/s",
                     "frac": ach_pairs / peak_pairs, "pairs_per_launch": mean_pairs}
            if 128 < L <= 4096 and os.environ.get("LTR_HINGE") != "pairs" \
                    and os.environ.get("LTR_KERNEL") not in ("tiles", "generic"):
                # the hinge losses do not enumerate pairs at these sizes (sort + scans, O(n log n)):
                # "achieved" is the pair rate an O(L^2) kernel would need to match it, not an issue rate
                issue.update({"bound": "latency (O(n log n) sort + scans; pairs are not enumerated)",
                              "frac": None, "equivalent_pair_rate_vs_fp32_issue_peak": ach_pairs / peak_pairs})
        else:
            mufu_rate, mufu_src = mufu_peak()
            peak_pairs = mufu_rate / MUFU_PER_PAIR
            issue = {"bound": "mufu", "achieved": ach_pairs, "peak": peak_pairs, "unit": "unordered pairs/s",
                     "frac": ach_pairs / peak_pairs, "pairs_per_launch": mean_pairs,
                     "mufu_per_pair": MUFU_PER_PAIR, "mufu_lane_ops_per_s": mufu_rate, "peak_source": mufu_src,
                     "frac_vs_2_mufu_per_pair": ach_pairs / (mufu_rate / 2.0),
                     "note": "binding roofline of the O(L^2) losses (SURVEY.md F4): the kernel needs lg2 of every "
                             "pair and one rcp per two pairs on the 16-lane/clk/SM MUFU pipe"}
    launches_per_step = 1 if is_metric else (2 if family == "listnet" else 3)
    if exchange is not None and exchange.timed_out():
        raise RuntimeError("a peer-memory exchange timed out")
    return {"ms": ms, "B": B, "L": L, "launch": launch, "exchange": exchange_kind, "pool_n": pool_n, "pool_bytes": pool_bytes,
            "windows": win, "parity": parity, "roofline": roofline, "issue": issue, "loss_fn": loss_fn,
            "is_metric": is_metric, "family": family, "mode": mode, "metric_ld": metric_ld,
            "launches_per_step": launches_per_step, "clock_window": clock_window}


def run_e2e(cfg, m, rank, world, dev, steps, compact):
    """Public API with pinned HOST tensors: H2D of scores / relevance / n, kernel, D2H of the loss and of
    the gradient inside the timed region.  `compact`: uint8 relevance and int32 n (extension)."""
    import torch
    import torch.distributed as dist
    B, L, is_metric, loss_fn = m["B"], m["L"], m["is_metric"], m["loss_fn"]
    host_n = 2 if B * L > (1 << 24) else 4
    host_pool = []
    for i in range(host_n):
        s, y, n = make_batch_numpy(99 + 1000 * rank + i, B, L, cfg.get("skew", False))
        ty, tn = torch.from_numpy(y), torch.from_numpy(n)
        if compact:
            ty, tn = ty.to(torch.uint8), tn.to(torch.int32)
        host_pool.append((torch.from_numpy(s).pin_memory().requires_grad_(not is_metric), ty.pin_memory(),
                          tn.pin_memory()))

    def e2e_step(i):
        s, y, n = host_pool[i % host_n]
        if is_metric:
            return loss_fn(s, y, n)        # H2D scores/relevance/n, kernel, D2H metric
        s.grad = None
        out = loss_fn(s, y, n)             # H2D scores/relevance/n, kernel, D2H loss
        out.mean().backward()              # device row scale by the broadcast 1/B, D2H gradient
        return out

    for i in range(3):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    total_queries = cfg["B"] if cfg.get("strong") else B * world
    rel_b, n_b = (1, 4) if compact else (8, 8)
    return {"value": total_queries * steps / dt, "unit": "queries/s",
            "h2d_bytes_per_step": B * L * (4 + rel_b) + B * n_b,
            "d2h_bytes_per_step": B * 4 * m["metric_ld"] if is_metric else B * 4 + B * L * 4,
            "steps": steps, "ms_per_step": dt / steps * 1e3,
            "dtypes": "float32 scores, %s relevance, %s n" % (("uint8", "int32") if compact else ("int64", "int64")),
            "path": "loss_fn(pinned CPU tensors).mean().backward(): loss and gradient returned as CPU tensors"
                    if not is_metric else "metric(pinned CPU tensors): result returned as a CPU tensor"}


def cpu_port_baseline(cfg, m):
    import oracle
    oracle.set_threads(os.cpu_count() or 1)
    threads = oracle.max_threads()
    B, L = m["B"], m["L"]
    family, mode = m["family"], m["mode"]
    s, y, n = make_batch_numpy(1234, min(B, 64), L, cfg.get("skew", False))
    oracle_step(oracle, family, mode, cfg, s, y, n)
    t0 = time.perf_counter()
    oracle_step(oracle, family, mode, cfg, s, y, n)
    per_q = (time.perf_counter() - t0) / len(n)
    sample_B = int(max(threads * 8, min(4 * B, 8.0 / max(per_q, 1e-9))))
    s, y, n = make_batch_numpy(1235, sample_B, L, cfg.get("skew", False))
    t0 = time.perf_counter()
    oracle_step(oracle, family, mode, cfg, s, y, n)
    dt = time.perf_counter() - t0
    return {"value": sample_B / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{sample_B} queries of the same synthetic workload in one call "
                      f"({dt:.1f} s wall), oracle/ltr_oracle.c with OpenMP over queries"}


def python_reference_baseline(name, cfg, dev):
    """The unmodified Python reference on the host cores and eager on this GPU: live when baseline/_ref
    is present, else the committed measurement."""
    ref = import_reference()
    if ref is None or reference_callable(ref, cfg)[0] is None:
        return committed_reference_timing(name)
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"source": "live: unmodified rjagerman/pytorchltr from baseline/_ref, "
                     "loss_fn(scores, relevance, n).mean().backward()",
           "cpu_cores": os.cpu_count()}
    chunk = pair_chunk(cfg, 4e9)
    t = time_python_reference(ref, cfg, torch.device("cpu"), chunk, 3)
    out.update({"cpu_queries_per_s": chunk / statistics.median(t), "cpu_chunk": chunk})
    try:
        gchunk = pair_chunk(cfg, 40e9)
        t = time_python_reference(ref, cfg, dev, gchunk, 5)
        out.update({"gpu_eager_queries_per_s": gchunk / statistics.median(t), "gpu_eager_chunk": gchunk})
    except Exception as e:  # pragma: no cover
        out["gpu_eager_unavailable"] = repr(e)[:200]
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- our arm
def measure_mlp(args, cfg, dev, rank=0):
    """Config c4mlp: `loss_fn(MLPRanker(F)(xs), ys, n).mean().backward()` -- the training step of the reference's
    getting-started guide -- timed as a CUDA graph, kernel by kernel through the C ABI, and against the same
    model built from torch.nn.Linear layers."""
    import numpy as np
    import torch

    import pytorchltr_b200.loss as losses
    from pytorchltr_b200 import _lib
    from pytorchltr_b200.fused import MLPRanker

    B, L, F = cfg["B"], cfg["L"], cfg["F"]
    rows = B * L
    _, y_np, n_np = make_batch_numpy(4321 + 1000 * rank, B, L, True)
    gen = torch.Generator(device=dev).manual_seed(78 + rank)
    xs = torch.randn(B, L, F, device=dev, generator=gen)            # 891 MB: larger than L2 by itself
    ys, ns = torch.from_numpy(y_np).to(dev), torch.from_numpy(n_np).to(dev)
    torch.manual_seed(6)
    model = MLPRanker(F).to(dev)
    plain = torch.nn.Sequential(torch.nn.Linear(F, 50), torch.nn.ReLU(), torch.nn.Linear(50, 10), torch.nn.ReLU(),
                                torch.nn.Linear(10, 1)).to(dev)
    with torch.no_grad():
        for a, b in zip((plain[0], plain[2], plain[4]), (model.l1, model.l2, model.l3)):
            a.weight.copy_(b.weight)
            a.bias.copy_(b.bias)
    loss_fn = getattr(losses, cfg["loss"])()

    def step(m=model):
        m.zero_grad(set_to_none=True)
        out = loss_fn(m(xs), ys, ns)
        out.mean().backward()
        return out

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    launch, run = "eager", step
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
            run, launch = graph.replay, "cuda_graph"
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture of the MLP step failed ({e!r}); eager\n")
            torch.cuda.synchronize()
    steps = max(10, min(args.steps, 50))
    ms = timed(run, steps, 3)
    unfused = {}
    old = torch.backends.cuda.matmul.allow_tf32
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        unfused["allow_tf32" if tf32 else "fp32"] = timed(lambda: step(plain), 5, 2)
    torch.backends.cuda.matmul.allow_tf32 = old

    # kernel by kernel through the C ABI
    lib = _lib.lib()
    p = [t.detach().contiguous() for t in (model.l1.weight, model.l1.bias, model.l2.weight, model.l2.bias,
                                           model.l3.weight, model.l3.bias)]
    x2 = xs.reshape(rows, F)
    scores = torch.empty(rows, device=dev)
    ds = torch.randn(rows, device=dev, generator=gen)
    glen = lib.ltr_mlp_grad_len(F, 50, 10)
    grads = torch.empty(glen, device=dev)
    wsb = lib.ltr_mlp_workspace_bytes(F, 50, 10)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    hz = torch.empty(rows, lib.ltr_mlp_hz_pitch(50, 10), device=dev)

    def fwd(keep=True):
        _lib.check(lib.ltr_mlp_scores(x2.data_ptr(), rows, F, p[0].data_ptr(), p[1].data_ptr(), 50, p[2].data_ptr(),
                                      p[3].data_ptr(), 10, p[4].data_ptr(), p[5].data_ptr(), scores.data_ptr(),
                                      hz.data_ptr() if keep else None, st))

    def bwd(kept=True):
        _lib.check(lib.ltr_mlp_backward(x2.data_ptr(), rows, F, p[0].data_ptr(), p[1].data_ptr(), 50, p[2].data_ptr(),
                                        p[3].data_ptr(), 10, p[4].data_ptr(), p[5].data_ptr(),
                                        hz.data_ptr() if kept else None, ds.data_ptr(), grads.data_ptr(),
                                        ws.data_ptr(), wsb, st))

    fwd_ms, bwd_ms = timed(fwd, 20, 3), timed(bwd, 20, 3)
    fwd_infer_ms = timed(lambda: fwd(False), 20, 3)
    bwd_recompute_ms = timed(lambda: bwd(False), 20, 3)
    sc2 = scores.reshape(B, L).clone()

    def loss_only():
        loss_fn(sc2, ys, ns)

    loss_ms = timed(loss_only, 20, 3)
    hzb = 4 * hz.shape[1]                                      # kept activation row
    alg = rows * (4 * F + 4 + hzb)                             # features + activations once + score / upstream gradient
    alg_infer = rows * (4 * F + 4)
    peak, peak_src = hbm_peak()
    flops_fwd = 2.0 * rows * (F * 50 + 50 * 10 + 10)
    flops_bwd = 2.0 * rows * (F * 50 + 50 * 10 + 10) + 2.0 * rows * (F * 50 + 2 * 50 * 10 + 10)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    ncu = {}
    if os.path.exists(tpath):
        ncu = json.load(open(tpath)).get("c4mlp") or {}

    # parity spot check against the float64 restatement with TF32 operands (not timed)
    import oracle
    pn = [t.cpu().numpy() for t in p]
    idx = torch.arange(0, rows, rows // 4096, device=dev)[:4096]
    fwd()
    torch.cuda.synchronize()
    xn = x2[idx].cpu().numpy()
    ref = oracle.mlp_scores(xn, *pn, tf32=True)
    got = scores[idx].cpu().double().numpy()
    serr = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
    nsub = 128 * 64
    dsn = ds[:nsub].clone()
    rc = lib.ltr_mlp_backward(x2.data_ptr(), nsub, F, p[0].data_ptr(), p[1].data_ptr(), 50, p[2].data_ptr(),
                              p[3].data_ptr(), 10, p[4].data_ptr(), p[5].data_ptr(), hz.data_ptr(), dsn.data_ptr(),
                              grads.data_ptr(), ws.data_ptr(), wsb, st)
    _lib.check(rc)
    torch.cuda.synchronize()
    gref = oracle.mlp_grads_kept(x2[:nsub].cpu().numpy(), *pn, dsn.cpu().numpy())
    gref = np.concatenate([g.reshape(-1) for g in gref])
    gerr = float(np.linalg.norm(grads.cpu().double().numpy() - gref) / max(1e-30, np.linalg.norm(gref)))
    parity = {"scores_max_err_rel": serr, "grads_norm_err_rel": gerr, "rows_checked": 4096 + nsub,
              "oracle": "oracle.mlp_scores(tf32=True) / mlp_grads_kept, float64 on TF32-truncated operands",
              "ok": bool(serr <= 2e-5 and gerr <= 2e-2)}
    return {
        "workload": cfg["workload"], "B": B, "L": L, "F": F, "steps": steps, "launch": launch,
        "ms_per_step": ms, "queries_per_s": B / (ms * 1e-3),
        # uniform keys of the sub-config block: the two scorer kernels together, against the HBM floor of both
        "kernel_ms": fwd_ms + bwd_ms, "hbm_frac": 2 * alg / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak,
        "hbm_gbs": 2 * alg / ((fwd_ms + bwd_ms) * 1e-3) / 1e9, "issue_frac": None, "issue_bound": None,
        "kernels_ms": {"ltr_mlp_scores (keeping H1 | Z2 for the backward pass)": fwd_ms,
                       "loss (" + cfg["loss"] + ", forward + gradient)": loss_ms,
                       "ltr_mlp_backward (from the kept activations)": bwd_ms,
                       "ltr_mlp_scores (inference: scores only)": fwd_infer_ms,
                       "ltr_mlp_backward (recomputing layer 1, no kept activations)": bwd_recompute_ms},
        "inference_hbm_frac": alg_infer / (fwd_infer_ms * 1e-3) / 1e9 / peak,
        "unfused_torch_ms_per_step": unfused,
        "roofline_forward": {"bound": "hbm", "kernel": "mlp_scores_kernel<50, 10>", "achieved": alg / (fwd_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg / (fwd_ms * 1e-3) / 1e9 / peak,
                             "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                             "tflops_tf32": flops_fwd / (fwd_ms * 1e-3) / 1e12, "traffic": ncu.get("fwd_dram_bytes"),
                             "tensor_pipe_active_pct": ncu.get("fwd_tensor_pct")},
        "roofline_backward": {"bound": "hbm", "kernel": "mlp_backward_hz_kernel<50, 10>",
                              "achieved": alg / (bwd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": alg / (bwd_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                              "peak_source": peak_src, "tflops_tf32": flops_bwd / (bwd_ms * 1e-3) / 1e12,
                              "traffic": ncu.get("bwd_dram_bytes"), "tensor_pipe_active_pct": ncu.get("bwd_tensor_pct")},
        "parity_spot_check": parity,
    }


def run_mlp(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure_mlp(args, cfg, dev, rank)
    sampler.stop()
    ms = m["ms_per_step"]
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        cb = config_block(cfg, world)
        cb.update({"F": cfg["F"], "model": "MLPRanker 136-50-10-1"})
        line = {
            "metric": "loss fwd+bwd queries/sec", "value": cfg["B"] * world / (ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": m["steps"], "warmup": 3, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (layer 1) / f32", "data": "synthetic", "config": cb,
            "details": {k: m[k] for k in ("launch", "kernels_ms", "unfused_torch_ms_per_step")},
            "roofline": m["roofline_forward"], "roofline_backward": m["roofline_backward"], "cpu_baseline": None,
            "e2e": None, "clocks": sampler.summary("timed region"), "gpu_launches": 6 * m["steps"],
            "gpu_launches_note": "per step: mlp_scores, loss kernel (+ ordering), row scale, mlp_backward, reduce",
            "parity_spot_check": m["parity_spot_check"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args, name, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(2, device=dev))          # communicator set-up before any capture
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    m = measure_config(args, name, cfg, rank, world, dev, args.steps, args.warmup, args.windows, sampler)
    ms = m["ms"]
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    B, L = m["B"], m["L"]
    total_queries = cfg["B"] if cfg.get("strong") else B * world
    value = total_queries * args.steps / (ms * 1e-3)

    e2e = e2e_compact = None
    if not args.no_e2e:
        e_steps = max(5, min(args.steps, 100)) if B * L <= (1 << 24) else max(3, min(args.steps, 10))
        e2e = run_e2e(cfg, m, rank, world, dev, e_steps, compact=False)
        e2e_compact = run_e2e(cfg, m, rank, world, dev, e_steps, compact=True)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_port_baseline(cfg, m)
        cpu_baseline["python_reference"] = python_reference_baseline(name, cfg, dev)

    # ---- the other BASELINE configurations, same run (N = 1 only) -----------------------------
    sub = {}
    if world == 1 and not args.no_sub and args.sub:
        import gc
        for sname in [x for x in args.sub.split(",") if x and x != name]:
            scfg = CONFIGS[sname]
            if scfg.get("fused") == "mlp":
                m.pop("loss_fn", None)
                gc.collect()
                torch.cuda.empty_cache()
                try:
                    sub[sname] = measure_mlp(args, scfg, dev)
                except Exception as e:  # pragma: no cover
                    sub[sname] = {"workload": scfg["workload"], "error": repr(e)}
                gc.collect()
                torch.cuda.empty_cache()
                continue
            if scfg.get("fused"):
                continue
            m.pop("loss_fn", None)
            gc.collect()
            torch.cuda.empty_cache()
            approx = {"ns": 0.9, "c3": 0.08, "c5": 4.0}.get(sname, 0.04)
            s_steps = int(max(20, min(2000, 60.0 / approx)))
            sm = measure_config(args, sname, scfg, 0, 1, dev, s_steps, max(3, s_steps // 10), 2)
            entry = {"workload": scfg["workload"], "B": sm["B"], "L": sm["L"], "steps": s_steps,
                     "ms_per_step": sm["ms"] / s_steps, "queries_per_s": sm["B"] * s_steps / (sm["ms"] * 1e-3),
                     "windows": sm["windows"], "kernel_ms": sm["roofline"]["kernel_ms"],
                     "hbm_frac": sm["roofline"]["frac"], "hbm_gbs": sm["roofline"]["achieved"],
                     "issue_frac": (sm["issue"] or {}).get("frac"), "issue_bound": (sm["issue"] or {}).get("bound"),
                     "pairs_per_s": (sm["issue"] or {}).get("achieved"),
                     "parity_spot_check": sm["parity"], "launch": sm["launch"]}
            if not args.no_e2e and sname in ("c2", "c4m"):
                entry["e2e"] = run_e2e(scfg, sm, 0, 1, dev, 50, compact=False)
                entry["e2e_compact"] = run_e2e(scfg, sm, 0, 1, dev, 50, compact=True)
            if not args.no_cpu_baseline:
                entry["python_reference"] = committed_reference_timing(sname)
            sub[sname] = entry

    if rank == 0:
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if cfg.get("strong") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, world),
            "details": {"B_per_gpu": B, "global_batch": total_queries, "launch": m["launch"],
                        "exchange": m["exchange"],
                        "step": "sharded_mean_loss(loss_fn, scores, relevance, n, global_count=B, exchange=...)"
                                ".backward() (scalar all-reduce, fed by the loss kernel's epilogue, inside the "
                                "captured step; one untimed aligning step between the host barrier and the first "
                                "timed step)" if world > 1 and not m["is_metric"]
                        else "loss_fn(scores, relevance, n).mean().backward()" if not m["is_metric"]
                        else "metric(scores, relevance, n)",
                        "l2_policy": f"inputs larger than L2: {m['pool_n']} distinct resident batches "
                                     f"({m['pool_bytes'] / 2**20:.0f} MiB) visited round-robin"},
            "windows": m["windows"],
            "roofline": m["roofline"], "issue_roofline": m["issue"], "cpu_baseline": cpu_baseline,
            "e2e": e2e, "e2e_compact": e2e_compact,
            "clocks": sampler.summary(m["clock_window"]),
            "gpu_launches": m["launches_per_step"] * args.steps,
            "gpu_launches_note": "per step: 1 fused kernel (+ 1 row-scale kernel in the backward pass, + 1 "
                                 "query-ordering kernel before the O(L^2) losses when queries queue) of "
                                 "libltr_sm100.so (plus torch's mean / fill kernels and, for N > 1, NCCL's all-reduce)",
            "parity_spot_check": m["parity"],
            "configs": sub or None,
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
        run_reference(args, args.config, cfg, rank, world)
        return
    if world != args.gpus and rank == 0:
        sys.stderr.write(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun "
                         f"for N > 1; running {world} rank(s)\n")
    if cfg.get("fused") == "mlp":
        run_mlp(args, cfg, rank, local_rank, world)
        return
    if cfg.get("fused"):
        run_fused(args, cfg, rank, local_rank, world)
        return
    run_ours(args, args.config, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
