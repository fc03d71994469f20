"""Throughput of the SVMrank text parser (N4) on a synthetic MSLR-shaped file (136 features per line): parse phase,
fill phase, thread scaling; optionally against the reference's Cython parser when /root/reference is present.
python tools/parser_bench.py [lines] [--ref]"""
import ctypes
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from pytorchltr_b200.datasets import svmrank as sv  # noqa: E402


def make_file(path, lines, feats=136, seed=0):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        chunk = 2000
        for start in range(0, lines, chunk):
            n = min(chunk, lines - start)
            y = rng.integers(0, 5, n)
            q = (start + np.arange(n)) // 120 + 1
            x = rng.random((n, feats)) * rng.choice([1.0, 100.0, 1e4], size=(1, feats))
            x[rng.random((n, feats)) < 0.3] = 0.0
            out = []
            for i in range(n):
                out.append(f"{y[i]} qid:{q[i]} " + " ".join(f"{j + 1}:{x[i, j]:.6g}" for j in range(feats)))
            f.write("\n".join(out) + "\n")


def main():
    lines = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 60000
    path = os.path.join(tempfile.gettempdir(), f"ltr_parser_bench_{lines}.txt")
    if not os.path.exists(path):
        make_file(path, lines)
    size = os.path.getsize(path)
    print(f"{path}: {size / 1e6:.1f} MB, {lines} lines")
    open(path, "rb").read()                                # page cache
    for threads in (1, 2, 4, 8, 0):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            xs, ys, qids = sv.parse_svmrank_file(path, n_threads=threads) if "n_threads" in sv.parse_svmrank_file.__code__.co_varnames \
                else sv.parse_svmrank_file(path)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"threads={threads or os.cpu_count()}: {best * 1e3:8.1f} ms  {size / best / 1e6:8.1f} MB/s  -> xs {xs.shape} {xs.dtype}")
    if "--ref" in sys.argv:
        sys.path.insert(0, "/root/repo/baseline/_ref")
        from pytorchltr.datasets.svmrank.parser import parse_svmrank_file as ref_parse
        t0 = time.perf_counter()
        rx, ry, rq = ref_parse(path)
        dt = time.perf_counter() - t0
        print(f"reference parser: {dt * 1e3:8.1f} ms  {size / dt / 1e6:8.1f} MB/s; identical: {bool((rx == xs).all())}")


if __name__ == "__main__":
    main()
