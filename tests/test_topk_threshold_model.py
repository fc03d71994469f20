"""The selection rule of topk_metrics_warp_kernel (csrc/ltr_metrics_warp.cuh) restated in numpy: with
the documents dealt to 32 lanes (document j on lane j % 32), the k-th smallest of the 32 lane minima
is a threshold T such that the documents with key <= T (a) number at least k and (b) contain the true
top-k -- whatever the data, ties included.  The kernel sorts those candidates exactly; if they number
more than 32 it falls back to one argmin per rank."""
import numpy as np
import pytest


@pytest.mark.parametrize("seed", range(20))
def test_threshold_from_lane_minima_keeps_the_top_k(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.integers(1, 1025))
    nb = int(rng.integers(0, L + 1))
    k = int(rng.integers(1, 33))
    PAD = np.uint64(0xFFFFFFFF)
    keys = rng.integers(0, 50 if seed % 3 == 0 else 2**32 - 1, size=L).astype(np.uint64)   # seed % 3 == 0: heavy ties
    keys[nb:] = PAD
    kv = min(k, L, nb)
    if kv == 0:
        return
    lanes = [keys[lane::32] for lane in range(32)]
    lane_min = np.array([x.min() if x.size else PAD for x in lanes], dtype=np.uint64)
    T = np.sort(lane_min)[kv - 1]
    assert T != PAD                                   # at least kv lanes hold a valid document
    cand = np.flatnonzero(keys <= T)
    assert cand.size >= kv
    exact = np.lexsort((np.arange(L), keys))[:kv]     # (key, index) order: ties lowest index first
    assert set(exact.tolist()) <= set(cand.tolist())
