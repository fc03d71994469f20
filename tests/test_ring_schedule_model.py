"""Exhaustive CPU check of the pair schedule of pair_ring_kernel (csrc/ltr_pair_ring.cuh, ring_split):
a Python restatement of its index arithmetic -- chunk ring, row groups, the static split over 8 warps
and the replica lanes of a sparsely filled last row group -- must

  * visit every unordered pair of 4-rank chunks exactly once (plus every chunk's own triangle),
  * never let two lanes of one warp step update the same column chunk (the kernel's column
    gradients are plain vector read-modify-writes, no atomics),

for every chunk count the kernel can see (C = 32 .. 256, i.e. n = 125 .. 1024).  The GPU parity tests
check the numbers; this checks the combinatorics for every size, not just the sampled ones.
"""
import pytest

WARPS = 8


def schedule(C):
    """Yields (warp, iteration, lane, row_chunk, col_chunk or None for the triangle, commits)."""
    G = (C + 31) >> 5
    M = C >> 1
    cl_n = C - ((G - 1) << 5)
    f = 1
    while cl_n * f * 2 <= 32:
        f <<= 1
    T = (M + f) // f
    while f > 1 and (T < cl_n or (f - 1) * T + cl_n > C):
        f >>= 1
        T = (M + f) // f
    U_full = (G - 1) * (M + 1)
    U = U_full + T
    even = C % 2 == 0
    out = []
    for warp in range(WARPS):
        u, u_end = U * warp // WARPS, U * (warp + 1) // WARPS
        it = 0
        while u < u_end:
            last = u >= U_full
            g = G - 1 if last else u // (M + 1)
            m0 = u - g * (M + 1)
            replicated = last and f > 1
            steps_g = T if replicated else M + 1
            m1 = min(steps_g, m0 + (u_end - u))
            u += m1 - m0
            for t in range(m0, m1):
                for lane in range(32):
                    sub = lane // cl_n if replicated else 0
                    cl = lane - sub * cl_n if replicated else lane
                    c = (g << 5) + cl
                    active = sub < f if replicated else c < C
                    if not active:
                        continue
                    m = sub * T + t if replicated else t
                    if m == 0:
                        if sub == 0:
                            out.append((warp, it, lane, c, None, True))
                        continue
                    if m > M:
                        continue
                    pc = c + m
                    wrapped = pc >= C
                    if wrapped:
                        pc -= C
                    commit = not (even and m == M and wrapped)
                    out.append((warp, it, lane, c, pc, commit))
                it += 1
    return out


@pytest.mark.parametrize("C", list(range(32, 257)))
def test_every_chunk_pair_once_and_no_column_collisions(C):
    sched = schedule(C)
    pairs, triangles = {}, {}
    by_step = {}
    for warp, it, lane, c, pc, commit in sched:
        if pc is None:
            triangles[c] = triangles.get(c, 0) + 1
            continue
        if commit:
            key = (min(c, pc), max(c, pc))
            assert c != pc
            pairs[key] = pairs.get(key, 0) + 1
            by_step.setdefault((warp, it), []).append(pc)
    assert triangles == {c: 1 for c in range(C)}
    assert len(pairs) == C * (C - 1) // 2 and set(pairs.values()) == {1}
    for cols in by_step.values():
        assert len(cols) == len(set(cols)), "two lanes of one step update the same column chunk"


def test_work_split_is_even():
    for C in (32, 33, 64, 65, 100, 129, 150, 192, 200, 255, 256):
        sched = schedule(C)
        steps = [len({it for w, it, *_ in sched if w == warp}) for warp in range(WARPS)]
        assert max(steps) - min(steps) <= 2, (C, steps)


@pytest.mark.parametrize("C", list(range(2, 33)))
def test_single_warp_ring_every_chunk_pair_once(C):
    """ring_pass (csrc/ltr_pair_tiles.cuh): one warp, lane l owns chunk l, step m meets chunk
    (l + m) mod C for m = 1 .. C // 2; on an even ring the last step is met from both ends and only
    the lanes below C / 2 commit."""
    steps = C // 2
    even = C % 2 == 0
    pairs = {}
    for m in range(1, steps + 1):
        cols = []
        for lane in range(C):
            pc = (lane + m) % C
            if even and m == steps and lane >= steps:
                continue
            cols.append(pc)
            key = (min(lane, pc), max(lane, pc))
            pairs[key] = pairs.get(key, 0) + 1
        assert len(cols) == len(set(cols))
    assert len(pairs) == C * (C - 1) // 2 and set(pairs.values()) == {1}
