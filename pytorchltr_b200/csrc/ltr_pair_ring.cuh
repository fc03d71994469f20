// ltr_pair_ring.cuh -- one CTA per query for list sizes 129 .. 1024 (the north-star point
// (4096, 1024), configs (65536, 512) and (1024, 1024)): every unordered pair once, on ONE ring
// of 4-rank chunks that spans the whole query, split evenly and statically over the warps.
//
// Per query:
//   1. the padded row (scores + relevance as stored in HBM) arrives by TMA bulk copy
//      (cp.async.bulk + mbarrier) into a staging buffer; it is unpacked at once (scores, int32
//      grades, 64-bit sort keys), which frees the buffer for the NEXT query's row: that copy is in
//      flight during the whole pair phase (plain coalesced loads when the row is not 16-byte
//      aligned);
//   2. rank_by_score (utils/tensor_operations.py:48-64): every warp sorts 128-key blocks in
//      registers (bitonic network on shuffles), only the strides >= 128 of the merge tree go
//      through shared memory: 10 CTA barriers at P = 1024 instead of 55;
//   3. ideal DCG from a shared-memory grade histogram (_max_dcg, pairwise_lambda.py:231-241),
//      per-document factors in rank order (structure of arrays, ltr_pair_tiles.cuh);
//   4. pair phase.  The n ranks form C = ceil(n / 4) chunks.  Lane l of row group g owns chunk
//      c = 32 g + l as ROWS (registers); in ring step m it meets the COLUMNS of chunk (c + m) mod C.
//      Steps m = 1 .. floor(C / 2) cover every unordered chunk pair once, step 0 is the triangle
//      inside a chunk.  The G x (M + 1) grid of (group, step) units is cut into equal contiguous
//      ranges, one per warp: no tile padding (a 128 x 128 tiling of n = 600 evaluates 14 % dead
//      pairs), no dynamic counter, and the warps finish within one step of each other.
//      Every warp adds its row and column gradients to a PRIVATE rank-order array with vector
//      read-modify-writes (the 32 lanes of a step touch 32 different chunks): no atomics, and the
//      result is bit-reproducible;
//   5. the private arrays are summed, scaled, scattered to document order (backward of the gather,
//      pairwise_lambda.py:69) and stored coalesced.
#pragma once

#include "ltr_pair_warp.cuh"

namespace ltr {

#ifndef LTR_RING_MIN_CTAS
#define LTR_RING_MIN_CTAS 3   // resident CTAs per SM the register allocation aims at (A-B timing: -DLTR_RING_MIN_CTAS=4)
#endif
#ifndef LTR_RING_SHORT_CTAS
#define LTR_RING_SHORT_CTAS 6   // same for the 4-warp CTAs of short lists
#endif
constexpr int kRingMaxL = 1024;
// Warps per CTA (= per query).  8 for lists up to 1024 documents; 4 for lists up to 512 (the (65536, 512)
// configuration): half the private gradient arrays (28 KB of shared memory per CTA at L = 512), every
// warp busy in the 128-key block sorts, and twice as many independent CTAs per SM, so that a CTA waiting
// at one of its barriers costs the SM half as much.
constexpr int kRingWarpsLong = 8;
constexpr int kRingWarpsShort = 4;
constexpr int kRingShortL = 640;                    // (4 warps still ahead of 8 at 516..640 documents, equal at 768: tools/ring_warps_ab2.py)
#ifndef LTR_RING_TINY_L
#define LTR_RING_TINY_L 256
#endif
constexpr int kRingTinyL = LTR_RING_TINY_L;         // one or two warps per query up to here

// ---- 128-key blocks sorted in registers -------------------------------------------------------------
// Element index inside the block = lane * 4 + r; `gbase` is the block's first index in the whole
// array, so that the directions of the sub-networks are those of the full bitonic network.
__device__ __forceinline__ void block_sort64(uint64_t (&k)[4], int lane, int gbase) {
  constexpr int E = 4;
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride < E) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int q = r ^ stride;
          if (q > r) {
            const bool up = ((gbase + lane * E + r) & size) == 0;
            const uint64_t a = k[r], b = k[q];
            const bool sw = (a > b) == up;
            k[r] = sw ? b : a;
            k[q] = sw ? a : b;
          }
        }
      } else {
        const int ls = stride / E;
        const bool keep_min = (((gbase + lane * E) & size) == 0) == ((lane & ls) == 0);
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint64_t mine = k[r];
          const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, ls);
          k[r] = (keep_min == (other < mine)) ? other : mine;
        }
      }
    }
  }
}

// strides 64 .. 1 of a merge step whose direction is the same for the whole block
__device__ __forceinline__ void block_merge64(uint64_t (&k)[4], int lane, bool up) {
  constexpr int E = 4;
#pragma unroll
  for (int stride = 16 * E; stride > 0; stride >>= 1) {
    if (stride < E) {
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int q = r ^ stride;
        if (q > r) {
          const uint64_t a = k[r], b = k[q];
          const bool sw = (a > b) == up;
          k[r] = sw ? b : a;
          k[q] = sw ? a : b;
        }
      }
    } else {
      const int ls = stride / E;
      const bool keep_min = up == ((lane & ls) == 0);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const uint64_t mine = k[r];
        const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, ls);
        k[r] = (keep_min == (other < mine)) ? other : mine;
      }
    }
  }
}

__device__ __forceinline__ void block_load(const uint64_t* keys, int blk, int lane, uint64_t (&k)[4]) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(keys + blk * 128 + lane * 4);
  const ulonglong2 a = p[0], b = p[1];
  k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
}
__device__ __forceinline__ void block_store(uint64_t* keys, int blk, int lane, const uint64_t (&k)[4]) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(keys + blk * 128 + lane * 4);
  p[0] = make_ulonglong2(k[0], k[1]);
  p[1] = make_ulonglong2(k[2], k[3]);
}

// Ascending sort of P >= 256 (power of two) 64-bit keys in shared memory by the whole CTA.
// Begins and ends with a barrier.
__device__ __forceinline__ void cta_block_sort(uint64_t* keys, int P, int lane, int warp, int nwarps) {
  const int nblk = P >> 7;
  __syncthreads();
  for (int blk = warp; blk < nblk; blk += nwarps) {
    uint64_t k[4];
    block_load(keys, blk, lane, k);
    block_sort64(k, lane, blk << 7);
    block_store(keys, blk, lane, k);
  }
  __syncthreads();
  for (int size = 256; size <= P; size <<= 1) {
    for (int j = size >> 1; j >= 128; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const bool up = (i & size) == 0;
        const uint64_t a = keys[i], b = keys[l];
        if ((a > b) == up) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
    for (int blk = warp; blk < nblk; blk += nwarps) {
      uint64_t k[4];
      block_load(keys, blk, lane, k);
      block_merge64(k, lane, ((blk << 7) & size) == 0);
      block_store(keys, blk, lane, k);
    }
    __syncthreads();
  }
}

// ---- shared memory ------------------------------------------------------------------------------------
struct RingSmem {
  PairSoA it;          // 4 x [Lp]   rank order, Lp = L rounded up to 128; it.a is reused as the
                       //            document-order gradient once the pair phase is over
  float* gw;           // [warps][Lp] private rank-order gradients.  Until the pair phase starts
                       //            the same storage holds the sort keys and the unpacked row:
  uint64_t* keys;      // [P]        (inside gw)
  float* raw_s;        // [L]        scores, document order (inside gw)
  int* raw_y;          // [L]        relevance, document order (inside gw)
  float* sym;          // [2 Lp + 8] sym[Lp + d] = delta[|d|]: delta windows by SIGNED rank distance
  uint16_t* doc;       // [Lp]       rank -> document
  float* red;          // [40]
  int* hist;           // [40]
  unsigned char* stage;   // TMA staging: [4 L] scores, then [rel_bytes L] relevance
  uint64_t* bar;          // [1]
};

__host__ __device__ inline size_t ring_align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

// tma_rel_bytes = 0: no staging buffer
__host__ __device__ inline size_t ring_smem_bytes(int L, int tma_rel_bytes, int warps) {
  const size_t Lp = (static_cast<size_t>(L) + 127) / 128 * 128;
  // the private arrays also hold the sort keys and the unpacked row before the pair phase: 8 P + 8 L <= 24 L
  const size_t gw = 4u * static_cast<size_t>(warps) * Lp < 24u * Lp ? 24u * Lp : 4u * static_cast<size_t>(warps) * Lp;
  size_t bytes = 16u * Lp + gw + 4u * (2 * Lp + 8) + 2u * Lp + 4u * 80;
  if (tma_rel_bytes) bytes += ring_align16(4u * L) + ring_align16(static_cast<size_t>(tma_rel_bytes) * L) + 16u;
  return bytes;
}

__device__ __forceinline__ RingSmem ring_carve(unsigned char* base, int L, int P, int tma_rel_bytes, int warps) {
  const int Lp = (L + 127) / 128 * 128;
  RingSmem m;
  // 8 P + 8 L <= 24 L bytes of keys and row fit in the private arrays (at least 24 Lp bytes)
  m.gw = reinterpret_cast<float*>(base);
  m.keys = reinterpret_cast<uint64_t*>(base);
  m.raw_s = reinterpret_cast<float*>(base + 8u * P);
  m.raw_y = reinterpret_cast<int*>(base + 8u * P + ring_align16(4u * L));
  base += (4u * warps * Lp < 24u * Lp) ? 24u * Lp : 4u * warps * Lp;
  m.it.a = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.b = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.e = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.g = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.sym = reinterpret_cast<float*>(base);                     base += 4u * (2 * Lp + 8);
  m.doc = reinterpret_cast<uint16_t*>(base);                  base += 2u * Lp;
  m.red = reinterpret_cast<float*>(base);                     base += 4u * 40;
  m.hist = reinterpret_cast<int*>(base);                      base += 4u * 40;
  m.stage = nullptr; m.bar = nullptr;
  if (tma_rel_bytes) {
    m.stage = base;   base += ring_align16(4u * L) + ring_align16(static_cast<size_t>(tma_rel_bytes) * L);
    m.bar = reinterpret_cast<uint64_t*>(base);
  }
  return m;
}

// ---- pair phase: this warp's share of the (row group, ring step) units ---------------------------------
// One ring step of one lane: the R rows it holds against the columns of chunk (c + m) mod C.
//   symc : centre of the signed-distance delta table (symc[d] = delta[|d|])
// FAST (warp-uniform): every lane of the group owns a chunk and the step is not the doubled last
// step of an even ring, so nothing is predicated and the rows accumulate in place.
template <int TW, bool FACTORED, bool FAST>
__device__ __forceinline__ void ring_step(const PairSoA& it, const float* __restrict__ colx,
                                          float* __restrict__ gw, const float* __restrict__ symc, int c,
                                          bool active, int m, int C, bool dup_step, const float (&ra)[4],
                                          const float (&re)[4], const float (&rg)[4], const float (&rv)[4],
                                          float (&racc)[4], float& lacc) {
  constexpr int R = 4;
  int pc = c + m;
  int sd = m;                                   // signed chunk distance column - row
  if (pc >= C) { pc -= C; sd = m - C; }         // wrapped: the column ranks above the row
  float4* g4 = reinterpret_cast<float4*>(gw) + pc;
  float4 gold = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  const bool commit = FAST || (active && !(dup_step && sd < 0));
  if (commit) gold = *g4;
  float dwin[8];
  if constexpr (TW == TW_DELTA) {
    // rank distance of row r and column k: |4 sd + (k - r)|, slot k - r + 3
    const float4* w4 = reinterpret_cast<const float4*>(symc + 4 * sd - 4);
    const float4 w0 = w4[0], w1 = w4[1];
    dwin[0] = w0.y; dwin[1] = w0.z; dwin[2] = w0.w;
    dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w; dwin[7] = 0.0f;
  }
  float cx[R], ce[R], cg[R];
  load_chunk<R>(colx, pc * R, cx);
  load_chunk<R>(it.g, pc * R, cg);
  if constexpr (FACTORED) load_chunk<R>(it.e, pc * R, ce);
  float tc[R];
#pragma unroll
  for (int k = 0; k < R; ++k) tc[k] = 0.0f;
  if constexpr (FAST) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float dw = rv[r];
        if constexpr (TW == TW_DELTA) dw = dwin[k - r + R - 1];
        pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[k], FACTORED ? ce[k] : 0.0f, cg[k], dw, lacc, racc[r],
                                tc[k]);
      }
    }
  } else {
    float tl = 0.0f, tr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) tr[r] = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float dw = rv[r];
        if constexpr (TW == TW_DELTA) dw = dwin[k - r + R - 1];
        pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[k], FACTORED ? ce[k] : 0.0f, cg[k], dw, tl, tr[r], tc[k]);
      }
    }
    if (commit) {
      lacc += tl;
#pragma unroll
      for (int r = 0; r < R; ++r) racc[r] += tr[r];
    }
  }
  if (commit) {
    gold.x += tc[0]; gold.y += tc[1]; gold.z += tc[2]; gold.w += tc[3];
    *g4 = gold;
  }
  __syncwarp();
}

// Factored form of ring_step: packed pair math (chunk_pairs, ltr_pair_tiles.cuh), accumulators in RowAcc.
template <int TW, bool FAST>
__device__ __forceinline__ void ring_step_fact(const PairSoA& it, float* __restrict__ gw,
                                               const float* __restrict__ symc, int c, bool active, int m, int C,
                                               bool dup_step, const float (&ra)[4], const float (&re)[4],
                                               const float (&rg)[4], const float (&rv)[4], RowAcc<4>& acc) {
  int pc = c + m;
  int sd = m;                                   // signed chunk distance column - row
  if (pc >= C) { pc -= C; sd = m - C; }         // wrapped: the column ranks above the row
  float4* g4 = reinterpret_cast<float4*>(gw) + pc;
  float dwin[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if constexpr (TW == TW_DELTA) {
    // rank distance of row r and column k: |4 sd + (k - r)|, slot k - r + 3
    const float4* w4 = reinterpret_cast<const float4*>(symc + 4 * sd - 4);
    const float4 w0 = w4[0], w1 = w4[1];
    dwin[0] = w0.y; dwin[1] = w0.z; dwin[2] = w0.w;
    dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w;
  }
  float cx[4], ce[4], cg[4];
  load_cols<4, 4>(it.b, pc, cx);
  load_cols<4, 4>(it.g, pc, cg);
  load_cols<4, 4>(it.e, pc, ce);
  if constexpr (FAST) {
    const float4 gold = *g4;
    float tc[4] = {gold.x, gold.y, gold.z, gold.w};
    chunk_pairs<TW, 4>(ra, re, rg, rv, cx, ce, cg, dwin, acc, tc);
    *g4 = make_float4(tc[0], tc[1], tc[2], tc[3]);
  } else {
    const bool commit = active && !(dup_step && sd < 0);
    RowAcc<4> tmp;
    tmp.clear();
    float tc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    chunk_pairs<TW, 4>(ra, re, rg, rv, cx, ce, cg, dwin, tmp, tc);
    if (commit) {
      acc.add(tmp);
      float4 gold = *g4;
      gold.x += tc[0]; gold.y += tc[1]; gold.z += tc[2]; gold.w += tc[3];
      *g4 = gold;
    }
  }
  __syncwarp();
}

//   it    : rank-ordered factors, padded to 4 C entries
//   gw    : THIS warp's private rank-order gradient array, zero on entry
//   symc  : centre of the signed-distance delta table in shared memory (TW_DELTA)
// Requires C = ceil(nb / 4) >= 32.  Returns the lane's partial loss.
template <int TW, bool FACTORED, int W>
__device__ __forceinline__ float ring_split(const PairSoA& it, float* __restrict__ gw,
                                            const float* __restrict__ symc, int nb, int lane, int warp) {
  constexpr int R = 4;
  const int C = (nb + R - 1) / R;
  const int G = (C + 31) >> 5;
  const int M = C >> 1;                 // ring steps 1 .. M; step 0 is the triangle inside a chunk
  // The last row group only has cl_n = C - 32 (G - 1) chunks.  When they fill at most half the warp,
  // f lanes share every chunk ("replicas") and split its M + 1 steps into f runs of T: replica s of
  // chunk c takes the steps s T .. s T + T - 1, so the group costs T instead of M + 1 steps.  All
  // lanes of one iteration still touch different column chunks as long as T >= cl_n (two lanes meet
  // the same column iff their chunk offsets differ by a multiple of T) and (f - 1) T + cl_n <= C (no
  // collision around the ring): f is halved until both hold.
  const int cl_n = C - ((G - 1) << 5);
  int f = 1;
  while (cl_n * f * 2 <= 32) f <<= 1;
  int T = (M + f) / f;                  // ceil((M + 1) / f)
  while (f > 1 && (T < cl_n || (f - 1) * T + cl_n > C)) { f >>= 1; T = (M + f) / f; }
  const int U_full = (G - 1) * (M + 1);
  const int U = U_full + T;             // f == 1: T == M + 1
  int u = static_cast<int>((static_cast<long long>(U) * warp) / W);
  const int u_end = static_cast<int>((static_cast<long long>(U) * (warp + 1)) / W);
  const bool even = (C & 1) == 0;
  const float* colx = FACTORED ? it.b : it.a;
  float lacc = 0.0f;

  while (u < u_end) {
    // ---- a run of steps [m0, m1) of row group g ------------------------------------------------------
    const bool last = u >= U_full;
    const int g = last ? G - 1 : u / (M + 1);
    int m0 = u - g * (M + 1);
    const int steps_g = (last && f > 1) ? T : M + 1;
    const int m1 = min(steps_g, m0 + (u_end - u));
    u += m1 - m0;
    const bool replicated = last && f > 1;             // warp-uniform
    // lane -> (chunk, replica); replica 0 of every chunk everywhere else
    const int sub = replicated ? lane / cl_n : 0;
    const int cl = replicated ? lane - sub * cl_n : lane;
    const int c = (g << 5) + cl;
    const bool active = replicated ? sub < f : c < C;
    const bool full = !replicated && ((g << 5) + 31) < C;   // warp-uniform
    const int me = active ? c : 0;                    // an idle lane mirrors chunk 0 and never commits
    float ra[R], re[R], rg[R], rv[R], racc[R];
    load_chunk<R>(it.a, me * R, ra);
    load_chunk<R>(it.g, me * R, rg);
    if constexpr (FACTORED) load_chunk<R>(it.e, me * R, re);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rv[r] = 1.0f;
      if constexpr (FACTORED) {
        const bool valid = active && (me * R + r < nb);
        ra[r] = valid ? ra[r] : 0.0f;
        rg[r] = valid ? rg[r] : (TW == TW_TWO ? 0.0f : kBigGain);
        rv[r] = valid ? 1.0f : 0.0f;
      } else {
        re[r] = 0.0f;
      }
      racc[r] = 0.0f;
    }
    RowAcc<R> acc;          // factored form: packed accumulators of this run (racc / run_l: triangle only)
    acc.clear();
    float run_l = 0.0f;     // this run's loss in pair_fact units (the factored winner losses accumulate loss / 2)
    if (m0 == 0) {
      // ---- triangle inside the chunk (rank distance k - r > 0); replica 0 only ----------------------------
      float dwin[8];
      if constexpr (TW == TW_DELTA) {
        const float4 w1 = *reinterpret_cast<const float4*>(symc);
        dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w;
        dwin[0] = dwin[1] = dwin[2] = dwin[7] = 0.0f;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dwin[k] = 0.0f;
      }
      float tl = 0.0f, tr[R];
#pragma unroll
      for (int r = 0; r < R; ++r) tr[r] = 0.0f;
      if constexpr (FACTORED) {
        float cx[4], ce[4], cg[4];
        load_cols<R, 4>(it.b, me, cx);
        load_cols<R, 4>(it.g, me, cg);
        load_cols<R, 4>(it.e, me, ce);
        chunk_triangle<TW, R>(ra, re, rg, rv, cx, ce, cg, dwin, tl, tr);
      } else {
        float cx[R], cg[R];
        load_chunk<R>(colx, me * R, cx);
        load_chunk<R>(it.g, me * R, cg);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int k = r + 1; k < R; ++k) {
            float dw = rv[r];
            if constexpr (TW == TW_DELTA) dw = dwin[k - r + R - 1];
            pair_once<TW, false>(ra[r], re[r], rg[r], cx[k], 0.0f, cg[k], dw, tl, tr[r], tr[k]);
          }
        }
      }
      if (active && sub == 0) {
        run_l += tl;
#pragma unroll
        for (int r = 0; r < R; ++r) racc[r] += tr[r];
      }
      if (!replicated) m0 = 1;
    }
    if (!replicated) {
      // even ring: the last step meets chunk c + C/2 from both ends; only the unwrapped end commits
      const int m_plain = (even && m1 == M + 1) ? M : m1;
      const int m_fast = full ? m_plain : m0;
      if constexpr (FACTORED) {
        for (int m = m0; m < m_fast; ++m)
          ring_step_fact<TW, true>(it, gw, symc, c, true, m, C, false, ra, re, rg, rv, acc);
        for (int m = max(m0, m_fast); m < m1; ++m)
          ring_step_fact<TW, false>(it, gw, symc, me, active, m, C, even && m == M, ra, re, rg, rv, acc);
      } else {
        for (int m = m0; m < m_fast; ++m)
          ring_step<TW, FACTORED, true>(it, colx, gw, symc, c, true, m, C, false, ra, re, rg, rv, racc, run_l);
        for (int m = max(m0, m_fast); m < m1; ++m)
          ring_step<TW, FACTORED, false>(it, colx, gw, symc, me, active, m, C, even && m == M, ra, re, rg, rv, racc,
                                         run_l);
      }
      // ---- flush the rows ----------------------------------------------------------------------------------
      if (active) {
        float4* g4 = reinterpret_cast<float4*>(gw) + c;
        float4 t = *g4;
        t.x += racc[0] + acc.row(0); t.y += racc[1] + acc.row(1);
        t.z += racc[2] + acc.row(2); t.w += racc[3] + acc.row(3);
        *g4 = t;
      }
      __syncwarp();
    } else {
      // replica `sub` of chunk c: iteration t is its ring step sub * T + t (step 0 is the triangle above)
      for (int t = m0; t < m1; ++t) {
        const int m = sub * T + t;
        const bool on = active && m >= 1 && m <= M;
        if constexpr (FACTORED)
          ring_step_fact<TW, false>(it, gw, symc, me, on, on ? m : 1, C, even && m == M, ra, re, rg, rv, acc);
        else
          ring_step<TW, FACTORED, false>(it, colx, gw, symc, me, on, on ? m : 1, C, even && m == M, ra, re, rg, rv,
                                         racc, run_l);
      }
      // the f replicas of a chunk add their row sums one after the other
      for (int s = 0; s < f; ++s) {
        if (active && sub == s) {
          float4* g4 = reinterpret_cast<float4*>(gw) + c;
          float4 tt = *g4;
          tt.x += racc[0] + acc.row(0); tt.y += racc[1] + acc.row(1);
          tt.z += racc[2] + acc.row(2); tt.w += racc[3] + acc.row(3);
          *g4 = tt;
        }
        __syncwarp();
      }
    }
    lacc += run_l + acc.loss();
  }
  return lacc;
}

// Short query (C < 32 chunks) inside a long-list batch: warp 0 runs the single-warp ring of
// ltr_pair_tiles.cuh and leaves the rank-order gradient in its private array.
template <int TW, bool FACTORED, int R>
__device__ __forceinline__ float ring_small(const PairSoA& it, float* __restrict__ gw, const PairTables& tb,
                                            int nb, int lane) {
  const int C = (nb + R - 1) / R;
  float racc[R];
  const float l = ring_pass<TW, FACTORED, R>(it, gw, tb.wtab[R - 1], C, nb, lane, racc);
  float tot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) tot[r] = lane < C ? gw[lane * 4 + r] + racc[r] : 0.0f;
  __syncwarp();
  *reinterpret_cast<float4*>(gw + lane * 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  __syncwarp();
  if (lane < C) {
#pragma unroll
    for (int r = 0; r < R; ++r) gw[lane * R + r] = tot[r];
  }
  return l;
}

template <int TW, bool FACTORED, int W>
__device__ __forceinline__ float ring_pairs(const RingSmem& m, const PairTables& tb, int Lp, int nb, int lane,
                                            int warp) {
  if (nb > 124) return ring_split<TW, FACTORED, W>(m.it, m.gw + warp * Lp, m.sym + Lp, nb, lane, warp);
  if (warp != 0) return 0.0f;
  const int R = (nb + 31) >> 5;
  if (R == 1) return ring_small<TW, FACTORED, 1>(m.it, m.gw, tb, nb, lane);
  if (R == 2) return ring_small<TW, FACTORED, 2>(m.it, m.gw, tb, nb, lane);
  if (R == 3) return ring_small<TW, FACTORED, 3>(m.it, m.gw, tb, nb, lane);
  return ring_small<TW, FACTORED, 4>(m.it, m.gw, tb, nb, lane);
}

template <int TW, int W>
__global__ void __launch_bounds__(W * 32, W == kRingWarpsLong ? LTR_RING_MIN_CTAS
                                          : (W == kRingWarpsShort ? LTR_RING_SHORT_CTAS
                                                                    : (W == 2 ? 2 * LTR_RING_SHORT_CTAS : 16)))
pair_ring_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                 const void* __restrict__ n, int n_bytes, int B, int L, int P, float sigma, int variant,
                 int tma, float* __restrict__ loss_out, float* __restrict__ grad_out,
                 int64_t* __restrict__ ranking_out, float* __restrict__ loss_sum,
                 unsigned int* __restrict__ queue, const unsigned int* __restrict__ order,
                 const PairTables* __restrict__ tabs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RingSmem m = ring_carve(smem_raw, L, P, tma ? rel_bytes : 0, W);
  const PairTables& tb = *tabs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Lp = (L + 127) / 128 * 128;
  const float gscale = sigma * kLog2e;
  const double kd = static_cast<double>(sigma) * 1.4426950408889634;
  const float k_hi = static_cast<float>(kd);
  const float k_lo = static_cast<float>(kd - static_cast<double>(k_hi));

  const bool rank_weighted = TW == TW_DELTA || (TW == TW_TWO && variant != 0);   // NDCG losses
  const bool sorted = rank_weighted || ranking_out != nullptr;
  if constexpr (TW == TW_DELTA) {
    for (int k = threadIdx.x; k < 2 * Lp + 8; k += blockDim.x) m.sym[k] = tb.delta[k < Lp ? Lp - k : k - Lp];
  }

  const uint32_t row_s_bytes = 4u * L, row_y_bytes = static_cast<uint32_t>(rel_bytes) * L;
  const float* stage_s = reinterpret_cast<const float*>(m.stage);
  const unsigned char* stage_y = m.stage + ring_align16(row_s_bytes);
  auto issue_row = [&](int q) {
    mbar_arrive_expect_tx(m.bar, row_s_bytes + row_y_bytes);
    tma_load_1d(m.stage, scores + static_cast<size_t>(q) * L, row_s_bytes, m.bar);
    tma_load_1d(m.stage + ring_align16(row_s_bytes),
                static_cast<const unsigned char*>(rel) + static_cast<size_t>(q) * row_y_bytes, row_y_bytes, m.bar);
  };
  if (tma) {
    if (threadIdx.x == 0) {
      mbar_init(m.bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
  }

  // ---- query schedule: the first query of every CTA is static, the following ones come from a
  // device-wide queue.  `order` (optional) lists the queries by decreasing size: the long ones
  // start first and the tail of the launch is made of short ones -----------------------------------
  // queue == nullptr (launch captured into a CUDA graph without a caller-owned workspace): static stride
  const bool dynamic = queue != nullptr && gridDim.x < static_cast<unsigned int>(B);
  int b = static_cast<int>(blockIdx.x) < B ? static_cast<int>(order ? order[blockIdx.x] : blockIdx.x) : -1;
  if (tma && threadIdx.x == 0 && b >= 0) issue_row(b);

  for (int iter = 0; b >= 0; ++iter) {
    __syncthreads();   // previous query fully consumed (gw / keys, raw_s, red, hist)
    if (threadIdx.x == 0) {
      int b_next = -1;
      if (dynamic) {
        const unsigned int qn = gridDim.x + atomicAdd(queue, 1u);
        if (qn < static_cast<unsigned int>(B)) b_next = static_cast<int>(order ? order[qn] : qn);
      } else {
        const unsigned int qn = blockIdx.x + (iter + 1) * gridDim.x;
        if (qn < static_cast<unsigned int>(B)) b_next = static_cast<int>(order ? order[qn] : qn);
      }
      m.hist[38] = b_next;
    }
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;

    // ---- unpack the row: scores, grades, sort keys ---------------------------------------------------
    if (tma) mbar_wait(m.bar, iter & 1);
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
      uint64_t key = ~0ull;
      if (j < L) {
        const float s = tma ? stage_s[j] : scores[base + j];
        const int y = tma ? load_int_clamped(stage_y, rel_bytes, j) : load_int_clamped(rel, rel_bytes, base + j);
        m.raw_s[j] = s;
        m.raw_y[j] = y;
        key = sorted ? pack_key(j < nb ? desc_key_f32(s) : kPadKey, j) : static_cast<uint64_t>(j);
      }
      m.keys[j] = key;
    }
    if (threadIdx.x < 36) m.hist[threadIdx.x] = 0;
    __syncthreads();
    const int b_next = m.hist[38];
    if (tma && threadIdx.x == 0 && b_next >= 0) {
      fence_proxy_async();   // the staging buffer was just read through the generic proxy
      issue_row(b_next);
    }

    // ---- rank_by_score ----------------------------------------------------------------------------------
    if (sorted) cta_block_sort(m.keys, P, lane, warp, W);

    // ---- ideal DCG ---------------------------------------------------------------------------------------
    float max_dcg = 1.0f;
    if (rank_weighted) {
      for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int y = m.raw_y[j];
        if (y < 0 || y > 31) m.hist[32] = 1;
        else if (y > 0) atomicAdd(&m.hist[y], 1);
      }
      __syncthreads();
      if (m.hist[32] == 0) {
        if (warp == 0) {
          // grade g (lane g) occupies the ideal ranks [start, start + cnt), start = count of higher grades
          const int cnt = lane >= 1 ? m.hist[lane] : 0;
          int above = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, above, o);
            if (lane + o < 32) above += t;
          }
          const int start = above - cnt;
          double part = 0.0;
          if (cnt > 0)
            part = static_cast<double>(gain_of_grade(lane)) * (tb.inv_disc_prefix[start + cnt] - tb.inv_disc_prefix[start]);
          // fixed order: descending grade, like the serial loop of the warp kernel
          double acc = 0.0;
          for (int gsel = 31; gsel >= 1; --gsel) {
            const double v = __shfl_sync(0xffffffffu, part, gsel);
            if (__shfl_sync(0xffffffffu, cnt, gsel) > 0) acc += v;
          }
          if (lane == 0) m.red[36] = static_cast<float>(acc);
        }
        __syncthreads();
        max_dcg = m.red[36];
      } else {
        // grades outside [0, 31]: sort them (the score ranking is saved in m.doc first)
        for (int r = threadIdx.x; r < L; r += blockDim.x) m.doc[r] = static_cast<int>(m.keys[r] & 0xffffffffu);
        __syncthreads();
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
          uint32_t key = kPadKey;
          if (j < nb) key = desc_key_i32(m.raw_y[j]);
          m.keys[j] = j < L ? pack_key(key, j) : ~0ull;
        }
        cta_block_sort(m.keys, P, lane, warp, W);
        float part = 0.0f;
        for (int r = threadIdx.x; r < nb; r += blockDim.x)
          part += exp_gain_f32(m.raw_y[static_cast<int>(m.keys[r] & 0xffffffffu)]) / tb.disc[r];
        max_dcg = cta_sum(part, m.red);
        for (int r = threadIdx.x; r < L; r += blockDim.x) m.keys[r] = static_cast<uint64_t>(m.doc[r]);
        __syncthreads();
      }
      if (max_dcg == 0.0f) max_dcg = 1.0f;
    }
    const float inv_max_dcg = 1.0f / max_dcg;

    // ---- score range, per-document factors in rank order ----------------------------------------------
    float smax = 0.0f, smin = 0.0f;
    if (sorted) {
      if (nb > 0) {
        smax = m.raw_s[static_cast<int>(m.keys[0] & 0xffffffffu)];
        smin = m.raw_s[static_cast<int>(m.keys[nb - 1] & 0xffffffffu)];
      }
    } else if constexpr (TW != TW_HINGE) {
      float lmax = -INFINITY, lmin = INFINITY;
      for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        lmax = fmaxf(lmax, m.raw_s[j]);
        lmin = fminf(lmin, m.raw_s[j]);
      }
      lmax = warp_max(lmax);
      lmin = -warp_max(-lmin);
      if (lane == 0) { m.red[warp] = lmax; m.red[8 + warp] = lmin; }
      __syncthreads();
      smax = m.red[0]; smin = m.red[8];
      for (int w = 1; w < W; ++w) { smax = fmaxf(smax, m.red[w]); smin = fminf(smin, m.red[8 + w]); }
    }
    const float mid = 0.5f * (smax + smin);
    const bool factored = TW != TW_HINGE && fabsf(sigma) * (smax - smin) * kLog2e <= kFactoredRange;
    // winner-by-relevance losses: padded columns carry the smallest valid weight, so that they lose (or tie)
    // every pair, in the factored and in the stable form.  Grades 0..31 give gains >= 0: the padding value 0 is small enough.
    float gpad = 0.0f;
    if constexpr (tw_winner(TW)) {
      if (!(TW == TW_DELTA && m.hist[32] == 0)) {
        float lmin = INFINITY;
        for (int j = threadIdx.x; j < nb; j += blockDim.x) {
          const int y = m.raw_y[j];
          lmin = fminf(lmin, TW == TW_DELTA ? gain_of_grade(y) * fabsf(inv_max_dcg) : static_cast<float>(y));
        }
        gpad = cta_min(lmin, m.red);
        if (!(gpad < INFINITY)) gpad = 0.0f;
      }
    }
    float diag = 0.0f;
    for (int p = threadIdx.x; p < Lp; p += blockDim.x) {
      float fa = factored ? 0.0f : -1.0e30f, fb = 0.0f, fe = 0.0f;                // padding
      float fg = TW == TW_HINGE ? -1.0e30f : gpad;
      int d = p;
      if (p < L) {
        d = static_cast<int>(m.keys[p] & 0xffffffffu);
        if (ranking_out) ranking_out[base + p] = d;
      }
      if (p < nb) {
        const float s = m.raw_s[d];
        const int y = m.raw_y[d];
        float w;
        if constexpr (TW == TW_DELTA) w = gain_of_grade(y) * fabsf(inv_max_dcg);   // |G_i - G_j|: :214-216
        else if (TW == TW_TWO && variant != 0) w = gain_of_grade(y) * inv_max_dcg / tb.disc[p];
        else w = static_cast<float>(y);
        if constexpr (TW == TW_TWO) diag += w;   // the pairs (i, i): w_i * log2(1 + e^0)
        fg = w;
        if constexpr (TW == TW_HINGE) {
          fa = s;                                // raw score: the hinge works on s_i - s_j itself
        } else if (factored) {
          doc_factors<TW>(s, mid, k_hi, k_lo, w, fa, fb, fe, fg);
        } else {
          fa = sigma * s;
        }
      }
      m.it.a[p] = fa; m.it.b[p] = fb; m.it.e[p] = fe; m.it.g[p] = fg;
      m.doc[p] = static_cast<uint16_t>(d);
    }
    __syncthreads();   // keys and row fully consumed: their storage becomes the private gradient arrays
    {
      float4* z = reinterpret_cast<float4*>(m.gw);
      const int nz = (nb > 124 ? W : 1) * (Lp >> 2);
      for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();

    // ---- all pairs, once ---------------------------------------------------------------------------------
    float wl = 0.0f;
    if (nb > 1) {
      if constexpr (TW == TW_HINGE) wl = ring_pairs<TW, false, W>(m, tb, Lp, nb, lane, warp);
      else wl = factored ? ring_pairs<TW, true, W>(m, tb, Lp, nb, lane, warp) : ring_pairs<TW, false, W>(m, tb, Lp, nb, lane, warp);
    }
    float loss = cta_sum(wl + diag, m.red);   // its barriers also publish the private arrays
    float gmul = gscale;
    if constexpr (TW == TW_HINGE) {
      gmul = 1.0f;
      if (variant) {
        // pairwise_additive.py:132-133: -1 / ln(2 + h); d/dh = 1 / ((2 + h) ln^2(2 + h))
        const float lg = logf(2.0f + loss);
        gmul = 1.0f / ((2.0f + loss) * lg * lg);
        loss = -1.0f / lg;
      }
    }
    if (threadIdx.x == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }

    // ---- gradient: sum of the private arrays, back to document order -----------------------------------
    if (grad_out) {
      float* gdoc = m.it.a;   // the factors are dead
      const int copies = nb > 124 ? W : 1;
      for (int p = threadIdx.x; p < L; p += blockDim.x) {
        float gsum = 0.0f;
        if (p < nb && nb > 1) {
          for (int w = 0; w < copies; ++w) gsum += m.gw[w * Lp + p];
        }
        gdoc[m.doc[p]] = gsum * gmul;
      }
      __syncthreads();
      float* __restrict__ go = grad_out + base;
      for (int j = threadIdx.x; j < L; j += blockDim.x) go[j] = gdoc[j];
    }
    b = b_next;
  }

  // ---- leave the queue clean for the next launch that uses this slot ----------------------------------
  if (dynamic && threadIdx.x == 0) {
    const unsigned int done = atomicAdd(queue + 1, 1u);
    if (done == gridDim.x - 1) {
      queue[0] = 0u;
      queue[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace ltr
