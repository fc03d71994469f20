// ltr_pair_warp.cuh -- one WARP per query for list sizes up to 128 (the headline (4096, 128)
// workload): no CTA barrier in the steady state, everything of a query lives in the warp's
// registers plus ~4 KB of shared memory.
//
// Per query: 128-bit loads of the padded row -> in-register bitonic argsort
// (rank_by_score, utils/tensor_operations.py:48-64) -> ideal DCG from a grade histogram
// (_max_dcg, pairwise_lambda.py:231-241) -> per-document factors into shared memory in rank
// order -> ring_pass (ltr_pair_tiles.cuh) -> gradient scattered back to document order
// (backward of the gather, pairwise_lambda.py:69) and stored with 128-bit writes.
//
// Queries are handed out dynamically (one atomic per query on a self-resetting device
// counter): list sizes vary by 2x, pair counts by 4x, so a static split leaves most of the SMs
// idle in the tail.
#pragma once

#include "ltr_pair_tiles.cuh"
#include "ltr_sm100.h"

namespace ltr {

constexpr int kWarpL = 128;          // max list size of the warp-per-query kernel
constexpr int kWarpE = 4;            // elements per lane in the sort (32 * 4 = 128)
constexpr int kWarpsPerCta = 4;
constexpr int kWarpSchedCtas = 0;    // cap on resident CTAs per SM under a longest-first schedule (0 = none)

// direction predicates of the bitonic network for element index lane * E + r
template <int E>
__device__ __forceinline__ bool bitonic_up(int lane, int r, int size) {
  return ((lane * E + r) & size) == 0;
}

// ---- warp-wide bitonic sort of 32*E 32-bit keys, element index = lane * E + r, ascending ----
template <int E>
__device__ __forceinline__ void warp_bitonic_sort32(uint32_t (&k)[E], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride < E) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int q = r ^ stride;
          if (q > r) {
            // one predicated min/max per output (VIMNMX takes the min/max choice as a predicate)
            const bool up = bitonic_up<E>(lane, r, size);
            const uint32_t a = k[r], b = k[q];
            k[r] = up ? min(a, b) : max(a, b);
            k[q] = up ? max(a, b) : min(a, b);
          }
        }
      } else {
        const int ls = stride / E;
        // for stride >= E the direction does not depend on r: (lane * E) & size
        const bool keep_min = (((lane * E) & size) == 0) == ((lane & ls) == 0);
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, k[r], ls);
          k[r] = keep_min ? min(k[r], other) : max(k[r], other);
        }
      }
    }
  }
}

// ---- same network on 64-bit keys: exact (score key, index) order, used when the packed
// 32-bit sort cannot separate two scores --------------------------------------------------------
template <int E>
__device__ __forceinline__ void warp_bitonic_sort64(uint64_t (&k)[E], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride < E) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int q = r ^ stride;
          if (q > r) {
            const bool up = bitonic_up<E>(lane, r, size);
            const uint64_t a = k[r], b = k[q];
            const bool sw = (a > b) == up;
            k[r] = sw ? b : a;
            k[q] = sw ? a : b;
          }
        }
      } else {
        const int ls = stride / E;
        const bool keep_min = (((lane * E) & size) == 0) == ((lane & ls) == 0);
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

struct __align__(16) WarpScratch {
  __align__(16) float fa[kWarpL];          // per-document factors, rank order (PairSoA)
  __align__(16) float fb[kWarpL];
  __align__(16) float fe[kWarpL];
  __align__(16) float fg[kWarpL];
  __align__(16) float gcol[2 * kWarpL];    // column gradients: chunk c owns gcol[4c .. 4c+3]; the upper half is
                                           // the dump of lanes without a chunk (ring_pass) and, before
                                           // the pair phase, the rank -> document map
  __align__(16) float raw_s[kWarpL];       // scores, document order; reused: rank-order gradient
  __align__(16) int raw_y[kWarpL];         // relevance, document order; reused: document-order gradient
};

// Score-independent tables, computed once per device by init_pair_tables_kernel and read
// through the read-only path (they stay L1-resident: ~10 KB are touched by the warp kernel).
struct __align__(16) PairTables {
  __align__(16) float wtab[4][window_table_floats()];   // delta windows for R = 1..4 (float4 loads)
  double inv_disc_prefix[LTR_MAX_LIST_SIZE + 1];        // S[p] = sum_{r<p} 1 / D(r)  (ideal DCG)
  __align__(16) float delta[LTR_MAX_LIST_SIZE + 8];     // delta[k] = |1/D(k) - 1/D(k+1)|, pairwise_lambda.py:206-211
  __align__(16) float disc[LTR_MAX_LIST_SIZE + 8];      // D(r) = log2(2 + r), pairwise_lambda.py:168-170
  __align__(16) float inv_disc[LTR_MAX_LIST_SIZE + 8];  // 1 / D(r)
};

__global__ void __launch_bounds__(1024) init_pair_tables_kernel(PairTables* __restrict__ t) {
  for (int k = threadIdx.x; k < LTR_MAX_LIST_SIZE + 8; k += blockDim.x) {
    const float d0 = log2f(2.0f + static_cast<float>(k));
    const float d1 = log2f(3.0f + static_cast<float>(k));
    t->disc[k] = d0;
    t->inv_disc[k] = 1.0f / d0;
    t->delta[k] = fabsf(1.0f / d0 - 1.0f / d1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * window_table_floats(); i += blockDim.x) {
    const int R = i / window_table_floats() + 1, rem = i % window_table_floats();
    const int d = (rem >> 3) - kMaxChunks, slot = rem & 7;
    int k = R * d + slot - (R - 1);
    k = k < 0 ? -k : k;
    t->wtab[R - 1][rem] = t->delta[k];
  }
  // S[p] = sum_{r<p} 1 / D(r) in double: every thread owns kPer consecutive entries, the thread totals are
  // scanned through shared memory
  constexpr int kPer = (LTR_MAX_LIST_SIZE + 1 + 1023) / 1024;
  __shared__ double tot[1024];
  const int first = threadIdx.x * kPer;
  double local[kPer];
  double run = 0.0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    local[j] = run;                                   // exclusive within the thread
    const int r = first + j;
    if (r <= LTR_MAX_LIST_SIZE) run += 1.0 / static_cast<double>(t->disc[r]);
  }
  tot[threadIdx.x] = run;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const double add = threadIdx.x >= off ? tot[threadIdx.x - off] : 0.0;
    __syncthreads();
    tot[threadIdx.x] += add;
    __syncthreads();
  }
  const double base = threadIdx.x > 0 ? tot[threadIdx.x - 1] : 0.0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int r = first + j;
    if (r <= LTR_MAX_LIST_SIZE) t->inv_disc_prefix[r] = base + local[j];
  }
}

template <int TW, bool FACTORED, int R>
__device__ __forceinline__ float ring_dispatch(WarpScratch& ws, const PairTables& tb, int n, int lane) {
  const int C = (n + R - 1) / R;
  float racc[R];
  const PairSoA it = {ws.fa, ws.fb, ws.fe, ws.fg};
  const float l = ring_pass<TW, FACTORED, R, 4>(it, ws.gcol, tb.wtab[R - 1], C, n, lane, racc);
  // rank-order gradient (unscaled): column part + row part
  float* glin = ws.raw_s;
  if (lane < C) {
#pragma unroll
    for (int r = 0; r < R; ++r) glin[lane * R + r] = ws.gcol[lane * 4 + r] + racc[r];
  }
  return l;
}

__device__ __forceinline__ int clamp_i64_to_i32(long long v) {
  const int lo = static_cast<int>(v), hi = static_cast<int>(v >> 32);
  return hi == (lo >> 31) ? lo : (hi < 0 ? -2147483647 : 2147483647);
}

// 2^y - 1 in float32 for an integer grade (exact for |y| <= 126, like the reference's 2 ** rel - 1)
__device__ __forceinline__ float gain_of_grade(int y) {
  y = y < -126 ? -126 : (y > 127 ? 127 : y);
  return __int_as_float((y + 127) << 23) - 1.0f;
}

template <int TW>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8)
pair_warp_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                 const void* __restrict__ n, int n_bytes, int B, int L, float sigma, int vec_ok,
                 int variant, float* __restrict__ loss_out, float* __restrict__ grad_out,
                 int64_t* __restrict__ ranking_out, float* __restrict__ loss_sum,
                 unsigned int* __restrict__ queue, const unsigned int* __restrict__ order,
                 const PairTables* __restrict__ tabs) {
  const PairTables& tb = *tabs;
  __shared__ WarpScratch scratch[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  WarpScratch& ws = scratch[warp];
  const float gscale = sigma * kLog2e;   // lambda = sigma / ln 2 * w * sigmoid(-x)
  // sigma * log2(e) split into two floats: e = c * k exactly to ~2^-48 relative
  const double kd = static_cast<double>(sigma) * 1.4426950408889634;
  const float k_hi = static_cast<float>(kd);
  const float k_lo = static_cast<float>(kd - static_cast<double>(k_hi));

  // ---- query schedule: the first query of every warp is static (no start-up burst on the queue
  // counter), the following ones are pulled from the device-wide queue as warps finish.  `order`
  // (optional) lists the queries by decreasing size, so the long ones start first and the tail of
  // the launch is made of short ones (longest-processing-time-first list scheduling) --------------
  const unsigned int total_warps = gridDim.x * kWarpsPerCta;
  // queue == nullptr (launch captured into a CUDA graph without a caller-owned workspace): static stride
  const bool dynamic = queue != nullptr && total_warps < static_cast<unsigned int>(B);   // else no queue traffic
  unsigned int q = blockIdx.x * kWarpsPerCta + warp;

  while (q < static_cast<unsigned int>(B)) {
    const unsigned int b = order ? order[q] : q;
    unsigned int b_next = 0xffffffffu;
    if (dynamic && lane == 0) b_next = total_warps + atomicAdd(queue, 1u);   // consumed at the end of this iteration

    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;

    // ---- load 4 consecutive documents per lane -------------------------------------------
    float sv[kWarpE];
    int yv[kWarpE];
    if (vec_ok && lane * kWarpE < L) {
      const float4 s4 = *reinterpret_cast<const float4*>(scores + base + lane * kWarpE);
      sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w;
      if (rel_bytes == 8) {
        const longlong2* rp = reinterpret_cast<const longlong2*>(
            reinterpret_cast<const long long*>(rel) + base + lane * kWarpE);
        const longlong2 r0 = rp[0], r1 = rp[1];
        yv[0] = clamp_i64_to_i32(r0.x); yv[1] = clamp_i64_to_i32(r0.y);
        yv[2] = clamp_i64_to_i32(r1.x); yv[3] = clamp_i64_to_i32(r1.y);
      } else if (rel_bytes == 4) {
        const int4 r4 = *reinterpret_cast<const int4*>(reinterpret_cast<const int*>(rel) + base + lane * kWarpE);
        yv[0] = r4.x; yv[1] = r4.y; yv[2] = r4.z; yv[3] = r4.w;
      } else if (rel_bytes == 2) {
        const short4 r4 = *reinterpret_cast<const short4*>(reinterpret_cast<const short*>(rel) + base + lane * kWarpE);
        yv[0] = r4.x; yv[1] = r4.y; yv[2] = r4.z; yv[3] = r4.w;
      } else {
        const uchar4 r4 =
            *reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(rel) + base + lane * kWarpE);
        yv[0] = r4.x; yv[1] = r4.y; yv[2] = r4.z; yv[3] = r4.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const int j = lane * kWarpE + r;
        sv[r] = j < L ? scores[base + j] : 0.0f;
        yv[r] = j < L ? load_int_clamped(rel, rel_bytes, base + j) : 0;
      }
    }
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) {
      ws.raw_s[lane * kWarpE + r] = sv[r];
      ws.raw_y[lane * kWarpE + r] = yv[r];
    }

    // ---- argsort by descending score; padding (and the slots beyond L) last, by index ------
    // Only LambdaNDCGLoss2 depends on the rank positions; the other losses are permutation
    // invariant and stay in document order unless the caller asked for the ranking.
    // Fast path: 25 key bits + 7 index bits in one register.  The result is then checked
    // against the exact (32-bit key, index) order; if two scores are closer than 128 ulps and
    // came out in the wrong order, the exact 64-bit network is run instead.
    int doc[kWarpE];
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) doc[r] = lane * kWarpE + r;
    const bool rank_weighted = TW == TW_DELTA || (TW == TW_TWO && variant != 0);   // NDCG losses
    if (rank_weighted || ranking_out != nullptr) {
      uint32_t ekey[kWarpE];   // exact keys (document order for now)
      uint32_t pk[kWarpE];
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const int j = lane * kWarpE + r;
        ekey[r] = j < nb ? desc_key_f32(sv[r]) : kPadKey;
        pk[r] = (ekey[r] & 0xffffff80u) | static_cast<uint32_t>(j);
      }
      warp_bitonic_sort32<kWarpE>(pk, lane);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) doc[r] = static_cast<int>(pk[r] & 127u);
      // The packed order is exact unless two neighbouring valid keys share their 25 key bits.
      const uint32_t pk_next = __shfl_down_sync(0xffffffffu, pk[0], 1);
      bool ambiguous = false;
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const uint32_t nxt = r + 1 < kWarpE ? pk[(r + 1) % kWarpE] : pk_next;
        ambiguous = ambiguous || (((pk[r] ^ nxt) < 128u) && (lane * kWarpE + r + 1 < nb));
      }
      if (__any_sync(0xffffffffu, ambiguous)) {
        // compare against the exact (32-bit key, index) order; redo with the 64-bit network if
        // the packed sort put two close scores the wrong way round
        uint64_t xk[kWarpE];
#pragma unroll
        for (int r = 0; r < kWarpE; ++r) {
          const int p = lane * kWarpE + r;
          xk[r] = pack_key(p < nb ? desc_key_f32(ws.raw_s[doc[r]]) : kPadKey, doc[r]);
        }
        const uint64_t next0 = __shfl_down_sync(0xffffffffu, xk[0], 1);
        const bool bad = (xk[0] > xk[1]) || (xk[1] > xk[2]) || (xk[2] > xk[3]) || (lane < 31 && xk[3] > next0);
        if (__any_sync(0xffffffffu, bad)) {
#pragma unroll
          for (int r = 0; r < kWarpE; ++r) xk[r] = pack_key(ekey[r], lane * kWarpE + r);
          warp_bitonic_sort64<kWarpE>(xk, lane);
#pragma unroll
          for (int r = 0; r < kWarpE; ++r) doc[r] = static_cast<int>(xk[r] & 0xffffffffu);
        }
      }
    }
    __syncwarp();

    // ---- ideal DCG over the valid documents --------------------------------------------------------
    float max_dcg = 1.0f;
    if (rank_weighted) {
      // Fast path, grades 0..7: a packed 8 x 8-bit histogram summed across the warp.
      bool wide = false;
      unsigned long long hist = 0ull;
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const bool valid = lane * kWarpE + r < nb;
        wide = wide || (valid && static_cast<unsigned int>(yv[r]) > 7u);
        if (valid) hist += 1ull << (8 * (yv[r] & 7));
      }
      const bool small_grades = !__any_sync(0xffffffffu, wide);
      int ymax = -2147483647, ymin = 2147483647;
      if (small_grades) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hist += __shfl_xor_sync(0xffffffffu, hist, o);   // counts <= 128 < 256
        double acc = 0.0;
        int start = 0;
#pragma unroll
        for (int g = 7; g >= 1; --g) {
          const int cnt = static_cast<int>((hist >> (8 * g)) & 0xffull);
          if (cnt > 0) {
            acc += static_cast<double>(gain_of_grade(g)) *
                   (tb.inv_disc_prefix[start + cnt] - tb.inv_disc_prefix[start]);
            start += cnt;
          }
        }
        max_dcg = static_cast<float>(acc);
      } else {
#pragma unroll
        for (int r = 0; r < kWarpE; ++r) {
          if (lane * kWarpE + r < nb) { ymax = max(ymax, yv[r]); ymin = min(ymin, yv[r]); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
          ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        }
      }
      if (small_grades) {
        // done above
      } else
      if (ymin >= 0 && ymax < 32) {
        // grade histogram by ballots; grade g occupies ideal ranks [start, start + cnt)
        double acc = 0.0;
        int start = 0;
        for (int g = ymax; g >= 1; --g) {
          int cnt = 0;
#pragma unroll
          for (int r = 0; r < kWarpE; ++r)
            cnt += __popc(__ballot_sync(0xffffffffu, lane * kWarpE + r < nb && yv[r] == g));
          acc += static_cast<double>(gain_of_grade(g)) *
                 (tb.inv_disc_prefix[start + cnt] - tb.inv_disc_prefix[start]);
          start += cnt;
        }
        max_dcg = static_cast<float>(acc);
      } else {
        uint32_t yk[kWarpE];
#pragma unroll
        for (int r = 0; r < kWarpE; ++r) yk[r] = lane * kWarpE + r < nb ? desc_key_i32(yv[r]) : kPadKey;
        warp_bitonic_sort32<kWarpE>(yk, lane);
        float part = 0.0f;
#pragma unroll
        for (int r = 0; r < kWarpE; ++r) {
          const int p = lane * kWarpE + r;
          if (p < nb) part += exp_gain_f32(static_cast<int>(~yk[r] ^ 0x80000000u)) / tb.disc[p];
        }
        max_dcg = warp_sum(part);
      }
      if (max_dcg == 0.0f) max_dcg = 1.0f;
    }
    const float inv_max_dcg = 1.0f / max_dcg;

    // ---- score range; rank -> document map for the chunk owners -------------------------------------
    float smax = -INFINITY, smin = INFINITY;
    int* rank2doc = reinterpret_cast<int*>(ws.gcol + kWarpL);
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) {
      const int p = lane * kWarpE + r;
      if (p < nb) { smax = fmaxf(smax, sv[r]); smin = fminf(smin, sv[r]); }   // document order: same extremes
      if (ranking_out && p < L) ranking_out[base + p] = doc[r];
    }
    *reinterpret_cast<int4*>(rank2doc + lane * kWarpE) = make_int4(doc[0], doc[1], doc[2], doc[3]);
    smax = warp_max(smax);
    smin = -warp_max(-smin);
    const float mid = 0.5f * (smax + smin);
    // NaN / inf scores fail this test and take the stable form
    const bool factored = TW != TW_HINGE && fabsf(sigma) * (smax - smin) * kLog2e <= kFactoredRange;
    // chunk size of the ring: R consecutive ranks per lane, the densest chunking that fits 32 lanes
    // (the stable sigmoid form always takes R = 4)
    const int Rq = (TW != TW_HINGE && !factored) ? 4 : max(1, (nb + 31) >> 5);
    __syncwarp();

    // ---- per-document factors -> shared memory: lane c writes chunk c = ranks [R c, R c + R) into the
    // 4 slots [4 c, 4 c + 4) (unused slots and ranks >= n carry the padding values) -------------------
    float diag = 0.0f;   // ARP1 / NDCG1 also count the pairs (i, i): w_i * log2(1 + e^0) = w_i
    {
      float fa[kWarpE], fb[kWarpE], fe[kWarpE], fg[kWarpE];
      float gmin = INFINITY;   // winner-by-relevance losses: padded columns carry the smallest valid weight
#pragma unroll
      for (int j = 0; j < kWarpE; ++j) {
        const int p = lane * Rq + j;
        fa[j] = factored ? 0.0f : -1.0e30f;   // padding
        fb[j] = 0.0f; fe[j] = 0.0f; fg[j] = TW == TW_HINGE ? -1.0e30f : 0.0f;
        if (j < Rq && p < nb) {
          const int d = rank2doc[p];
          const float sd = ws.raw_s[d];
          const int yd = ws.raw_y[d];
          float w;
          if constexpr (TW == TW_DELTA) w = gain_of_grade(yd) * fabsf(inv_max_dcg);   // |G_i - G_j|: :214-216
          else if (TW == TW_TWO && variant != 0) w = gain_of_grade(yd) * inv_max_dcg / tb.disc[p];
          else w = static_cast<float>(yd);
          if constexpr (TW == TW_TWO) diag += w;
          fg[j] = w;
          if constexpr (TW == TW_HINGE) {
            fa[j] = sd;                          // raw score: the hinge works on s_i - s_j itself
          } else if (factored) {
            doc_factors<TW>(sd, mid, k_hi, k_lo, w, fa[j], fb[j], fe[j], fg[j]);
            gmin = fminf(gmin, w);
          } else {
            fa[j] = sigma * sd;
            gmin = fminf(gmin, w);
          }
        }
      }
      if constexpr (tw_winner(TW)) {
        // padded columns carry the smallest valid weight: they lose (or tie) every pair, in both forms
        gmin = -warp_max(-gmin);
        if (!(gmin < INFINITY)) gmin = 0.0f;
#pragma unroll
        for (int j = 0; j < kWarpE; ++j)
          if (!(j < Rq && lane * Rq + j < nb)) fg[j] = gmin;
      }
      __syncwarp();   // raw_s / raw_y / rank2doc fully consumed: they are reused below
      const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      reinterpret_cast<float4*>(ws.fa)[lane] = make_float4(fa[0], fa[1], fa[2], fa[3]);
      reinterpret_cast<float4*>(ws.fb)[lane] = make_float4(fb[0], fb[1], fb[2], fb[3]);
      reinterpret_cast<float4*>(ws.fe)[lane] = make_float4(fe[0], fe[1], fe[2], fe[3]);
      reinterpret_cast<float4*>(ws.fg)[lane] = make_float4(fg[0], fg[1], fg[2], fg[3]);
      reinterpret_cast<float4*>(ws.gcol)[lane] = zero;
      reinterpret_cast<float4*>(ws.raw_s)[lane] = zero;   // rank-order gradient of ranks no ring lane owns
    }
    __syncwarp();

    // ---- all pairs, once ----------------------------------------------------------------------------
    float lacc = 0.0f;
    if (nb > 1) {
      if (factored) {
        if (Rq == 1) lacc = ring_dispatch<TW, true, 1>(ws, tb, nb, lane);
        else if (Rq == 2) lacc = ring_dispatch<TW, true, 2>(ws, tb, nb, lane);
        else if (Rq == 3) lacc = ring_dispatch<TW, true, 3>(ws, tb, nb, lane);
        else lacc = ring_dispatch<TW, true, 4>(ws, tb, nb, lane);
      } else if constexpr (TW == TW_HINGE) {
        if (Rq == 1) lacc = ring_dispatch<TW, false, 1>(ws, tb, nb, lane);
        else if (Rq == 2) lacc = ring_dispatch<TW, false, 2>(ws, tb, nb, lane);
        else if (Rq == 3) lacc = ring_dispatch<TW, false, 3>(ws, tb, nb, lane);
        else lacc = ring_dispatch<TW, false, 4>(ws, tb, nb, lane);
      } else {
        lacc = ring_dispatch<TW, false, 4>(ws, tb, nb, lane);
      }
    }
    __syncwarp();
    float loss = warp_sum(lacc + diag);
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
    if (lane == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }

    // ---- gradient back to document order -----------------------------------------------------------
    if (grad_out) {
      float* gdoc = reinterpret_cast<float*>(ws.raw_y);
      const float4 g4 = *reinterpret_cast<const float4*>(ws.raw_s + lane * kWarpE);
      const float gl[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const int p = lane * kWarpE + r;
        gdoc[doc[r]] = p < nb ? gl[r] * gmul : 0.0f;
      }
      __syncwarp();
      if (vec_ok) {
        if (lane * kWarpE < L)
          *reinterpret_cast<float4*>(grad_out + base + lane * kWarpE) =
              *reinterpret_cast<const float4*>(gdoc + lane * kWarpE);
      } else {
        for (int j = lane; j < L; j += 32) grad_out[base + j] = gdoc[j];
      }
    }
    __syncwarp();
    if (!dynamic) { q += total_warps; continue; }
    q = __shfl_sync(0xffffffffu, b_next, 0);
  }

  // ---- leave the queue clean for the next launch that uses this slot ----------------------------------
  if (dynamic && lane == 0) {
    const unsigned int done = atomicAdd(queue + 1, 1u);
    if (done == total_warps - 1) {
      queue[0] = 0u;
      queue[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace ltr
