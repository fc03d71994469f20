// ltr_pair_warp.cuh -- one WARP per query for list sizes up to 128 (the headline (4096, 128)
// workload): no CTA barrier in the steady state, everything of a query lives in the warp's
// registers plus 3.5 KB of shared memory.
//
// Per query: 128-bit loads of the padded row -> in-register bitonic argsort of (score, index)
// keys (rank_by_score, utils/tensor_operations.py:48-64) -> optional second sort of the
// relevance grades for the ideal DCG (_max_dcg, pairwise_lambda.py:231-241) -> per-document
// factors into shared memory in rank order -> ring_pass (ltr_pair_tiles.cuh) -> gradient
// scattered back to document order (backward of the gather, pairwise_lambda.py:69) and stored
// with 128-bit writes.
#pragma once

#include "ltr_pair_tiles.cuh"

namespace ltr {

constexpr int kWarpL = 128;          // max list size of the warp-per-query kernel
constexpr int kWarpE = 4;            // elements per lane in the sort (32 * 4 = 128)
constexpr int kWarpsPerCta = 8;

// ---- warp-wide bitonic sort of 32*E 64-bit keys, element index = lane * E + r, ascending ----
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
            const bool up = (((lane * E + r) & size) == 0);
            const uint64_t a = k[r], b = k[q];
            const bool sw = (a > b) == up;
            k[r] = sw ? b : a;
            k[q] = sw ? a : b;
          }
        }
      } else {
        const int ls = stride / E;
        const bool lower = (lane & ls) == 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint64_t mine = k[r];
          const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, ls);
          const bool up = (((lane * E + r) & size) == 0);
          const bool keep_min = (lower == up);
          const bool other_smaller = other < mine;
          k[r] = (keep_min == other_smaller) ? other : mine;
        }
      }
    }
  }
}

// Same for 32-bit keys (relevance grades of the ideal ranking; no payload needed).
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
            const bool up = (((lane * E + r) & size) == 0);
            const uint32_t lo = min(k[r], k[q]), hi = max(k[r], k[q]);
            k[r] = up ? lo : hi;
            k[q] = up ? hi : lo;
          }
        }
      } else {
        const int ls = stride / E;
        const bool lower = (lane & ls) == 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, k[r], ls);
          const bool up = (((lane * E + r) & size) == 0);
          k[r] = (lower == up) ? min(k[r], other) : max(k[r], other);
        }
      }
    }
  }
}

__device__ __forceinline__ float score_from_desc_key(uint32_t key) {
  const uint32_t asc = ~key;
  const uint32_t u = (asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc;
  return __uint_as_float(u);
}

struct WarpScratch {
  PairItem items[kWarpL];   // rank order
  float gcol[kWarpL];       // column gradients, rank order
  int raw_y[kWarpL];        // relevance in document order; reused as the document-order gradient
};

struct WarpTables {
  float delta[kWarpL + 8];                 // delta[k] = |1/D(k) - 1/D(k+1)|
  float disc[kWarpL];                      // D(r) = log2(2 + r)
  float wtab[4][window_table_floats()];    // window tables for R = 1..4
};

template <int TW, bool FACTORED, int R>
__device__ __forceinline__ float ring_dispatch(WarpScratch& ws, const WarpTables& tb, int n, int lane) {
  const int C = (n + R - 1) / R;
  float racc[R];
  const float l = ring_pass<TW, FACTORED, R>(ws.items, ws.gcol, tb.wtab[R - 1], C, n, lane, racc);
  if (lane < C) {
#pragma unroll
    for (int r = 0; r < R; ++r) ws.gcol[lane * R + r] += racc[r];
  }
  return l;
}

template <int TW>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
pair_warp_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                 const void* __restrict__ n, int n_bytes, int B, int L, float sigma, int vec_ok,
                 float* __restrict__ loss_out, float* __restrict__ grad_out,
                 int64_t* __restrict__ ranking_out, float* __restrict__ loss_sum) {
  __shared__ WarpTables tb;
  __shared__ WarpScratch scratch[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  if constexpr (TW == TW_DELTA) {
    for (int k = threadIdx.x; k < kWarpL + 8; k += blockDim.x) {
      const float d0 = log2f(2.0f + static_cast<float>(k));
      const float d1 = log2f(3.0f + static_cast<float>(k));
      tb.delta[k] = fabsf(1.0f / d0 - 1.0f / d1);
      if (k < kWarpL) tb.disc[k] = d0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * window_table_floats(); i += blockDim.x) {
      const int R = i / window_table_floats() + 1;
      const int rem = i % window_table_floats();
      const int d = rem / 8 - kMaxChunks, slot = rem % 8;
      int k = R * d + slot - (R - 1);
      k = k < 0 ? -k : k;
      tb.wtab[R - 1][rem] = k < kWarpL + 8 ? tb.delta[k] : 0.0f;
    }
    __syncthreads();
  }

  WarpScratch& ws = scratch[warp];
  const float gscale = sigma * kLog2e;   // lambda = sigma / ln 2 * w * sigmoid(-x)

  for (int b = blockIdx.x * kWarpsPerCta + warp; b < B; b += gridDim.x * kWarpsPerCta) {
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
        const long long t[4] = {r0.x, r0.y, r1.x, r1.y};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          long long v = t[r];
          v = v < -2147483647LL ? -2147483647LL : (v > 2147483647LL ? 2147483647LL : v);
          yv[r] = static_cast<int>(v);
        }
      } else {
        const int4 r4 = *reinterpret_cast<const int4*>(reinterpret_cast<const int*>(rel) + base + lane * kWarpE);
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

    // ---- argsort by descending score; padding (and the slots beyond L) last, by index ------
    uint64_t key[kWarpE];
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) {
      const int j = lane * kWarpE + r;
      ws.raw_y[j] = yv[r];
      key[r] = pack_key(j < nb ? desc_key_f32(sv[r]) : kPadKey, j);
    }
    warp_bitonic_sort64<kWarpE>(key, lane);
    __syncwarp();

    // ---- ideal DCG: relevance descending over the valid documents ---------------------------
    float inv_max_dcg = 1.0f;
    if constexpr (TW == TW_DELTA) {
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
      float max_dcg = warp_sum(part);
      if (max_dcg == 0.0f) max_dcg = 1.0f;
      inv_max_dcg = max_dcg;   // divided below exactly as the reference does (gain / max_dcg)
    }

    // ---- rank-ordered scores / relevance, score range ----------------------------------------
    float ss[kWarpE];
    int ys[kWarpE], doc[kWarpE];
    float smax = -INFINITY, smin = INFINITY;
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) {
      const int p = lane * kWarpE + r;
      doc[r] = static_cast<int>(key[r] & 0xffffffffu);
      ss[r] = score_from_desc_key(static_cast<uint32_t>(key[r] >> 32));
      ys[r] = ws.raw_y[doc[r]];
      if (p < nb) { smax = fmaxf(smax, ss[r]); smin = fminf(smin, ss[r]); }
      if (ranking_out && p < L) ranking_out[base + p] = doc[r];
    }
    smax = warp_max(smax);
    smin = -warp_max(-smin);
    const float mid = 0.5f * (smax + smin);
    // NaN / inf scores fail this test and take the stable form
    const bool factored = fabsf(sigma) * (smax - smin) * kLog2e <= kFactoredRange;

    // ---- per-document factors -> shared memory (rank order) ------------------------------------
#pragma unroll
    for (int r = 0; r < kWarpE; ++r) {
      const int p = lane * kWarpE + r;
      PairItem it;
      if (p < nb) {
        if constexpr (TW == TW_DELTA) it.g = exp_gain_f32(ys[r]) / inv_max_dcg;
        else it.g = static_cast<float>(ys[r]);
        if (factored) {
          const double ed = static_cast<double>(ss[r] - mid) * static_cast<double>(sigma) * 1.4426950408889634;
          const float eh = static_cast<float>(ed);
          const float el = static_cast<float>(ed - static_cast<double>(eh)) * kLn2;
          it.e = eh;
          it.a = exp2f(-eh) * (1.0f - el);
          it.b = exp2f(eh) * (1.0f + el);
        } else {
          it.a = sigma * ss[r];
          it.b = 0.0f;
          it.e = 0.0f;
        }
      } else {
        it.a = factored ? 0.0f : -1.0e30f;
        it.b = 0.0f;
        it.e = 0.0f;
        it.g = 0.0f;
      }
      ws.items[p] = it;
      ws.gcol[p] = 0.0f;
    }
    __syncwarp();

    // ---- all pairs, once ----------------------------------------------------------------------------
    float lacc = 0.0f;
    if (nb > 1) {
      if (factored) {
        const int R = (nb + 31) >> 5;
        if (R == 1) lacc = ring_dispatch<TW, true, 1>(ws, tb, nb, lane);
        else if (R == 2) lacc = ring_dispatch<TW, true, 2>(ws, tb, nb, lane);
        else if (R == 3) lacc = ring_dispatch<TW, true, 3>(ws, tb, nb, lane);
        else lacc = ring_dispatch<TW, true, 4>(ws, tb, nb, lane);
      } else {
        lacc = ring_dispatch<TW, false, 4>(ws, tb, nb, lane);
      }
    }
    __syncwarp();
    const float loss = warp_sum(lacc);
    if (lane == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }

    // ---- gradient back to document order -----------------------------------------------------------
    if (grad_out) {
      float* gdoc = reinterpret_cast<float*>(ws.raw_y);
#pragma unroll
      for (int r = 0; r < kWarpE; ++r) {
        const int p = lane * kWarpE + r;
        gdoc[doc[r]] = p < nb ? ws.gcol[p] * gscale : 0.0f;
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
  }
}

}  // namespace ltr
