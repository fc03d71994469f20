// ltr_pair_cta.cuh -- one CTA per query for list sizes 129 .. LTR_MAX_LIST_SIZE (the north-star
// point (4096, 1024), config (65536, 512)): the sigmoid-weighted pair losses evaluated ONCE per
// unordered pair on 128 x 128 rank tiles.
//
// Per query: the padded row (scores + relevance) is staged into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier), double buffered so that the NEXT query's row is in flight while
// the current one is being processed (plain coalesced loads when the row is not 16-byte
// aligned or shared memory is short) -> CTA-wide bitonic argsort in shared memory
// (rank_by_score, utils/tensor_operations.py:48-64) -> ideal DCG from a shared-memory grade
// histogram (_max_dcg, pairwise_lambda.py:231-241; a second sort if the grades do not fit) ->
// per-document factors in rank order -> the S (S + 1) / 2 rank tiles are handed to the warps of
// the CTA through a shared counter: diagonal tiles run ring_pass, off-diagonal tiles tile_pass
// (ltr_pair_tiles.cuh) -> per-tile row / column gradients are added to the CTA's rank-order
// gradient with shared-memory atomics (8 per lane per 512 pairs) -> scatter to document order,
// coalesced store.
#pragma once

#include "ltr_pair_warp.cuh"

namespace ltr {

constexpr int kCtaWarps = 8;

struct CtaSmem {
  uint64_t* keys;     // [P]      sort keys
  PairSoA it;         // 4 x [Lp] rank order, Lp = L rounded up to 128
  float* delta;       // [Lp + 8] delta table (copied from PairTables once per CTA)
  float* raw_s;       // [L]      scores, document order; reused: document-order gradient
  int* raw_y;         // [L]      relevance, document order
  int* doc;           // [Lp]     rank -> document
  float* gacc;        // [Lp]     rank-order gradient (unscaled)
  float* gcol;        // [W*256]  per-warp column accumulators of the current tile (128) + the dump of ring
                      //          lanes that own no chunk (ring_pass_fact: chunk 32 + lane)
  float* red;         // [40]
  int* hist;          // [36]     grade histogram + flags + tile counter
  // TMA staging (tma != 0): two row buffers and their mbarriers; raw_s then aliases tma_s[cur]
  float* tma_s1;          // [L]  second score buffer (the first one is raw_s)
  unsigned char* tma_y0;  // [L * rel_bytes] relevance as stored in HBM, buffer 0
  unsigned char* tma_y1;  // buffer 1
  uint64_t* bars;         // [2]
};

__host__ __device__ inline size_t cta_align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

// rel_bytes = 0: no TMA staging buffers
__host__ __device__ inline size_t cta_smem_bytes(int L, int P, int tma_rel_bytes) {
  const size_t Lp = (static_cast<size_t>(L) + 127) / 128 * 128;
  size_t bytes = 8u * P + 16u * Lp + 4u * (Lp + 8) + 4u * kCtaWarps * 256 + 4u * Lp + 4u * Lp +
                 cta_align16(4u * L) + cta_align16(4u * L) + 4u * 40 + 4u * 40;
  if (tma_rel_bytes) bytes += cta_align16(4u * L) + 2u * cta_align16(static_cast<size_t>(tma_rel_bytes) * L) + 16u;
  return bytes;
}

__device__ __forceinline__ CtaSmem cta_carve(unsigned char* base, int L, int P, int tma_rel_bytes) {
  const int Lp = (L + 127) / 128 * 128;
  CtaSmem m;
  m.keys = reinterpret_cast<uint64_t*>(base);                 base += 8u * P;
  m.it.a = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.b = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.e = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.it.g = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.delta = reinterpret_cast<float*>(base);                   base += 4u * (Lp + 8);
  m.gcol = reinterpret_cast<float*>(base);                    base += 4u * kCtaWarps * 256;
  m.gacc = reinterpret_cast<float*>(base);                    base += 4u * Lp;
  m.doc = reinterpret_cast<int*>(base);                       base += 4u * Lp;
  m.raw_s = reinterpret_cast<float*>(base);                   base += cta_align16(4u * L);
  m.raw_y = reinterpret_cast<int*>(base);                     base += cta_align16(4u * L);
  m.red = reinterpret_cast<float*>(base);                     base += 4u * 40;
  m.hist = reinterpret_cast<int*>(base);                      base += 4u * 40;
  m.tma_s1 = nullptr; m.tma_y0 = nullptr; m.tma_y1 = nullptr; m.bars = nullptr;
  if (tma_rel_bytes) {
    m.tma_s1 = reinterpret_cast<float*>(base);                base += cta_align16(4u * L);
    m.tma_y0 = base;                                          base += cta_align16(static_cast<size_t>(tma_rel_bytes) * L);
    m.tma_y1 = base;                                          base += cta_align16(static_cast<size_t>(tma_rel_bytes) * L);
    m.bars = reinterpret_cast<uint64_t*>(base);
  }
  return m;
}

template <int TW, bool FACTORED>
__device__ __forceinline__ float cta_tiles(const CtaSmem& m, const PairTables& tb, int nb, int lane, int warp) {
  const int S = (nb + 127) >> 7;
  const int T = S * (S + 1) / 2;
  float* gcol = m.gcol + warp * 256;
  float wl = 0.0f;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&m.hist[34], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= T) break;
    // strip A holds the tiles (A, A), (A, A + 1), ..., (A, S - 1)
    int A = 0, rem = t;
    while (rem >= S - A) { rem -= S - A; ++A; }
    const int Bc = A + rem;
    const int row_base = A << 7, col_base = Bc << 7;
    *reinterpret_cast<float4*>(gcol + lane * 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncwarp();
    if (Bc == A) {
      // diagonal block: ring over the (possibly partial) block with the densest chunking
      const int nl = min(128, nb - row_base);
      const int R = (nl + 31) >> 5;
      const PairSoA it = {m.it.a + row_base, m.it.b + row_base, m.it.e + row_base, m.it.g + row_base};
      float racc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int C;
      float l;
      if (R == 1) { float ra1[1]; C = nl; l = ring_pass<TW, FACTORED, 1>(it, gcol, tb.wtab[0], C, nl, lane, ra1); racc[0] = ra1[0]; }
      else if (R == 2) { float ra2[2]; C = (nl + 1) / 2; l = ring_pass<TW, FACTORED, 2>(it, gcol, tb.wtab[1], C, nl, lane, ra2); racc[0] = ra2[0]; racc[1] = ra2[1]; }
      else if (R == 3) { float ra3[3]; C = (nl + 2) / 3; l = ring_pass<TW, FACTORED, 3>(it, gcol, tb.wtab[2], C, nl, lane, ra3); racc[0] = ra3[0]; racc[1] = ra3[1]; racc[2] = ra3[2]; }
      else { C = (nl + 3) / 4; l = ring_pass<TW, FACTORED, 4>(it, gcol, tb.wtab[3], C, nl, lane, racc); }
      wl += l;
      __syncwarp();
      if (lane < C) {
        for (int r = 0; r < R; ++r) {
          const int p = lane * R + r;
          if (p < nl) atomicAdd(&m.gacc[row_base + p], gcol[lane * 4 + r] + racc[r]);
        }
      }
    } else {
      float racc[4];
      wl += tile_pass<TW, FACTORED>(m.it, row_base, col_base, m.delta, gcol, lane, racc);
      __syncwarp();
      const float4 gc = *reinterpret_cast<const float4*>(gcol + lane * 4);
      const float gcv[4] = {gc.x, gc.y, gc.z, gc.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        atomicAdd(&m.gacc[row_base + lane * 4 + r], racc[r]);
        if (col_base + lane * 4 + r < nb) atomicAdd(&m.gacc[col_base + lane * 4 + r], gcv[r]);
      }
    }
    __syncwarp();
  }
  return wl;
}

template <int TW>
__global__ void __launch_bounds__(kCtaWarps * 32, 3)
pair_cta_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                const void* __restrict__ n, int n_bytes, int B, int L, int P, float sigma, int variant,
                int tma, float* __restrict__ loss_out, float* __restrict__ grad_out,
                int64_t* __restrict__ ranking_out, float* __restrict__ loss_sum,
                const PairTables* __restrict__ tabs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaSmem m = cta_carve(smem_raw, L, P, tma ? rel_bytes : 0);
  const PairTables& tb = *tabs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float gscale = sigma * kLog2e;
  const double kd = static_cast<double>(sigma) * 1.4426950408889634;
  const float k_hi = static_cast<float>(kd);
  const float k_lo = static_cast<float>(kd - static_cast<double>(k_hi));

  const bool rank_weighted = TW == TW_DELTA || (TW == TW_TWO && variant != 0);   // NDCG losses
  if constexpr (TW == TW_DELTA) {
    const int Lp0 = (L + 127) / 128 * 128;
    for (int k = threadIdx.x; k < Lp0 + 8; k += blockDim.x) m.delta[k] = tb.delta[k];
  }

  // TMA row staging: one thread arms the buffer's mbarrier with the byte count and issues the two
  // bulk copies; everybody waits on the barrier's phase before touching the row.
  const uint32_t row_s_bytes = 4u * L, row_y_bytes = static_cast<uint32_t>(rel_bytes) * L;
  float* const buf_s0 = m.raw_s;
  float* const buf_s1 = m.tma_s1;
  auto issue_row = [&](int q, int buf) {
    uint64_t* bar = m.bars + buf;
    mbar_arrive_expect_tx(bar, row_s_bytes + row_y_bytes);
    tma_load_1d(buf ? buf_s1 : buf_s0, scores + static_cast<size_t>(q) * L, row_s_bytes, bar);
    tma_load_1d(buf ? m.tma_y1 : m.tma_y0,
                static_cast<const unsigned char*>(rel) + static_cast<size_t>(q) * row_y_bytes, row_y_bytes, bar);
  };
  if (tma) {
    if (threadIdx.x == 0) {
      mbar_init(m.bars, 1);
      mbar_init(m.bars + 1, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < B) issue_row(blockIdx.x, 0);
  }

  int iter = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x, ++iter) {
    __syncthreads();   // previous query fully consumed
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    const int Lp = ((nb + 127) >> 7) << 7;   // ranks covered by tiles
    if (tma) {
      const int cur = iter & 1;
      if (threadIdx.x == 0 && b + static_cast<int>(gridDim.x) < B) {
        fence_proxy_async();   // the other buffer was last written through the generic proxy
        issue_row(b + gridDim.x, cur ^ 1);
      }
      mbar_wait(m.bars + cur, (iter >> 1) & 1);
      m.raw_s = cur ? buf_s1 : buf_s0;
      const unsigned char* ybuf = cur ? m.tma_y1 : m.tma_y0;
      for (int j = threadIdx.x; j < L; j += blockDim.x) m.raw_y[j] = load_int_clamped(ybuf, rel_bytes, j);
    } else {
      for (int j = threadIdx.x; j < L; j += blockDim.x) {
        m.raw_s[j] = scores[base + j];
        m.raw_y[j] = load_int_clamped(rel, rel_bytes, base + j);
      }
    }
    if (threadIdx.x < 36) m.hist[threadIdx.x] = 0;
    __syncthreads();

    // ---- rank_by_score (only LambdaNDCGLoss2 needs rank order; the other losses are permutation
    // invariant and stay in document order unless the ranking was asked for) -----------------------
    const bool sorted = rank_weighted || ranking_out != nullptr;
    if (sorted) {
      for (int j = threadIdx.x; j < P; j += blockDim.x) {
        uint32_t key = kPadKey;
        if (j < nb) key = desc_key_f32(m.raw_s[j]);
        m.keys[j] = j < L ? pack_key(key, j) : ~0ull;
      }
      cta_bitonic_sort(m.keys, P);
    } else {
      for (int j = threadIdx.x; j < L; j += blockDim.x) m.keys[j] = static_cast<uint64_t>(j);
      __syncthreads();
    }

    // ---- ideal DCG ---------------------------------------------------------------------------------
    float max_dcg = 1.0f;
    if (rank_weighted) {
      for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int y = m.raw_y[j];
        if (y < 0 || y > 31) m.hist[32] = 1;
        else if (y > 0) atomicAdd(&m.hist[y], 1);
      }
      __syncthreads();
      if (m.hist[32] == 0) {
        if (threadIdx.x == 0) {
          double acc = 0.0;
          int start = 0;
          for (int g = 31; g >= 1; --g) {
            const int cnt = m.hist[g];
            if (cnt > 0) {
              acc += static_cast<double>(gain_of_grade(g)) *
                     (tb.inv_disc_prefix[start + cnt] - tb.inv_disc_prefix[start]);
              start += cnt;
            }
          }
          m.red[36] = static_cast<float>(acc);
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
        cta_bitonic_sort(m.keys, P);
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
      for (int w = 1; w < kCtaWarps; ++w) { smax = fmaxf(smax, m.red[w]); smin = fminf(smin, m.red[8 + w]); }
      __syncthreads();
    }
    const float mid = 0.5f * (smax + smin);
    const bool factored = TW != TW_HINGE && fabsf(sigma) * (smax - smin) * kLog2e <= kFactoredRange;
    const int fill = Lp > L ? Lp : L;
    // winner-by-relevance losses: padded columns carry the smallest valid weight (both forms)
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
    for (int p = threadIdx.x; p < fill; p += blockDim.x) {
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
      if (p < Lp) { m.it.a[p] = fa; m.it.b[p] = fb; m.it.e[p] = fe; m.it.g[p] = fg; m.gacc[p] = 0.0f; }
      if (p < L) m.doc[p] = d;
    }
    __syncthreads();

    // ---- all pairs, once ---------------------------------------------------------------------------------
    float wl = 0.0f;
    if (nb > 1) {
      if constexpr (TW == TW_HINGE) wl = cta_tiles<TW, false>(m, tb, nb, lane, warp);
      else wl = factored ? cta_tiles<TW, true>(m, tb, nb, lane, warp) : cta_tiles<TW, false>(m, tb, nb, lane, warp);
    }
    float loss = cta_sum(wl + diag, m.red);   // barriers inside also publish gacc
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

    // ---- gradient back to document order ----------------------------------------------------------------
    if (grad_out) {
      float* gdoc = m.raw_s;
      for (int p = threadIdx.x; p < L; p += blockDim.x) gdoc[m.doc[p]] = p < nb ? m.gacc[p] * gmul : 0.0f;
      __syncthreads();
      float* __restrict__ go = grad_out + base;
      for (int j = threadIdx.x; j < L; j += blockDim.x) go[j] = gdoc[j];
    }
  }
}

}  // namespace ltr
