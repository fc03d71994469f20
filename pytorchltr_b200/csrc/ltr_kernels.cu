// ltr_kernels.cu -- sm_100a kernels + C ABI of libltr_sm100.so (see include/ltr_sm100.h).
//
// One fused launch per loss call: each CTA stages a query's padded row into shared
// memory, ranks it (bitonic argsort), builds the O(L) weight tables the loss needs
// (gains, discounts, delta) and then generates the O(L^2) document pairs on the fly
// from shared memory, producing the per-query loss AND d loss / d scores in the same
// pass.  Nothing of size L^2 ever exists in memory (the reference materialises
// ~78 L^2 bytes per query, utils/tensor_operations.py:94-119).
//
// Reference citations are relative to the reference root (pytorchltr/...).
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>

#include "ltr_collate.cuh"
#include "ltr_common.cuh"
#include "ltr_hinge_sorted.cuh"
#include "ltr_host.cuh"
#include "ltr_linear_listnet.cuh"
#include "ltr_metrics_warp.cuh"
#include "ltr_pair_cta.cuh"
#include "ltr_pair_ring.cuh"
#include "ltr_pair_warp.cuh"
#include "ltr_p2p.cuh"
#include "ltr_sm100.h"

namespace ltr {

enum PairMode : int {
  PM_HINGE = 0,
  PM_DCG_HINGE = 1,
  PM_LOGISTIC = 2,
  PM_ARP1 = 3,
  PM_ARP2 = 4,
  PM_NDCG1 = 5,
  PM_NDCG2 = 6,
};

__host__ __device__ constexpr bool mode_needs_rank(int m) { return m == PM_NDCG1 || m == PM_NDCG2; }
__host__ __device__ constexpr bool mode_is_lambda(int m) { return m >= PM_ARP1; }

// Shared-memory carve-up of the generic pair kernel (and the metric kernels).
struct RowSmem {
  uint64_t* keys;  // [P]   sort keys
  float* raw_s;    // [L]   scores, document order
  int* raw_y;      // [L]   relevance, document order
  float* s;        // [L]   scores, rank order
  int* y;          // [L]   relevance, rank order
  float* w;        // [L]   per-item weight (gain G, or G/D, or float(rel)), rank order
  int* doc;        // [L]   rank -> document index
  float* delta;    // [L]   |1/D(k) - 1/D(k+1)|, pairwise_lambda.py:206-211
  float* disc;     // [L]   D(r) = log2(2 + r),    pairwise_lambda.py:168-170 / dcg.py:93
  float* gout;     // [L]   gradient, document order
  float* red;      // [40]  reduction scratch
};

__host__ __device__ inline size_t row_smem_bytes(int L, int P) {
  return sizeof(uint64_t) * static_cast<size_t>(P) + static_cast<size_t>(L) * 4u * 9u + 40u * 4u;
}

__device__ __forceinline__ RowSmem carve(unsigned char* base, int L, int P) {
  RowSmem m;
  m.keys = reinterpret_cast<uint64_t*>(base);
  float* f = reinterpret_cast<float*>(base + sizeof(uint64_t) * static_cast<size_t>(P));
  m.raw_s = f;
  m.raw_y = reinterpret_cast<int*>(f + L);
  m.s = f + 2 * L;
  m.y = reinterpret_cast<int*>(f + 3 * L);
  m.w = f + 4 * L;
  m.doc = reinterpret_cast<int*>(f + 5 * L);
  m.delta = f + 6 * L;
  m.disc = f + 7 * L;
  m.gout = f + 8 * L;
  m.red = f + 9 * L;
  return m;
}

// Stage one padded row into shared memory (coalesced).
__device__ __forceinline__ void stage_row(const RowSmem& m, const float* __restrict__ scores,
                                          const void* __restrict__ rel, int rel_bytes, int b, int L) {
  const size_t base = static_cast<size_t>(b) * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    m.raw_s[j] = scores[base + j];
    m.raw_y[j] = load_int_clamped(rel, rel_bytes, base + j);
  }
}

// rank_by_score (utils/tensor_operations.py:48-64): valid documents by descending score,
// padded documents last.  Leaves the sorted (key, doc) pairs in m.keys.
__device__ __forceinline__ void rank_row_by_score(const RowSmem& m, int nb, int L, int P) {
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    uint32_t key = kPadKey;
    if (j < nb) key = desc_key_f32(m.raw_s[j]);
    m.keys[j] = j < L ? pack_key(key, j) : ~0ull;
  }
  cta_bitonic_sort(m.keys, P);
}

// Ideal ranking: relevance descending, padding last (_max_dcg, pairwise_lambda.py:233;
// ndcg's dcg(relevance.float(), ...), dcg.py:36).
__device__ __forceinline__ void rank_row_by_relevance(const RowSmem& m, int nb, int L, int P) {
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    uint32_t key = kPadKey;
    if (j < nb) key = desc_key_i32(m.raw_y[j]);
    m.keys[j] = j < L ? pack_key(key, j) : ~0ull;
  }
  cta_bitonic_sort(m.keys, P);
}

// ---------------------------------------------------------------------------------------
// Row pass of the generic pair kernel: thread owns item `a` (rank order) and visits every
// b in [0, nb).  Each ordered pair is seen from both endpoints, so a thread only ever
// accumulates its own gradient: no synchronisation, at the price of evaluating the
// sigmoid of every unordered pair twice (the tiled kernels in ltr_pair_tiles.cuh
// evaluate it once).
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void row_pass(const RowSmem& m, int a, int nb, float sigma,
                                         float& lacc, float& gacc) {
  const float sa = m.s[a];
  const int ya = m.y[a];
  const float wa = m.w[a];
  for (int b = 0; b < nb; ++b) {
    const float sb = m.s[b];
    const int yb = m.y[b];
    if constexpr (MODE == PM_HINGE || MODE == PM_DCG_HINGE) {
      // pairwise_additive.py:108-112: loss = 1.0 - (s_i - s_j) (two float32 roundings);
      // 0 where rel_i - rel_j <= 0; 0 where loss < 0 (the kink loss == 0 keeps grad -1/+1).
      const float d = sa - sb;
      if (ya > yb) {
        const float l = 1.0f - d;
        if (!(l < 0.0f)) { lacc += l; gacc -= 1.0f; }
      } else if (yb > ya) {
        const float l = 1.0f + d;  // == 1.0f - (s_b - s_a) bit for bit
        if (!(l < 0.0f)) gacc += 1.0f;
      }
    } else if constexpr (MODE == PM_LOGISTIC || MODE == PM_ARP2 || MODE == PM_NDCG2) {
      // loss_ij = w_ij * log2(1 + exp(-sigma (s_i - s_j))) for rel_i > rel_j with
      //   w = 1                         pairwise_additive.py:161-162
      //   w = rel_i - rel_j             pairwise_lambda.py:137-140
      //   w = delta_|i-j| |G_i - G_j|   pairwise_lambda.py:201-218
      const int yd = ya - yb;
      if (yd != 0) {
        float w;
        if constexpr (MODE == PM_LOGISTIC) w = 1.0f;
        else if constexpr (MODE == PM_ARP2) w = fabsf(static_cast<float>(yd));
        else w = m.delta[a > b ? a - b : b - a] * fabsf(wa - m.w[b]);
        float x = sigma * (sa - sb);
        if (yd < 0) x = -x;                    // sigma * (s_winner - s_loser)
        const float u = -fabsf(x) * kLog2e;
        const float t = ex2_approx(u);         // exp(-|x|) in (0, 1]: never overflows
        const float p = 1.0f + t;
        const float r = rcp_approx(p);
        const float sg = x >= 0.0f ? t * r : r;  // sigmoid(-x)
        const float lam = w * sg;
        if (yd > 0) {
          gacc -= lam;
          float l = lg2_approx(p);             // log2(1 + e^-x) = max(-x, 0) log2 e + log2(1 + e^-|x|)
          if (x < 0.0f) l -= u;
          lacc = fmaf(w, l, lacc);
        } else {
          gacc += lam;
        }
      }
    } else {
      // ARP1 / NDCG1: every ordered pair incl. the diagonal, weight of the FIRST index:
      //   w_i = rel_i        pairwise_lambda.py:114-117
      //   w_i = G_i / D_i    pairwise_lambda.py:165-173
      const float wb = m.w[b];
      const float x = sigma * (sa - sb);
      const float u = -fabsf(x) * kLog2e;
      const float t = ex2_approx(u);
      const float p = 1.0f + t;
      const float r = rcp_approx(p);
      const float tr = t * r;
      const float s_neg = x >= 0.0f ? tr : r;  // sigmoid(-x): pair (a, b)
      const float s_pos = x >= 0.0f ? r : tr;  // sigmoid(+x): pair (b, a)
      float l = lg2_approx(p);
      if (x < 0.0f) l -= u;
      lacc = fmaf(wa, l, lacc);
      gacc += wb * s_pos - wa * s_neg;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(1024)
pair_loss_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                 const void* __restrict__ n, int n_bytes, int B, int L, int P, float sigma,
                 float* __restrict__ loss_out, float* __restrict__ grad_out,
                 int64_t* __restrict__ ranking_out, float* __restrict__ loss_sum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RowSmem m = carve(smem_raw, L, P);
  const bool want_rank = mode_needs_rank(MODE) || ranking_out != nullptr;

  // Score-independent tables, once per (persistent) CTA.
  if constexpr (MODE == PM_NDCG1 || MODE == PM_NDCG2) {
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
      const float d0 = log2f(2.0f + static_cast<float>(k));
      const float d1 = log2f(3.0f + static_cast<float>(k));
      m.disc[k] = d0;
      m.delta[k] = fabsf(1.0f / d0 - 1.0f / d1);
    }
  }

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();  // previous query fully consumed
    const int nb = load_n(n, n_bytes, b, L);
    stage_row(m, scores, rel, rel_bytes, b, L);
    __syncthreads();

    if (want_rank) {
      rank_row_by_score(m, nb, L, P);
      for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const int d = static_cast<int>(m.keys[r] & 0xffffffffu);
        m.doc[r] = d;
        m.s[r] = m.raw_s[d];
        m.y[r] = m.raw_y[d];
        if (ranking_out) ranking_out[static_cast<size_t>(b) * L + r] = d;
      }
    } else {
      for (int r = threadIdx.x; r < L; r += blockDim.x) {
        m.doc[r] = r;
        m.s[r] = m.raw_s[r];
        m.y[r] = m.raw_y[r];
      }
    }

    if constexpr (MODE == PM_NDCG1 || MODE == PM_NDCG2) {
      // _max_dcg (pairwise_lambda.py:231-241) then _ndcg_gains (:221-228)
      __syncthreads();  // everyone is done reading the score ranking out of m.keys
      rank_row_by_relevance(m, nb, L, P);
      float part = 0.0f;
      for (int r = threadIdx.x; r < nb; r += blockDim.x) {
        const int d = static_cast<int>(m.keys[r] & 0xffffffffu);
        part += exp_gain_f32(m.raw_y[d]) / m.disc[r];
      }
      float max_dcg = cta_sum(part, m.red);
      if (max_dcg == 0.0f) max_dcg = 1.0f;
      for (int r = threadIdx.x; r < L; r += blockDim.x) {
        const float g = exp_gain_f32(m.y[r]) / max_dcg;
        m.w[r] = MODE == PM_NDCG1 ? g / m.disc[r] : g;
      }
    } else {
      for (int r = threadIdx.x; r < L; r += blockDim.x) m.w[r] = static_cast<float>(m.y[r]);
    }
    for (int j = threadIdx.x; j < L; j += blockDim.x) m.gout[j] = 0.0f;
    __syncthreads();

    float lacc = 0.0f;
    for (int a = threadIdx.x; a < nb; a += blockDim.x) {
      float gacc = 0.0f;
      row_pass<MODE>(m, a, nb, sigma, lacc, gacc);
      m.gout[m.doc[a]] = gacc;   // backward of the gather (pairwise_lambda.py:69)
    }
    float loss = cta_sum(lacc, m.red);   // barriers inside also publish gout

    float gscale = 1.0f;
    if constexpr (MODE == PM_DCG_HINGE) {
      // pairwise_additive.py:132-133: -1 / ln(2 + h); d/dh = 1 / ((2 + h) ln^2(2 + h))
      const float lg = logf(2.0f + loss);
      gscale = 1.0f / ((2.0f + loss) * lg * lg);
      loss = -1.0f / lg;
    } else if constexpr (MODE != PM_HINGE) {
      gscale = sigma * kLog2e;           // lambda = -sigma w sigmoid(-x) / ln 2
    }
    if (threadIdx.x == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }
    if (grad_out) {
      float* __restrict__ go = grad_out + static_cast<size_t>(b) * L;
      for (int j = threadIdx.x; j < L; j += blockDim.x) go[j] = m.gout[j] * gscale;
    }
  }
}

// ---------------------------------------------------------------------------------------
// ListNet: one warp per query, three passes over the (L1-resident) row.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
listnet_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
               const void* __restrict__ n, int n_bytes, int B, int L,
               float* __restrict__ loss_out, float* __restrict__ grad_out,
               float* __restrict__ loss_sum) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  for (int b = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); b < B; b += gridDim.x * warps_per_cta) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    float ms = -INFINITY, my = -INFINITY;
    for (int j = lane; j < nb; j += 32) {
      ms = fmaxf(ms, scores[base + j]);
      my = fmaxf(my, static_cast<float>(load_int_clamped(rel, rel_bytes, base + j)));
    }
    ms = warp_max(ms);
    my = warp_max(my);
    float zs = 0.0f, zy = 0.0f, a = 0.0f;
    for (int j = lane; j < nb; j += 32) {
      const float ds = scores[base + j] - ms;
      const float ey = expf(static_cast<float>(load_int_clamped(rel, rel_bytes, base + j)) - my);
      zs += expf(ds);
      zy += ey;
      a = fmaf(ey, ds, a);
    }
    zs = warp_sum(zs);
    zy = warp_sum(zy);
    a = warp_sum(a);
    // loss = -sum_j P_j (s_j - ms - ln zs) = ln zs - (sum_j e^{y_j - my} (s_j - ms)) / zy
    const float loss = nb > 0 ? logf(zs) - a / zy : 0.0f;
    if (lane == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }
    if (grad_out) {
      const float izs = nb > 0 ? 1.0f / zs : 0.0f, izy = nb > 0 ? 1.0f / zy : 0.0f;
      for (int j = lane; j < L; j += 32) {
        float g = 0.0f;
        if (j < nb) {
          const float ey = expf(static_cast<float>(load_int_clamped(rel, rel_bytes, base + j)) - my);
          g = expf(scores[base + j] - ms) * izs - ey * izy;
        }
        grad_out[base + j] = g;
      }
    }
  }
}

// Register-resident form for L <= 32 K: the row is read ONCE (K coalesced loads per lane, all
// in flight together), reduced with shuffles and the gradient written once: 12 L bytes read and
// 4 L written per query, which is the algorithmic minimum (HBM-bound).
template <int K>
__global__ void __launch_bounds__(256)
listnet_reg_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                   const void* __restrict__ n, int n_bytes, int B, int L,
                   float* __restrict__ loss_out, float* __restrict__ grad_out,
                   float* __restrict__ loss_sum) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  for (int b = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); b < B; b += gridDim.x * warps_per_cta) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    float s[K], y[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = lane + 32 * k;
      s[k] = j < nb ? scores[base + j] : -INFINITY;
    }
    if (rel_bytes == 8) {
      const long long* r8 = reinterpret_cast<const long long*>(rel) + base;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        y[k] = j < nb ? static_cast<float>(r8[j]) : -INFINITY;
      }
    } else if (rel_bytes == 4) {
      const int* r4 = reinterpret_cast<const int*>(rel) + base;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        y[k] = j < nb ? static_cast<float>(r4[j]) : -INFINITY;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        y[k] = j < nb ? static_cast<float>(load_int_clamped(rel, rel_bytes, base + j)) : -INFINITY;
      }
    }
    float ms = -INFINITY, my = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) { ms = fmaxf(ms, s[k]); my = fmaxf(my, y[k]); }
    ms = warp_max(ms);
    my = warp_max(my);
    float zs = 0.0f, zy = 0.0f, a = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool valid = lane + 32 * k < nb;
      const float ds = valid ? s[k] - ms : 0.0f;
      const float es = valid ? ex2_approx(ds * kLog2e) : 0.0f;
      const float ey = valid ? ex2_approx((y[k] - my) * kLog2e) : 0.0f;
      s[k] = es;
      y[k] = ey;
      zs += es;
      zy += ey;
      a = fmaf(ey, ds, a);
    }
    zs = warp_sum(zs);
    zy = warp_sum(zy);
    a = warp_sum(a);
    const float loss = nb > 0 ? logf(zs) - a / zy : 0.0f;
    if (lane == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }
    if (grad_out) {
      const float izs = nb > 0 ? 1.0f / zs : 0.0f, izy = nb > 0 ? 1.0f / zy : 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        if (j < L) grad_out[base + j] = s[k] * izs - y[k] * izy;   // exactly 0 on padding
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Ranking metrics: dcg / ndcg (evaluation/dcg.py) and arp (evaluation/arp.py).
// ---------------------------------------------------------------------------------------

// Inclusive scan of v[0..L) in shared memory (Hillis-Steele, ping-pong between v and tmp).
// Returns the buffer holding the result.
__device__ __forceinline__ float* cta_inclusive_scan(float* v, float* tmp, int L) {
  float* src = v;
  float* dst = tmp;
  __syncthreads();
  for (int off = 1; off < L; off <<= 1) {
    for (int j = threadIdx.x; j < L; j += blockDim.x)
      dst[j] = j >= off ? src[j] + src[j - off] : src[j];
    __syncthreads();
    float* t = src; src = dst; dst = t;
  }
  return src;
}

// per-rank dcg terms of the ranking held in m.keys -> out[r] (dcg.py:85-93).  The relevance
// of padded documents is NOT masked, exactly as in the reference.
__device__ __forceinline__ void dcg_terms(const RowSmem& m, float* out, int L, int exp_gain) {
  for (int r = threadIdx.x; r < L; r += blockDim.x) {
    const int d = static_cast<int>(m.keys[r] & 0xffffffffu);
    float g = static_cast<float>(m.raw_y[d]);
    if (exp_gain) g = exp2f(g) - 1.0f;
    out[r] = g / log2f(static_cast<float>(r) + 2.0f);
  }
}

__global__ void __launch_bounds__(1024)
rank_metrics_kernel(int metric, const float* __restrict__ scores, const void* __restrict__ rel,
                    int rel_bytes, const void* __restrict__ n, int n_bytes, int B, int L, int P,
                    int k, int exp_gain, float* __restrict__ out, int out_ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RowSmem m = carve(smem_raw, L, P);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const int nb = load_n(n, n_bytes, b, L);
    stage_row(m, scores, rel, rel_bytes, b, L);
    __syncthreads();
    rank_row_by_score(m, nb, L, P);

    if (metric == LTR_METRIC_ARP) {
      // arp.py:32-42: sum_r (r + 1) rel_sort[r] / sum_r rel_sort[r] over valid ranks
      float srp = 0.0f, nrp = 0.0f;
      for (int r = threadIdx.x; r < nb; r += blockDim.x) {
        const float y = static_cast<float>(m.raw_y[static_cast<int>(m.keys[r] & 0xffffffffu)]);
        srp = fmaf(static_cast<float>(r + 1), y, srp);
        nrp += y;
      }
      srp = cta_sum(srp, m.red);
      nrp = cta_sum(nrp, m.red);
      if (nrp == 0.0f) nrp = 1.0f;
      if (threadIdx.x == 0) out[static_cast<size_t>(b) * out_ld] = srp / nrp;
      continue;
    }

    const int kk = k > 0 ? (k < L ? k : L) : 0;
    dcg_terms(m, m.s, L, exp_gain);
    if (kk > 0) {
      // dcg[:, :k][:, -1]: only the first kk ranks contribute
      float part = 0.0f;
      for (int r = threadIdx.x; r < kk; r += blockDim.x) part += m.s[r];
      __syncthreads();
      float v = cta_sum(part, m.red);
      if (metric == LTR_METRIC_NDCG) {
        rank_row_by_relevance(m, nb, L, P);
        dcg_terms(m, m.w, L, exp_gain);
        float ip = 0.0f;
        __syncthreads();
        for (int r = threadIdx.x; r < kk; r += blockDim.x) ip += m.w[r];
        float iv = cta_sum(ip, m.red);
        if (iv == 0.0f) iv = 1.0f;       // dcg.py:37
        v = v / iv;
      }
      if (threadIdx.x == 0) out[static_cast<size_t>(b) * out_ld] = v;
    } else {
      float* cum = cta_inclusive_scan(m.s, m.gout, L);   // dcg.py:94 cumsum
      float* icum = nullptr;
      if (metric == LTR_METRIC_NDCG) {
        rank_row_by_relevance(m, nb, L, P);
        dcg_terms(m, m.w, L, exp_gain);
        icum = cta_inclusive_scan(m.w, m.delta, L);
      }
      float* __restrict__ o = out + static_cast<size_t>(b) * out_ld;
      for (int r = threadIdx.x; r < L; r += blockDim.x) {
        float v = cum[r];
        if (icum) { float iv = icum[r]; if (iv == 0.0f) iv = 1.0f; v = v / iv; }
        o[r] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(1024)
rank_by_score_kernel(const float* __restrict__ scores, const void* __restrict__ n, int n_bytes,
                     int B, int L, int P, int64_t* __restrict__ ranking_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RowSmem m = carve(smem_raw, L, P);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    for (int j = threadIdx.x; j < L; j += blockDim.x) m.raw_s[j] = scores[base + j];
    __syncthreads();
    rank_row_by_score(m, nb, L, P);
    for (int r = threadIdx.x; r < L; r += blockDim.x)
      ranking_out[base + r] = static_cast<int64_t>(m.keys[r] & 0xffffffffu);
  }
}

// out[b, j] = g[b] * d[b, j]: the backward of every loss (HBM-bound streaming pass).
__global__ void __launch_bounds__(256)
scale_rows_kernel(const float* __restrict__ g, int g_stride, float g_scalar, const float* __restrict__ d,
                  float* __restrict__ out, int B, int L, int vec_ok) {
  const size_t total = static_cast<size_t>(B) * L;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (g == nullptr) {
    // one scalar for every row, passed by value (host callers: the broadcast gradient of .sum() / .mean())
    if (vec_ok) {
      const size_t total4 = total >> 2;
      const float4* __restrict__ d4 = reinterpret_cast<const float4*>(d);
      float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
      for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4; i += stride) {
        float4 v = d4[i];
        v.x *= g_scalar; v.y *= g_scalar; v.z *= g_scalar; v.w *= g_scalar;
        o4[i] = v;
      }
    } else {
      for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += stride)
        out[i] = g_scalar * d[i];
    }
    return;
  }
  if (vec_ok) {
    const size_t total4 = total >> 2;
    const int L4 = L >> 2;
    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(d);
    float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4; i += stride) {
      const float gb = g[(i / L4) * g_stride];
      float4 v = d4[i];
      v.x *= gb; v.y *= gb; v.z *= gb; v.w *= gb;
      o4[i] = v;
    }
  } else {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += stride)
      out[i] = g[(i / L) * g_stride] * d[i];
  }
}

// ---------------------------------------------------------------------------------------
// Host side of the C ABI
// ---------------------------------------------------------------------------------------
thread_local int tls_cuda_error = 0;

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

inline int check_common(const void* scores, const void* n, int n_bytes, int B, int L) {
  if (B < 0 || L < 1) return LTR_EINVAL;
  if (L > LTR_MAX_LIST_SIZE) return LTR_EUNSUPPORTED;
  if (B > 0 && (!scores || !n)) return LTR_EINVAL;
  if (n_bytes != 4 && n_bytes != 8) return LTR_EINVAL;
  return LTR_OK;
}

// relevance element widths: int64 (the reference's dtype), int32, int16, uint8
inline bool rel_width_ok(int bytes) { return bytes == 8 || bytes == 4 || bytes == 2 || bytes == 1; }

// TMA bulk row staging: 16-byte aligned rows of a multiple of 16 bytes, for scores and relevance alike
inline bool rows_tma_ok(const void* scores, const void* rel, int rel_bytes, int L) {
  return (L % 4 == 0) && ((static_cast<size_t>(rel_bytes) * L) % 16 == 0) && aligned16(scores) && aligned16(rel);
}

inline int cta_threads_for(int L) {
  int t = (L + 31) / 32 * 32;
  if (t < 64) t = 64;
  if (t > 1024) t = 1024;
  return t;
}

// Persistent grid: as many CTAs as fit on the device at once, never more than queries.
template <typename K>
inline int persistent_grid(K kernel, int threads, size_t smem, int B, const DeviceInfo& di, int* grid) {
  if (smem > 48 * 1024)
    LTR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) return LTR_EUNSUPPORTED;
  long long g = static_cast<long long>(per_sm) * di.sms;
  *grid = static_cast<int>(g < B ? g : B);
  return LTR_OK;
}

template <int MODE>
int launch_pair(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                int B, int L, float sigma, float* loss_out, float* grad_out, int64_t* ranking_out,
                float* loss_sum, cudaStream_t st, const DeviceInfo& di) {
  const int P = next_pow2(L);
  const int threads = cta_threads_for(L);
  const size_t smem = row_smem_bytes(L, P);
  int grid = 0;
  int rc = persistent_grid(pair_loss_kernel<MODE>, threads, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  pair_loss_kernel<MODE><<<grid, threads, smem, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, P, sigma,
                                                      loss_out, grad_out, ranking_out, loss_sum);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

// LTR_KERNEL=generic forces the generic one-CTA-per-query kernel (debugging / A-B timing).
inline bool force_generic() {
  const char* v = getenv("LTR_KERNEL");
  return v && strcmp(v, "generic") == 0;
}

// Work queues of the warp-per-query kernel: {next query, warps done} pairs in device memory.
// A launch takes the next slot round-robin; the kernel leaves its slot zeroed when its last
// warp retires, so a slot is clean again long before the rotation comes back to it.
constexpr int kQueueSlots = 1024;
__device__ unsigned int g_work_queues[2 * kQueueSlots];

inline int next_queue(cudaStream_t st, unsigned int** out) {
  static std::atomic<unsigned int> counter{0};
  // A launch that is being captured would bake its slot into the graph, and two graphs replayed
  // concurrently could then share a counter: captured launches without a caller-owned workspace take
  // the queries in a static stride instead (queue == nullptr).
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  LTR_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) {
    *out = nullptr;
    return LTR_OK;
  }
  unsigned int* base = nullptr;
  LTR_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_work_queues));
  *out = base + 2 * (counter.fetch_add(1, std::memory_order_relaxed) % kQueueSlots);
  return LTR_OK;
}

// Score-independent tables (delta windows, discounts, ideal-DCG prefix sums): filled on first use per
// device by a kernel on the caller's stream, followed by an event.  Nothing synchronises: until the
// event has completed, launches on any stream are made to wait for it (cudaStreamWaitEvent); a first
// use that is being captured into a CUDA graph puts the (idempotent) fill into that graph instead.
__device__ PairTables g_pair_tables;

struct TableState {
  std::atomic<int> state{0};   // 0: not filled, 1: fill enqueued and `ev` recorded behind it, 2: fill complete
  cudaEvent_t ev = nullptr;
  std::atomic_flag busy = ATOMIC_FLAG_INIT;
};

inline int pair_tables(cudaStream_t st, const PairTables** out) {
  static TableState states[64];
  int dev = 0;
  LTR_CUDA(cudaGetDevice(&dev));
  PairTables* t = nullptr;
  LTR_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&t), g_pair_tables));
  *out = t;
  if (dev < 0 || dev >= 64) {
    init_pair_tables_kernel<<<1, 1024, 0, st>>>(t);
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  TableState& ts = states[dev];
  if (ts.state.load(std::memory_order_acquire) == 2) return LTR_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  LTR_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) {
    init_pair_tables_kernel<<<1, 1024, 0, st>>>(t);
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  while (ts.busy.test_and_set(std::memory_order_acquire)) {}
  int rc = LTR_OK;
  const int state = ts.state.load(std::memory_order_acquire);
  cudaError_t e = cudaSuccess;
  if (state == 0) {
    init_pair_tables_kernel<<<1, 1024, 0, st>>>(t);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ts.ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ts.ev, st);
    if (e == cudaSuccess) ts.state.store(1, std::memory_order_release);
  } else if (state == 1) {
    const cudaError_t q = cudaEventQuery(ts.ev);
    if (q == cudaSuccess) ts.state.store(2, std::memory_order_release);
    else if (q == cudaErrorNotReady) e = cudaStreamWaitEvent(st, ts.ev, 0);
    else e = q;
  }
  ts.busy.clear(std::memory_order_release);
  if (e != cudaSuccess) rc = cuda_fail(e);
  return rc;
}

// ---- longest-first query schedule ------------------------------------------------------------------
// Pair counts grow with n^2, so the queries of one batch differ 4x and more in cost.  With a
// scheduling workspace the launch is preceded by a counting sort of the queries by decreasing n
// (one CTA: shared-memory histogram, scan, scatter); the loss kernels then start the long queries
// first and hand out the rest through a device-wide counter, which turns the tail of the launch
// into short queries (longest-processing-time-first list scheduling).
// Workspace layout: queue {next, done} + padding (16 bytes) | order uint32 [B].
constexpr int kOrderThreads = 1024;
__global__ void __launch_bounds__(kOrderThreads)
order_queries_kernel(const void* __restrict__ n, int n_bytes, int B, int L, unsigned int* __restrict__ queue,
                     unsigned int* __restrict__ order) {
  __shared__ unsigned int hist[LTR_MAX_LIST_SIZE + 1];
  __shared__ unsigned int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the first kOrderKeep values of every thread stay in registers for the scatter pass
  constexpr int kOrderKeep = 8;
  int keep[kOrderKeep];
#pragma unroll
  for (int r = 0; r < kOrderKeep; ++r) {
    const int i = tid + r * kOrderThreads;
    keep[r] = i < B ? load_n(n, n_bytes, i, L) : 0;
  }
  for (int i = tid; i <= L; i += kOrderThreads) hist[i] = 0u;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kOrderKeep; ++r)
    if (tid + r * kOrderThreads < B) atomicAdd(&hist[keep[r]], 1u);
  for (int i = tid + kOrderKeep * kOrderThreads; i < B; i += kOrderThreads)
    atomicAdd(&hist[load_n(n, n_bytes, i, L)], 1u);
  __syncthreads();
  // exclusive scan in order of DECREASING n: entry e stands for n = L - e
  const int K = (L + kOrderThreads) / kOrderThreads;   // ceil((L + 1) / threads)
  unsigned int local = 0u;
  for (int e = tid * K; e < (tid + 1) * K && e <= L; ++e) local += hist[L - e];
  unsigned int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  unsigned int run = incl - local + (warp > 0 ? wsum[warp - 1] : 0u);
  for (int e = tid * K; e < (tid + 1) * K && e <= L; ++e) {
    const unsigned int cnt = hist[L - e];
    hist[L - e] = run;
    run += cnt;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kOrderKeep; ++r) {
    const int i = tid + r * kOrderThreads;
    if (i < B) order[atomicAdd(&hist[keep[r]], 1u)] = static_cast<unsigned int>(i);
  }
  for (int i = tid + kOrderKeep * kOrderThreads; i < B; i += kOrderThreads) {
    const unsigned int pos = atomicAdd(&hist[load_n(n, n_bytes, i, L)], 1u);
    order[pos] = static_cast<unsigned int>(i);
  }
  if (tid == 0) { queue[0] = 0u; queue[1] = 0u; }
}

// Large batches: the same counting sort spread over many CTAs (the single-CTA form serialises 65536
// shared-memory atomics on ~250 distinct list sizes: 63 us at (65536, 512)).
//   order_hist_kernel    per-CTA shared histogram of its slice of n, added to the global histogram
//   order_scatter_kernel every CTA scans the global histogram (decreasing n) into shared memory, then
//                        places its queries: position = first slot of the size class + a global cursor
// Workspace: queue[4] | order[B] | hist[LTR_MAX_LIST_SIZE + 1] | cursor[LTR_MAX_LIST_SIZE + 1]; queue, hist
// and cursor are zeroed by one cudaMemsetAsync.  The order inside a size class depends on the atomics'
// arrival order: the schedule changes from run to run, the results (per-query, independent) do not.
constexpr int kOrderBins = LTR_MAX_LIST_SIZE + 1;
constexpr int kOrderParThreads = 256;
constexpr int kOrderParMinB = 16384;

__global__ void __launch_bounds__(kOrderParThreads)
order_hist_kernel(const void* __restrict__ n, int n_bytes, int B, int L, unsigned int* __restrict__ ghist) {
  __shared__ unsigned int hist[kOrderBins];
  for (int i = threadIdx.x; i <= L; i += kOrderParThreads) hist[i] = 0u;
  __syncthreads();
  for (int i = blockIdx.x * kOrderParThreads + threadIdx.x; i < B; i += gridDim.x * kOrderParThreads)
    atomicAdd(&hist[load_n(n, n_bytes, i, L)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i <= L; i += kOrderParThreads)
    if (hist[i]) atomicAdd(&ghist[i], hist[i]);
}

__global__ void __launch_bounds__(kOrderParThreads)
order_scatter_kernel(const void* __restrict__ n, int n_bytes, int B, int L, const unsigned int* __restrict__ ghist,
                     unsigned int* __restrict__ cursor, unsigned int* __restrict__ order) {
  __shared__ unsigned int first[kOrderBins];   // first[v] = number of queries with n > v
  __shared__ unsigned int wsum[kOrderParThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // exclusive scan in order of decreasing n: entry e stands for n = L - e
  const int K = (L + kOrderParThreads) / kOrderParThreads;
  unsigned int local = 0u;
  for (int e = tid * K; e < (tid + 1) * K && e <= L; ++e) local += ghist[L - e];
  unsigned int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = lane < kOrderParThreads / 32 ? wsum[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kOrderParThreads / 32) wsum[lane] = w;
  }
  __syncthreads();
  unsigned int run = incl - local + (warp > 0 ? wsum[warp - 1] : 0u);
  for (int e = tid * K; e < (tid + 1) * K && e <= L; ++e) {
    first[L - e] = run;
    run += ghist[L - e];
  }
  __syncthreads();
  for (int i = blockIdx.x * kOrderParThreads + tid; i < B; i += gridDim.x * kOrderParThreads) {
    const int v = load_n(n, n_bytes, i, L);
    order[first[v] + atomicAdd(&cursor[v], 1u)] = static_cast<unsigned int>(i);
  }
}

struct Schedule {
  unsigned int* queue;          // {next query, finished CTAs / warps}
  const unsigned int* order;    // queries by decreasing n, or nullptr (natural order)
};

inline size_t schedule_bytes(int B) {
  const size_t b = static_cast<size_t>(B > 0 ? B : 0);
  // queue | order [B] | (large batches) histogram + cursors of the multi-CTA counting sort
  return 16u + 4u * b + (B >= kOrderParMinB ? 2u * 4u * kOrderBins : 0u);
}

// `slots` = queries that start at once (resident CTAs / warps): below that nothing queues and the
// order is irrelevant.  LTR_SCHEDULE=natural disables the sort (A-B timing).
inline int make_schedule(const void* n, int n_bytes, int B, int L, long long slots, void* ws, size_t ws_bytes,
                         cudaStream_t st, Schedule* out) {
  out->order = nullptr;
  static const bool natural = [] {
    const char* v = getenv("LTR_SCHEDULE");
    return v && strcmp(v, "natural") == 0;
  }();
  static const bool always = [] {
    const char* v = getenv("LTR_SCHEDULE");
    return v && strcmp(v, "always") == 0;
  }();
  if (ws && ws_bytes >= schedule_bytes(B) && (B > slots || always) && !natural &&
      (reinterpret_cast<uintptr_t>(ws) & 15u) == 0) {
    unsigned int* q = static_cast<unsigned int*>(ws);
    if (B >= kOrderParMinB) {
      unsigned int* hist = q + 4 + B;
      unsigned int* cursor = hist + kOrderBins;
      LTR_CUDA(cudaMemsetAsync(q, 0, 16, st));
      LTR_CUDA(cudaMemsetAsync(hist, 0, 2u * 4u * kOrderBins, st));
      DeviceInfo di;
      int rc = device_info(&di);
      if (rc != LTR_OK) return rc;
      const int want = (B + kOrderParThreads * 4 - 1) / (kOrderParThreads * 4);
      const int grid = want < di.sms ? want : di.sms;
      order_hist_kernel<<<grid, kOrderParThreads, 0, st>>>(n, n_bytes, B, L, hist);
      LTR_CUDA(cudaGetLastError());
      order_scatter_kernel<<<grid, kOrderParThreads, 0, st>>>(n, n_bytes, B, L, hist, cursor, q + 4);
      LTR_CUDA(cudaGetLastError());
    } else {
      order_queries_kernel<<<1, kOrderThreads, 0, st>>>(n, n_bytes, B, L, q, q + 4);
      LTR_CUDA(cudaGetLastError());
    }
    out->queue = q;
    out->order = q + 4;
    return LTR_OK;
  }
  return next_queue(st, &out->queue);
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <int TW>
int launch_pair_warp(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                     int B, int L, float sigma, int dcg_mod, float* loss_out, float* grad_out, int64_t* ranking_out,
                     float* loss_sum, void* ws, size_t ws_bytes, cudaStream_t st, const DeviceInfo& di) {
  const int threads = kWarpsPerCta * 32;
  int per_sm = 0;
  LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_warp_kernel<TW>, threads, 0));
  if (per_sm < 1) return LTR_EUNSUPPORTED;
  // LTR_WARP_CTAS caps the resident CTAs per SM so that more of the batch is handed out
  // dynamically (measured on (4096, 128): 26 us all-resident vs 34 - 39 us with 3 - 6 CTAs per
  // SM, so the default keeps every warp slot busy and queues only what does not fit).
  static const int sched_ctas = env_int("LTR_WARP_CTAS", kWarpSchedCtas);
  const bool can_order = ws && ws_bytes >= schedule_bytes(B);
  if (can_order && per_sm > sched_ctas && sched_ctas > 0) per_sm = sched_ctas;
  const long long want = (static_cast<long long>(B) + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long cap = static_cast<long long>(per_sm) * di.sms;
  const int grid = static_cast<int>(want < cap ? want : cap);
  const int vec_ok = (L % 4 == 0) && aligned16(scores) && aligned16(rel) && (!grad_out || aligned16(grad_out));
  Schedule sc;
  int rc = make_schedule(n, n_bytes, B, L, static_cast<long long>(grid) * kWarpsPerCta, ws, ws_bytes, st, &sc);
  if (rc != LTR_OK) return rc;
  const PairTables* tabs = nullptr;
  rc = pair_tables(st, &tabs);
  if (rc != LTR_OK) return rc;
  pair_warp_kernel<TW><<<grid, threads, 0, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, vec_ok,
                                                 dcg_mod, loss_out, grad_out, ranking_out, loss_sum, sc.queue,
                                                 sc.order, tabs);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

template <int TW>
int launch_pair_cta(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                    int B, int L, float sigma, int dcg_mod, float* loss_out, float* grad_out, int64_t* ranking_out,
                    float* loss_sum, void*, size_t, cudaStream_t st, const DeviceInfo& di) {
  const int P = next_pow2(L);
  const int threads = kCtaWarps * 32;
  // TMA bulk staging needs 16-byte aligned rows of a multiple of 16 bytes, and room for the
  // double buffer next to everything else
  int tma = rows_tma_ok(scores, rel, rel_bytes, L) && cta_smem_bytes(L, P, rel_bytes) <= 200u * 1024u;
  if (const char* v = getenv("LTR_TMA")) tma = tma && strcmp(v, "0") != 0;
  const size_t smem = cta_smem_bytes(L, P, tma ? rel_bytes : 0);
  int grid = 0;
  int rc = persistent_grid(pair_cta_kernel<TW>, threads, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  const PairTables* tabs = nullptr;
  rc = pair_tables(st, &tabs);
  if (rc != LTR_OK) return rc;
  pair_cta_kernel<TW><<<grid, threads, smem, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, P, sigma, dcg_mod,
                                                   tma, loss_out, grad_out, ranking_out, loss_sum, tabs);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

// LTR_KERNEL=tiles keeps the 128 x 128 rank-tile kernel for every L > 128 (A-B timing).
inline bool force_tiles() {
  const char* v = getenv("LTR_KERNEL");
  return v && strcmp(v, "tiles") == 0;
}

template <int TW, int W>
int launch_pair_ring_w(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                       int B, int L, float sigma, int dcg_mod, float* loss_out, float* grad_out, int64_t* ranking_out,
                       float* loss_sum, void* ws, size_t ws_bytes, cudaStream_t st, const DeviceInfo& di) {
  const int P = next_pow2(L);
  const int threads = W * 32;
  // TMA bulk staging needs 16-byte aligned rows of a multiple of 16 bytes
  int tma = rows_tma_ok(scores, rel, rel_bytes, L);
  if (const char* v = getenv("LTR_TMA")) tma = tma && strcmp(v, "0") != 0;
  const size_t smem = ring_smem_bytes(L, tma ? rel_bytes : 0, W);
  int grid = 0;
  int rc = persistent_grid(pair_ring_kernel<TW, W>, threads, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  Schedule sc;
  rc = make_schedule(n, n_bytes, B, L, grid, ws, ws_bytes, st, &sc);
  if (rc != LTR_OK) return rc;
  const PairTables* tabs = nullptr;
  rc = pair_tables(st, &tabs);
  if (rc != LTR_OK) return rc;
  pair_ring_kernel<TW, W><<<grid, threads, smem, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, P, sigma, dcg_mod,
                                                       tma, loss_out, grad_out, ranking_out, loss_sum, sc.queue,
                                                       sc.order, tabs);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

template <int TW>
int launch_pair_ring(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                     int B, int L, float sigma, int dcg_mod, float* loss_out, float* grad_out, int64_t* ranking_out,
                     float* loss_sum, void* ws, size_t ws_bytes, cudaStream_t st, const DeviceInfo& di) {
  // LTR_RING_WARPS=8 keeps eight warps per query for every list size (A-B timing)
  // (2 warps per query up to 256 documents: 8-10 % over 4 warps for the sigmoid losses at 132..256 documents,
  // tools/ring_warps_ab.py)
  static const int forced = env_int("LTR_RING_WARPS", 0);
  const bool short_lists = forced ? forced == kRingWarpsShort : L <= kRingShortL;
  // one warp per query for the losses that need no ranking (logistic, ARP, hinge): no CTA barrier left, up to 29 %
  // at 132 documents, 5 % at 256; the NDCG losses keep two warps (the 256-key sort wants both)
  const bool ranked = TW == TW_DELTA || (TW == TW_TWO && dcg_mod != 0) || ranking_out != nullptr;
  if ((forced == 1 || (forced == 0 && !ranked)) && L <= kRingTinyL)
    return launch_pair_ring_w<TW, 1>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, dcg_mod, loss_out, grad_out,
                                     ranking_out, loss_sum, ws, ws_bytes, st, di);
  if ((forced == 2 || forced == 0) && L <= kRingTinyL)
    return launch_pair_ring_w<TW, 2>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, dcg_mod, loss_out, grad_out,
                                     ranking_out, loss_sum, ws, ws_bytes, st, di);
  if (short_lists)
    return launch_pair_ring_w<TW, kRingWarpsShort>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, dcg_mod, loss_out,
                                                   grad_out, ranking_out, loss_sum, ws, ws_bytes, st, di);
  return launch_pair_ring_w<TW, kRingWarpsLong>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, dcg_mod, loss_out,
                                                grad_out, ranking_out, loss_sum, ws, ws_bytes, st, di);
}

// PairwiseHingeLoss / PairwiseDCGHingeLoss for long lists: O(n log n) by sorting (ltr_hinge_sorted.cuh).  The
// sorted kernel pays ~17 ns per query whatever the list size (one 256-thread CTA, a dozen barriers), the pair
// kernels ~0.4 ns per pair: measured crossover between 320 and 400 documents (tools/hinge_sweep.py: 8192 x 200
// takes 86 us by pairs against 145 us sorted; 1024 x 1024 takes 155 us against 57 us).  Same gradients bit
// for bit either way.  LTR_HINGE=pairs / sorted forces one form for every list above 128 (A-B timing, cross-check).
constexpr int kHingeSortMinL = 384;
inline int hinge_form() {            // 0 = by list size, 1 = pairs, 2 = sorted
  const char* v = getenv("LTR_HINGE");
  if (!v) return 0;
  return strcmp(v, "pairs") == 0 ? 1 : (strcmp(v, "sorted") == 0 ? 2 : 0);
}
inline bool hinge_sorted_for(int L) {
  const int f = hinge_form();
  return f == 2 || (f == 0 && L > kHingeSortMinL);
}

constexpr int kHingeMaxL = 4096;

template <int THREADS>
int launch_hinge_sorted_t(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes, int B,
                          int L, int dcg_mod, float* loss_out, float* grad_out, float* loss_sum, cudaStream_t st,
                          const DeviceInfo& di) {
  const int P = next_pow2(L);
  const size_t smem = hinge_smem_bytes(L, P);
  int grid = 0;
  int rc = persistent_grid(hinge_sorted_kernel<THREADS>, THREADS, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  hinge_sorted_kernel<THREADS><<<grid, THREADS, smem, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, P, dcg_mod,
                                                            loss_out, grad_out, loss_sum);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int launch_hinge_sorted(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes, int B,
                        int L, int dcg_mod, float* loss_out, float* grad_out, float* loss_sum, cudaStream_t st,
                        const DeviceInfo& di) {
  if (L <= 1024)
    return launch_hinge_sorted_t<256>(scores, rel, rel_bytes, n, n_bytes, B, L, dcg_mod, loss_out, grad_out,
                                      loss_sum, st, di);
  return launch_hinge_sorted_t<1024>(scores, rel, rel_bytes, n, n_bytes, B, L, dcg_mod, loss_out, grad_out, loss_sum,
                                     st, di);
}

int dispatch_pair(int pm, const float* scores, const void* rel, int rel_bytes, const void* n,
                  int n_bytes, int B, int L, float sigma, float* loss_out, float* grad_out,
                  int64_t* ranking_out, float* loss_sum, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_common(scores, n, n_bytes, B, L);
  if (rc != LTR_OK) return rc;
  if (!rel_width_ok(rel_bytes)) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!rel || !loss_out) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!force_generic()) {
    // every unordered pair once: one warp per query for short lists, one CTA per query on a
    // query-wide chunk ring up to 1024 documents, 128 x 128 rank tiles beyond
    const bool ring = L <= kRingMaxL && !force_tiles();
    if ((pm == PM_HINGE || pm == PM_DCG_HINGE) && L > kWarpL && L <= kHingeMaxL && !ranking_out &&
        !force_tiles() && hinge_sorted_for(L))
      return launch_hinge_sorted(scores, rel, rel_bytes, n, n_bytes, B, L, pm == PM_DCG_HINGE ? 1 : 0, loss_out,
                                 grad_out, loss_sum, st, di);
#define LTR_TILED(TWMODE, DCG)                                                                          \
  return L <= kWarpL                                                                                    \
             ? launch_pair_warp<TWMODE>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, DCG, loss_out,  \
                                        grad_out, ranking_out, loss_sum, ws, ws_bytes, st, di)          \
         : ring                                                                                         \
             ? launch_pair_ring<TWMODE>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, DCG, loss_out,  \
                                        grad_out, ranking_out, loss_sum, ws, ws_bytes, st, di)          \
             : launch_pair_cta<TWMODE>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, DCG, loss_out,   \
                                       grad_out, ranking_out, loss_sum, ws, ws_bytes, st, di)
    if (pm == PM_LOGISTIC) LTR_TILED(TW_UNIT, 0);
    if (pm == PM_ARP2) LTR_TILED(TW_DIFF, 0);
    if (pm == PM_NDCG2) LTR_TILED(TW_DELTA, 0);
    if (pm == PM_HINGE) LTR_TILED(TW_HINGE, 0);
    if (pm == PM_DCG_HINGE) LTR_TILED(TW_HINGE, 1);
    if (pm == PM_ARP1) LTR_TILED(TW_TWO, 0);
    if (pm == PM_NDCG1) LTR_TILED(TW_TWO, 1);
#undef LTR_TILED
  }
#define LTR_CASE(M)                                                                               \
  case M:                                                                                         \
    return launch_pair<M>(scores, rel, rel_bytes, n, n_bytes, B, L, sigma, loss_out, grad_out,    \
                          ranking_out, loss_sum, st, di)
  switch (pm) {
    LTR_CASE(PM_HINGE);
    LTR_CASE(PM_DCG_HINGE);
    LTR_CASE(PM_LOGISTIC);
    LTR_CASE(PM_ARP1);
    LTR_CASE(PM_ARP2);
    LTR_CASE(PM_NDCG1);
    LTR_CASE(PM_NDCG2);
    default:
      return LTR_EINVAL;
  }
#undef LTR_CASE
}

}  // namespace ltr

using namespace ltr;

extern "C" {

int ltr_version(void) { return LTR_VERSION; }

const char* ltr_strerror(int rc) {
  switch (rc) {
    case LTR_OK: return "success";
    case LTR_EINVAL: return "invalid argument";
    case LTR_EUNSUPPORTED: return "unsupported list size or device (needs sm_100, L <= LTR_MAX_LIST_SIZE)";
    case LTR_ECUDA: return cudaGetErrorString(static_cast<cudaError_t>(tls_cuda_error));
    default: return "unknown error";
  }
}

int ltr_last_cuda_error(void) { return tls_cuda_error; }

int ltr_pairwise_additive(int mode, const float* scores, const void* rel, int rel_bytes,
                          const void* n, int n_bytes, int B, int L, float sigma, float* loss_out,
                          float* dscores_out, float* loss_sum, void* stream) {
  if (mode < LTR_ADD_HINGE || mode > LTR_ADD_LOGISTIC) return LTR_EINVAL;
  return dispatch_pair(PM_HINGE + mode, scores, rel, rel_bytes, n, n_bytes, B, L, sigma, loss_out,
                       dscores_out, nullptr, loss_sum, nullptr, 0, stream);
}

int ltr_pairwise_additive_ws(int mode, const float* scores, const void* rel, int rel_bytes,
                             const void* n, int n_bytes, int B, int L, float sigma, float* loss_out,
                             float* dscores_out, float* loss_sum, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (mode < LTR_ADD_HINGE || mode > LTR_ADD_LOGISTIC) return LTR_EINVAL;
  return dispatch_pair(PM_HINGE + mode, scores, rel, rel_bytes, n, n_bytes, B, L, sigma, loss_out,
                       dscores_out, nullptr, loss_sum, workspace, workspace_bytes, stream);
}

size_t ltr_schedule_workspace_bytes(int B) { return schedule_bytes(B); }

int ltr_lambda(int mode, const float* scores, const void* rel, int rel_bytes, const void* n,
               int n_bytes, int B, int L, float sigma, float* loss_out, float* dscores_out,
               int64_t* ranking_out, float* loss_sum, void* stream) {
  if (mode < LTR_LAM_ARP1 || mode > LTR_LAM_NDCG2) return LTR_EINVAL;
  return dispatch_pair(PM_ARP1 + mode, scores, rel, rel_bytes, n, n_bytes, B, L, sigma, loss_out,
                       dscores_out, ranking_out, loss_sum, nullptr, 0, stream);
}

int ltr_lambda_ws(int mode, const float* scores, const void* rel, int rel_bytes, const void* n,
                  int n_bytes, int B, int L, float sigma, float* loss_out, float* dscores_out,
                  int64_t* ranking_out, float* loss_sum, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (mode < LTR_LAM_ARP1 || mode > LTR_LAM_NDCG2) return LTR_EINVAL;
  return dispatch_pair(PM_ARP1 + mode, scores, rel, rel_bytes, n, n_bytes, B, L, sigma, loss_out,
                       dscores_out, ranking_out, loss_sum, workspace, workspace_bytes, stream);
}

int ltr_listnet(const float* scores, const void* rel, int rel_bytes, const void* n, int n_bytes,
                int B, int L, float* loss_out, float* dscores_out, float* loss_sum, void* stream) {
  int rc = check_common(scores, n, n_bytes, B, L);
  if (rc != LTR_OK) return rc;
  if (!rel_width_ok(rel_bytes)) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!rel || !loss_out) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  const int threads = 256, wpc = threads / 32;
  long long want = (static_cast<long long>(B) + wpc - 1) / wpc;
  long long cap = static_cast<long long>(di.sms) * 8;
  const int grid = static_cast<int>(want < cap ? want : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LTR_LISTNET(K)                                                                              \
  listnet_reg_kernel<K><<<grid, threads, 0, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, loss_out, \
                                                  dscores_out, loss_sum)
  if (L <= 32) LTR_LISTNET(1);
  else if (L <= 64) LTR_LISTNET(2);
  else if (L <= 128) LTR_LISTNET(4);
  else if (L <= 256) LTR_LISTNET(8);
  else if (L <= 512) LTR_LISTNET(16);
  else if (L <= 1024) LTR_LISTNET(32);
  else
    listnet_kernel<<<grid, threads, 0, st>>>(scores, rel, rel_bytes, n, n_bytes, B, L, loss_out,
                                             dscores_out, loss_sum);
#undef LTR_LISTNET
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_rank_metrics(int metric, const float* scores, const void* rel, int rel_bytes,
                     const void* n, int n_bytes, int B, int L, int k, int exp_gain, float* out,
                     int out_ld, void* stream) {
  if (metric < LTR_METRIC_DCG || metric > LTR_METRIC_ARP) return LTR_EINVAL;
  int rc = check_common(scores, n, n_bytes, B, L);
  if (rc != LTR_OK) return rc;
  if (!rel_width_ok(rel_bytes)) return LTR_EINVAL;
  if (k < 0) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!rel || !out) return LTR_EINVAL;
  const bool all_k = metric != LTR_METRIC_ARP && k == 0;
  if (out_ld < (all_k ? L : 1)) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  static const bool no_topk = [] {
    const char* v = getenv("LTR_TOPK");
    return v && strcmp(v, "0") == 0;
  }();
  if (metric != LTR_METRIC_ARP && k > 0 && k <= 32 && L <= 1024 && !force_generic() && !no_topk) {
    // dcg@k / ndcg@k with a small cut-off: top-k selection, one warp per query, no full sort
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PairTables* tabs = nullptr;
    rc = pair_tables(st, &tabs);
    if (rc != LTR_OK) return rc;
    const long long want = (static_cast<long long>(B) + kTopkWarps - 1) / kTopkWarps;
    // TMA bulk staging needs 16-byte aligned rows of a multiple of 16 bytes
    int tma = rows_tma_ok(scores, rel, rel_bytes, L);
    if (const char* v = getenv("LTR_TMA")) tma = tma && strcmp(v, "0") != 0;
#define LTR_TOPK_LAUNCH(E)                                                                                  \
  do {                                                                                                      \
    int per_sm = 0;                                                                                         \
    LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, topk_metrics_warp_kernel<E>,            \
                                                           kTopkWarps * 32, 0));                            \
    if (per_sm < 1) return LTR_EUNSUPPORTED;                                                                \
    const long long cap = static_cast<long long>(per_sm) * di.sms;                                          \
    /* every warp the same number of queries: no second partial wave */                                     \
    const long long rounds = (want + cap - 1) / cap;                                                        \
    const int grid = static_cast<int>((want + rounds - 1) / rounds);                                        \
    topk_metrics_warp_kernel<E><<<grid, kTopkWarps * 32, 0, st>>>(metric, scores, rel, rel_bytes, n, n_bytes, B, \
                                                                  L, k, exp_gain, tma, out, out_ld, tabs);  \
  } while (0)
    // E = documents per lane: the smallest instantiated size that covers L (the kernel walks the
    // row in strides of 32, so E need not be a power of two)
    if (L <= 32) LTR_TOPK_LAUNCH(1);
    else if (L <= 64) LTR_TOPK_LAUNCH(2);
    else if (L <= 96) LTR_TOPK_LAUNCH(3);
    else if (L <= 128) LTR_TOPK_LAUNCH(4);
    else if (L <= 160) LTR_TOPK_LAUNCH(5);
    else if (L <= 192) LTR_TOPK_LAUNCH(6);
    else if (L <= 224) LTR_TOPK_LAUNCH(7);
    else if (L <= 256) LTR_TOPK_LAUNCH(8);
    else if (L <= 384) LTR_TOPK_LAUNCH(12);
    else if (L <= 512) LTR_TOPK_LAUNCH(16);
    else if (L <= 768) LTR_TOPK_LAUNCH(24);
    else LTR_TOPK_LAUNCH(32);
#undef LTR_TOPK_LAUNCH
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  if (L <= 1024 && !force_generic()) {
    // one warp per query, in-register ranking (E = 1 .. 32 documents per lane), double-buffered TMA row staging.
    // Up to 1024 documents since round 2: the CTA-per-query kernel behind it costs 8x more per query at 260
    // documents than this one at 256 (tools/metric_sweep.py)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PairTables* tabs = nullptr;
    rc = pair_tables(st, &tabs);
    if (rc != LTR_OK) return rc;
    int tma = rows_tma_ok(scores, rel, rel_bytes, L);
    if (const char* v = getenv("LTR_TMA")) tma = tma && strcmp(v, "0") != 0;
#define LTR_RANKM_LAUNCH(E, WPB)                                                                               \
  do {                                                                                                        \
    int per_sm = 0;                                                                                           \
    const long long want_ctas = (static_cast<long long>(B) + (WPB) - 1) / (WPB);                              \
    LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rank_metrics_warp_kernel<E, WPB>,         \
                                                           (WPB) * 32, 0));                                   \
    if (per_sm < 1) return LTR_EUNSUPPORTED;                                                                  \
    const long long cap = static_cast<long long>(per_sm) * di.sms;                                            \
    /* every warp the same number of queries: no second partial wave */                                       \
    const long long rounds = (want_ctas + cap - 1) / cap;                                                     \
    const int grid = static_cast<int>((want_ctas + rounds - 1) / rounds);                                     \
    rank_metrics_warp_kernel<E, WPB><<<grid, (WPB) * 32, 0, st>>>(metric, scores, rel, rel_bytes, n, n_bytes, B, L, \
                                                                  k, exp_gain, tma, out, out_ld, tabs);       \
  } while (0)
    if (L <= 32) LTR_RANKM_LAUNCH(1, 4);
    else if (L <= 64) LTR_RANKM_LAUNCH(2, 4);
    else if (L <= 128) LTR_RANKM_LAUNCH(4, 4);
    else if (L <= 256) LTR_RANKM_LAUNCH(8, 4);
    else if (L <= 512) LTR_RANKM_LAUNCH(16, 2);
    else LTR_RANKM_LAUNCH(32, 1);
#undef LTR_RANKM_LAUNCH
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  const int P = next_pow2(L);
  const int threads = cta_threads_for(L);
  const size_t smem = row_smem_bytes(L, P);
  int grid = 0;
  rc = persistent_grid(rank_metrics_kernel, threads, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  rank_metrics_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      metric, scores, rel, rel_bytes, n, n_bytes, B, L, P, k, exp_gain, out, out_ld);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_rank_by_score(const float* scores, const void* n, int n_bytes, int B, int L,
                      int64_t* ranking_out, void* stream) {
  int rc = check_common(scores, n, n_bytes, B, L);
  if (rc != LTR_OK) return rc;
  if (B == 0) return LTR_OK;
  if (!ranking_out) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  if (L <= 1024 && !force_generic()) {
    // one warp per query up to 1024 documents (16 / 32 per lane beyond 256: the CTA-per-query kernel behind
    // costs 8x more per query at 260 documents than this one at 256)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LTR_RBS_LAUNCH(E, WPB)                                                                              \
  do {                                                                                                     \
    const long long want = (static_cast<long long>(B) + (WPB) - 1) / (WPB);                                \
    const long long cap = static_cast<long long>(di.sms) * 12;                                             \
    const int grid = static_cast<int>(want < cap ? want : cap);                                            \
    rank_by_score_warp_kernel<E, WPB><<<grid, (WPB) * 32, 0, st>>>(scores, n, n_bytes, B, L, ranking_out); \
  } while (0)
    if (L <= 128) LTR_RBS_LAUNCH(4, 4);
    else if (L <= 256) LTR_RBS_LAUNCH(8, 4);
    else if (L <= 512) LTR_RBS_LAUNCH(16, 2);
    else LTR_RBS_LAUNCH(32, 1);
#undef LTR_RBS_LAUNCH
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  const int P = next_pow2(L);
  const int threads = cta_threads_for(L);
  const size_t smem = row_smem_bytes(L, P);
  int grid = 0;
  rc = persistent_grid(rank_by_score_kernel, threads, smem, B, di, &grid);
  if (rc != LTR_OK) return rc;
  rank_by_score_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      scores, n, n_bytes, B, L, P, ranking_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

static int launch_scale_rows(const float* g, int g_stride, float g_scalar, const float* dscores, float* out, int B,
                             int L, cudaStream_t st) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  const size_t total = static_cast<size_t>(B) * L;
  // 128-bit path: rows of a multiple of 4 floats in 16-byte aligned buffers
  const int vec_ok = (L & 3) == 0 && aligned16(dscores) && aligned16(out);
  const size_t work = vec_ok ? total / 4 : total;
  long long want = static_cast<long long>((work + 255) / 256);
  long long cap = static_cast<long long>(di.sms) * 8;
  const int grid = static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
  scale_rows_kernel<<<grid, 256, 0, st>>>(g, g_stride, g_scalar, dscores, out, B, L, vec_ok);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_scale_rows(const float* g, int g_stride, const float* dscores, float* out, int B, int L,
                   void* stream) {
  if (B < 0 || L < 1 || (g_stride != 0 && g_stride != 1)) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!g || !dscores || !out) return LTR_EINVAL;
  return launch_scale_rows(g, g_stride, 0.0f, dscores, out, B, L, static_cast<cudaStream_t>(stream));
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// Position-biased click model (click_simulation/pbm.py:12-63): the document at rank r of query b
// is observed with probability 1 / (2 + r)^eta if r < min(n, cutoff) (else 0) and, once observed,
// clicked with probability relevance_probs[grade].  Written in DOCUMENT order (the reference inverts
// the ranking with a second argsort, :55-62); the Bernoulli draw stays with the caller's generator.
__global__ void __launch_bounds__(256)
pbm_probabilities_kernel(const int64_t* __restrict__ rankings, const void* __restrict__ ys, int ys_bytes,
                         const void* __restrict__ n, int n_bytes, const float* __restrict__ relevance_probs,
                         int n_probs, int cutoff, float eta, int B, int L, float* __restrict__ click_prob_out,
                         float* __restrict__ propensity_out) {
  const size_t total = static_cast<size_t>(B) * L;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / L), r = static_cast<int>(i - static_cast<size_t>(b) * L);
    int nb = load_n(n, n_bytes, b, L);
    if (cutoff > 0 && cutoff < nb) nb = cutoff;                       // :33-34
    long long d = rankings[i];
    d = d < 0 ? 0 : (d >= L ? L - 1 : d);
    const float obs = r < nb ? 1.0f / powf(2.0f + static_cast<float>(r), eta) : 0.0f;   // :37-42
    int y = load_int_clamped(ys, ys_bytes, static_cast<size_t>(b) * L + d);
    y = y < 0 ? 0 : (y >= n_probs ? n_probs - 1 : y);
    const size_t o = static_cast<size_t>(b) * L + d;
    propensity_out[o] = obs;
    click_prob_out[o] = relevance_probs[y] * obs;                     // :45-53
  }
}

int ltr_pbm_probabilities(const int64_t* rankings, const void* ys, int ys_bytes, const void* n, int n_bytes,
                          const float* relevance_probs, int n_probs, int cutoff, float eta, int B, int L,
                          float* click_prob_out, float* propensity_out, void* stream) {
  // elementwise kernel: any list size (no LTR_MAX_LIST_SIZE limit here)
  if (B < 0 || L < 1 || (n_bytes != 4 && n_bytes != 8)) return LTR_EINVAL;
  if (B > 0 && (!rankings || !n)) return LTR_EINVAL;
  int rc = LTR_OK;
  if (!rel_width_ok(ys_bytes) || n_probs < 1 || cutoff < 0) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!ys || !relevance_probs || !click_prob_out || !propensity_out) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  const size_t total = static_cast<size_t>(B) * L;
  long long want = static_cast<long long>((total + 255) / 256);
  const long long cap = static_cast<long long>(di.sms) * 8;
  const int grid = static_cast<int>(want < cap ? want : cap);
  pbm_probabilities_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rankings, ys, ys_bytes, n, n_bytes, relevance_probs, n_probs, cutoff, eta, B, L, click_prob_out,
      propensity_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_collate(const float* features, const int64_t* relevance, const int64_t* offsets, const int64_t* qidx,
                int B, int L, int F, float* feat_out, int64_t* rel_out, int64_t* n_out, int64_t* count_out,
                void* stream) {
  if (B < 0 || L < 1 || F < 1) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!features || !relevance || !offsets || !qidx || !feat_out || !rel_out || !n_out) return LTR_EINVAL;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  const size_t row_floats = static_cast<size_t>(L) * F;
  const long long slabs = static_cast<long long>((row_floats + kCollateSlab - 1) / kCollateSlab);
  const long long grid = slabs * B;
  if (grid > 2147483647LL) return LTR_EUNSUPPORTED;
  const int vec_ok = (F % 4 == 0) && aligned16(features) && aligned16(feat_out);
  collate_kernel<<<static_cast<unsigned int>(grid), kCollateThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      features, relevance, offsets, qidx, B, L, F, static_cast<int>(slabs), vec_ok, feat_out, rel_out, n_out,
      count_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_collate_sampled(const float* features, const int64_t* relevance, const int64_t* offsets,
                        const int64_t* qidx, const int64_t* sel, int sel_ld, int B, int L, int F, float* feat_out,
                        int64_t* rel_out, int64_t* n_out, int64_t* count_out, void* stream) {
  if (B < 0 || L < 1 || F < 1 || sel_ld < 0) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!features || !relevance || !offsets || !qidx || !feat_out || !rel_out || !n_out) return LTR_EINVAL;
  if (!sel && sel_ld != 0) return LTR_EINVAL;
  if (sel && sel_ld < L) return LTR_EINVAL;
  const long long groups = (L + kGatherRowsPerCta - 1) / kGatherRowsPerCta;
  const long long grid = groups * B;
  if (grid > 2147483647LL) return LTR_EUNSUPPORTED;
  const int vec_ok = (F % 4 == 0) && aligned16(features) && aligned16(feat_out);
  collate_gather_kernel<<<static_cast<unsigned int>(grid), kCollateThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      features, relevance, offsets, qidx, sel, sel_ld, B, L, F, static_cast<int>(groups), vec_ok, feat_out, rel_out,
      n_out, count_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_collate_sparse(const int64_t* indptr, const int64_t* indices, const float* values, const int64_t* relevance,
                       const int64_t* offsets, const int64_t* qidx, const int64_t* sel, int sel_ld,
                       const int64_t* out_ptr, int B, int L, int64_t nnz_out, int64_t* coo_out, float* val_out,
                       int64_t* rel_out, int64_t* n_out, void* stream) {
  if (B < 0 || L < 1 || nnz_out < 0 || sel_ld < 0) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!indptr || !relevance || !offsets || !qidx || !out_ptr || !rel_out || !n_out) return LTR_EINVAL;
  if (nnz_out > 0 && (!indices || !values || !coo_out || !val_out)) return LTR_EINVAL;
  if (sel && sel_ld < L) return LTR_EINVAL;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  const long long rows = static_cast<long long>(B) * L;
  const long long want = (rows + kCollateThreads / 32 - 1) / (kCollateThreads / 32);
  const long long cap = static_cast<long long>(di.sms) * 8;
  const int grid = static_cast<int>(want < cap ? want : cap);
  collate_sparse_kernel<<<grid, kCollateThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      indptr, indices, values, relevance, offsets, qidx, sel, sel_ld, out_ptr, B, L, nnz_out, coo_out, val_out, rel_out,
      n_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

size_t ltr_linear_listnet_workspace_bytes(int F) {
  return F < 1 ? 0 : static_cast<size_t>(kFusedBwdCtas) * (static_cast<size_t>(F) + 1) * sizeof(float);
}

int ltr_linear_listnet(const float* features, const float* weight, const float* bias, const void* rel,
                       int rel_bytes, const void* n, int n_bytes, int B, int L, int F, float* scores_out,
                       float* loss_out, float* dscores_out, float* qgrad_out, float* loss_sum, void* stream) {
  int rc = check_common(features, n, n_bytes, B, L);
  if (rc != LTR_OK) return rc;
  if (!rel_width_ok(rel_bytes)) return LTR_EINVAL;
  if (F < 1) return LTR_EINVAL;
  if (B > 0 && (!weight || !rel || !loss_out || !qgrad_out)) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  if (B == 0) return LTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t max_smem = 227u * 1024u;
  // The fast kernel keeps a whole L x F block in shared memory and moves it by TMA bulk copies; every
  // other shape (F % 4 != 0, unaligned buffers, blocks larger than shared memory) takes the tiled kernel.
  const bool fast_ok = F % 4 == 0 && F <= kFusedMaxF && (static_cast<size_t>(rel_bytes) * L) % 16 == 0 &&
                       aligned16(features) && aligned16(rel) && fused_smem_bytes(L, F, rel_bytes, 1) <= max_smem;
  static const bool force_tiled = [] {
    const char* v = getenv("LTR_FUSED");
    return v && strcmp(v, "tiled") == 0;
  }();
  if (!fast_ok || force_tiled) {
    const size_t smem = fused_tiled_smem_bytes(L, F);
    if (smem > max_smem) return LTR_EUNSUPPORTED;
    if (smem > 48 * 1024)
      LTR_CUDA(cudaFuncSetAttribute(linear_listnet_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    const int vec_ok = F % 4 == 0 && aligned16(features);
    const long long cap = 2LL * di.sms;
    const int grid = static_cast<int>(B < cap ? B : cap);
    linear_listnet_tiled_kernel<<<grid, kFusedThreads, smem, st>>>(features, weight, bias, rel, rel_bytes, n, n_bytes,
                                                                   B, L, F, vec_ok, scores_out, loss_out, dscores_out,
                                                                   qgrad_out, loss_sum);
    LTR_CUDA(cudaGetLastError());
    return LTR_OK;
  }
  int nbuf = 2;
  if (fused_smem_bytes(L, F, rel_bytes, 2) > max_smem) nbuf = 1;
  const size_t smem = fused_smem_bytes(L, F, rel_bytes, nbuf);
  const int grid = di.sms < B ? di.sms : B;
  if (nbuf == 2) {
    LTR_CUDA(cudaFuncSetAttribute(linear_listnet_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    linear_listnet_kernel<2><<<grid, kFusedThreads, smem, st>>>(features, weight, bias, rel, rel_bytes, n, n_bytes, B,
                                                                L, F, scores_out, loss_out, dscores_out, qgrad_out,
                                                                loss_sum);
  } else {
    LTR_CUDA(cudaFuncSetAttribute(linear_listnet_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    linear_listnet_kernel<1><<<grid, kFusedThreads, smem, st>>>(features, weight, bias, rel, rel_bytes, n, n_bytes, B,
                                                                L, F, scores_out, loss_out, dscores_out, qgrad_out,
                                                                loss_sum);
  }
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

int ltr_linear_listnet_backward(const float* qgrad, const float* g, int g_stride, int B, int F,
                                float* dweight_out, float* dbias_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (B < 0 || F < 1 || (g_stride != 0 && g_stride != 1)) return LTR_EINVAL;
  if (!dweight_out || !workspace || (B > 0 && (!qgrad || !g))) return LTR_EINVAL;
  if (workspace_bytes < ltr_linear_listnet_workspace_bytes(F)) return LTR_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partials = static_cast<float*>(workspace);
  const int rows = B < kFusedBwdCtas ? (B > 0 ? B : 1) : kFusedBwdCtas;
  weighted_colsum_kernel<<<rows, 256, 0, st>>>(qgrad, g, g_stride, B, F + 1, partials);
  LTR_CUDA(cudaGetLastError());
  reduce_partials_kernel<<<1, 256, 0, st>>>(partials, rows, F + 1, dweight_out, dbias_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

// ---- all-reduce over NVLink peer memory (ltr_p2p.cuh) ---------------------------------------------------
// The loss path's only exchange is the [sum of losses (, count)] behind a global mean: a few floats per step,
// pure latency -- one single-CTA kernel instead of a library collective (~20 us inside a step of ~450 us at
// 8 GPUs).  The vector form carries the parameter gradient of a data-parallel ranker (ltr_mlp.cu).
__global__ void __launch_bounds__(kP2PMaxRanks * kP2PMaxValues)
p2p_allreduce_kernel(float* __restrict__ values, int k, int rank, int world, P2PMailbox* __restrict__ mine) {
  __shared__ unsigned int seq_s;
  __shared__ float recv[kP2PMaxRanks][kP2PMaxValues];
  const int t = threadIdx.x;
  if (t == 0) {
    seq_s = mine->seq + 1u;
    mine->seq = seq_s;
  }
  __syncthreads();
  const unsigned int seq = seq_s;
  const int par = static_cast<int>(seq & 1u);
  if (t < world * k) {
    const int peer = t / k, j = t - peer * k;
    const unsigned long long packed =
        (static_cast<unsigned long long>(seq) << 32) | static_cast<unsigned long long>(__float_as_uint(values[j]));
    volatile unsigned long long* dst = &mine->peers[peer]->slot[par][rank][j];
    *dst = packed;                                            // remote store over NVLink (local for peer == rank)
    const volatile unsigned long long* src = &mine->slot[par][peer][j];
    unsigned long long w = *src;
    const long long t0 = clock64();
    while (static_cast<unsigned int>(w >> 32) != seq) {
      if (clock64() - t0 > kP2PTimeoutCycles) { mine->error = 1u; break; }
      w = *src;
    }
    recv[peer][j] = __uint_as_float(static_cast<unsigned int>(w & 0xffffffffull));
  }
  __syncthreads();
  if (t < k) {
    float s = 0.0f;
    for (int r = 0; r < world; ++r) s += recv[r][t];          // rank order: the same bits on every rank
    values[t] = s;
  }
}

// values[0..k) (any k, pieces of kP2PVecCapacity per launch): replaced by the sum over the ranks
__global__ void __launch_bounds__(256)
p2p_allreduce_vec_kernel(float* __restrict__ values, int k, int rank, int world, P2PMailbox* __restrict__ mine) {
  const unsigned int seq = p2p_vec_begin(mine);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x)
    values[j] = p2p_vec_element(mine, rank, world, seq, j, values[j]);
  p2p_vec_finish(mine, seq);
}

int ltr_p2p_create(int rank, int world, ltr_p2p** out, unsigned char* handle_out) {
  if (!out || !handle_out || world < 1 || world > kP2PMaxRanks || rank < 0 || rank >= world) return LTR_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  ltr_p2p* p = new (std::nothrow) ltr_p2p();
  if (!p) return LTR_EINVAL;
  p->rank = rank;
  p->world = world;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p->mine), p2p_mailbox_bytes(world));
  if (e == cudaSuccess) e = cudaMemset(p->mine, 0, p2p_mailbox_bytes(world));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->mine);
  if (e != cudaSuccess) {
    if (p->mine) cudaFree(p->mine);
    delete p;
    return cuda_fail(e);
  }
  memcpy(handle_out, &h, 64);
  *out = p;
  return LTR_OK;
}

// handles: world x 64 bytes in rank order (every rank's ltr_p2p_create output, exchanged by the caller)
int ltr_p2p_connect(ltr_p2p* p, const unsigned char* handles) {
  if (!p || !handles) return LTR_EINVAL;
  P2PMailbox* table[kP2PMaxRanks] = {};
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) {
      table[r] = p->mine;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    void* ptr = nullptr;
    LTR_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->opened[r] = ptr;
    table[r] = static_cast<P2PMailbox*>(ptr);
  }
  LTR_CUDA(cudaMemcpy(p->mine->peers, table, sizeof(table), cudaMemcpyHostToDevice));
  const unsigned int w = static_cast<unsigned int>(p->world);
  LTR_CUDA(cudaMemcpy(&p->mine->world, &w, sizeof(w), cudaMemcpyHostToDevice));
  return LTR_OK;
}

// values [k] (device, k <= 4): replaced by the sum over all ranks.  Every rank must call it the same number
// of times.  Enqueued on `stream`; CUDA-graph capturable.
int ltr_p2p_allreduce_sum(ltr_p2p* p, float* values, int k, void* stream) {
  if (!p || !values || k < 1 || k > kP2PMaxValues) return LTR_EINVAL;
  p2p_allreduce_kernel<<<1, kP2PMaxRanks * kP2PMaxValues, 0, static_cast<cudaStream_t>(stream)>>>(
      values, k, p->rank, p->world, p->mine);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

// values [k] (device, any k >= 1): replaced by the sum over all ranks, kP2PVecCapacity floats per launch.
// Every rank must call it with the same k.  Enqueued on `stream`; CUDA-graph capturable.
int ltr_p2p_allreduce_vec(ltr_p2p* p, float* values, long long k, void* stream) {
  if (!p || !values || k < 1) return LTR_EINVAL;
  for (long long off = 0; off < k; off += kP2PVecCapacity) {
    const int piece = static_cast<int>(k - off < kP2PVecCapacity ? k - off : kP2PVecCapacity);
    const int grid = (piece + 255) / 256 < 64 ? (piece + 255) / 256 : 64;
    p2p_allreduce_vec_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(values + off, piece, p->rank,
                                                                                 p->world, p->mine);
    LTR_CUDA(cudaGetLastError());
  }
  return LTR_OK;
}

// 1 if an exchange gave up waiting for a peer (synchronises the device)
int ltr_p2p_error(ltr_p2p* p) {
  if (!p) return LTR_EINVAL;
  unsigned int err = 0;
  LTR_CUDA(cudaMemcpy(&err, &p->mine->error, sizeof(err), cudaMemcpyDeviceToHost));
  return static_cast<int>(err);
}

void ltr_p2p_destroy(ltr_p2p* p) {
  if (!p) return;
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
  if (p->mine) cudaFree(p->mine);
  delete p;
}

size_t ltr_host_workspace_bytes(int B, int L) {
  if (B < 0 || L < 1) return 0;
  const size_t bl = static_cast<size_t>(B) * L;
  // scores f32 | relevance (sized for int64) | n (sized for int64) | loss f32 | dscores f32 | schedule
  return align256(bl * 4) + align256(bl * 8) + align256(static_cast<size_t>(B) * 8) +
         align256(static_cast<size_t>(B) * 4) + align256(bl * 4) + align256(schedule_bytes(B));
}

size_t ltr_host_workspace_dscores_offset(int B, int L) {
  if (B < 0 || L < 1) return 0;
  const size_t bl = static_cast<size_t>(B) * L;
  return align256(bl * 4) + align256(bl * 8) + align256(static_cast<size_t>(B) * 8) +
         align256(static_cast<size_t>(B) * 4);
}

}  // extern "C"

namespace ltr {

// ---- host-buffer pipeline ------------------------------------------------------------------------------
// Large batches are cut into chunks of queries: the H2D copy of chunk c + 1 (copy-in stream), the kernel
// of chunk c (caller's stream) and the D2H copy of chunk c - 1 (copy-out stream) overlap, so the call is
// bound by the slower PCIe direction instead of the sum of the three.  The two internal streams and the
// events are created once per device on first use; the caller's stream waits for everything at the end.
constexpr int kPipeMaxChunks = 16;
constexpr size_t kPipeChunkBytes = 16u << 20;

struct HostPipe {
  std::atomic_flag busy = ATOMIC_FLAG_INIT;
  bool ready = false;
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t start = nullptr, done = nullptr, ev_in[kPipeMaxChunks], ev_k[kPipeMaxChunks];
};

struct PipeLock {
  HostPipe* p;
  explicit PipeLock(HostPipe* pipe) : p(pipe) { while (p->busy.test_and_set(std::memory_order_acquire)) {} }
  ~PipeLock() { p->busy.clear(std::memory_order_release); }
};

inline int host_pipe(HostPipe** out) {
  static HostPipe pipes[64];
  int dev = 0;
  LTR_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return LTR_EUNSUPPORTED;
  HostPipe* p = &pipes[dev];
  PipeLock lock(p);
  if (!p->ready) {
    LTR_CUDA(cudaStreamCreateWithFlags(&p->in, cudaStreamNonBlocking));
    LTR_CUDA(cudaStreamCreateWithFlags(&p->out, cudaStreamNonBlocking));
    LTR_CUDA(cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming));
    LTR_CUDA(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
    for (int i = 0; i < kPipeMaxChunks; ++i) {
      LTR_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
      LTR_CUDA(cudaEventCreateWithFlags(&p->ev_k[i], cudaEventDisableTiming));
    }
    p->ready = true;
  }
  *out = p;
  return LTR_OK;
}

// queries per chunk (a multiple of 64), 0 = do not pipeline
inline int pipe_chunk_queries(int B, size_t bytes_per_query) {
  const size_t total = static_cast<size_t>(B) * bytes_per_query;
  if (total < 2 * kPipeChunkBytes) return 0;
  size_t chunks = (total + kPipeChunkBytes - 1) / kPipeChunkBytes;
  if (chunks > static_cast<size_t>(kPipeMaxChunks)) chunks = kPipeMaxChunks;
  size_t bc = (static_cast<size_t>(B) + chunks - 1) / chunks;
  bc = (bc + 63) / 64 * 64;
  return static_cast<int>(bc);
}

inline int launch_family(int family, int mode, const float* d_scores, const void* d_rel, int rel_bytes,
                         const void* d_n, int n_bytes, int B, int L, float sigma, float* d_loss, float* d_grad,
                         void* d_sched, size_t sched_bytes, void* stream) {
  switch (family) {
    case LTR_FAMILY_ADDITIVE:
      return ltr_pairwise_additive_ws(mode, d_scores, d_rel, rel_bytes, d_n, n_bytes, B, L, sigma, d_loss, d_grad,
                                      nullptr, d_sched, sched_bytes, stream);
    case LTR_FAMILY_LAMBDA:
      return ltr_lambda_ws(mode, d_scores, d_rel, rel_bytes, d_n, n_bytes, B, L, sigma, d_loss, d_grad, nullptr,
                           nullptr, d_sched, sched_bytes, stream);
    case LTR_FAMILY_LISTNET:
      return ltr_listnet(d_scores, d_rel, rel_bytes, d_n, n_bytes, B, L, d_loss, d_grad, nullptr, stream);
    default:
      return LTR_EINVAL;
  }
}

}  // namespace ltr

extern "C" {

int ltr_loss_host_ex(int family, int mode, const float* h_scores, const void* h_rel, int rel_bytes,
                     const void* h_n, int n_bytes, int B, int L, float sigma, float* h_loss_out,
                     float* h_dscores_out, int keep_dscores, void* workspace, size_t workspace_bytes,
                     void* stream) {
  if (B < 0 || L < 1) return LTR_EINVAL;
  if (L > LTR_MAX_LIST_SIZE) return LTR_EUNSUPPORTED;
  if (!rel_width_ok(rel_bytes) || (n_bytes != 4 && n_bytes != 8)) return LTR_EINVAL;
  if (family < LTR_FAMILY_ADDITIVE || family > LTR_FAMILY_LISTNET) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!h_scores || !h_rel || !h_n || !h_loss_out || !workspace) return LTR_EINVAL;
  if (workspace_bytes < ltr_host_workspace_bytes(B, L)) return LTR_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bl = static_cast<size_t>(B) * L;
  unsigned char* p = static_cast<unsigned char*>(workspace);
  float* d_scores = reinterpret_cast<float*>(p);            p += align256(bl * 4);
  unsigned char* d_rel = p;                                 p += align256(bl * 8);
  unsigned char* d_n = p;                                   p += align256(static_cast<size_t>(B) * 8);
  float* d_loss = reinterpret_cast<float*>(p);              p += align256(static_cast<size_t>(B) * 4);
  float* d_grad = reinterpret_cast<float*>(p);              p += align256(bl * 4);
  void* d_sched = p;
  const bool want_grad = h_dscores_out != nullptr || keep_dscores != 0;
  const unsigned char* hr = static_cast<const unsigned char*>(h_rel);
  const unsigned char* hn = static_cast<const unsigned char*>(h_n);
  const size_t row_in = static_cast<size_t>(L) * (4 + rel_bytes) + n_bytes;

  // chunk rows must keep every chunk's device pointers 16-byte aligned (vector / TMA paths): 64 queries do
  const int bc = pipe_chunk_queries(B, row_in + (h_dscores_out ? 4u * L : 0u));
  if (bc <= 0 || bc >= B) {
    LTR_CUDA(cudaMemcpyAsync(d_scores, h_scores, bl * 4, cudaMemcpyHostToDevice, st));
    LTR_CUDA(cudaMemcpyAsync(d_rel, h_rel, bl * rel_bytes, cudaMemcpyHostToDevice, st));
    LTR_CUDA(cudaMemcpyAsync(d_n, h_n, static_cast<size_t>(B) * n_bytes, cudaMemcpyHostToDevice, st));
    int rc = launch_family(family, mode, d_scores, d_rel, rel_bytes, d_n, n_bytes, B, L, sigma, d_loss,
                           want_grad ? d_grad : nullptr, d_sched, schedule_bytes(B), stream);
    if (rc != LTR_OK) return rc;
    LTR_CUDA(cudaMemcpyAsync(h_loss_out, d_loss, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, st));
    if (h_dscores_out) LTR_CUDA(cudaMemcpyAsync(h_dscores_out, d_grad, bl * 4, cudaMemcpyDeviceToHost, st));
    return LTR_OK;
  }

  HostPipe* pipe = nullptr;
  int rc = host_pipe(&pipe);
  if (rc != LTR_OK) return rc;
  PipeLock lock(pipe);
  LTR_CUDA(cudaEventRecord(pipe->start, st));
  LTR_CUDA(cudaStreamWaitEvent(pipe->in, pipe->start, 0));
  LTR_CUDA(cudaStreamWaitEvent(pipe->out, pipe->start, 0));
  int c = 0;
  for (int b0 = 0; b0 < B; b0 += bc, ++c) {
    const int nb = B - b0 < bc ? B - b0 : bc;
    const size_t e0 = static_cast<size_t>(b0) * L, en = static_cast<size_t>(nb) * L;
    LTR_CUDA(cudaMemcpyAsync(d_scores + e0, h_scores + e0, en * 4, cudaMemcpyHostToDevice, pipe->in));
    LTR_CUDA(cudaMemcpyAsync(d_rel + e0 * rel_bytes, hr + e0 * rel_bytes, en * rel_bytes, cudaMemcpyHostToDevice,
                             pipe->in));
    LTR_CUDA(cudaMemcpyAsync(d_n + static_cast<size_t>(b0) * n_bytes, hn + static_cast<size_t>(b0) * n_bytes,
                             static_cast<size_t>(nb) * n_bytes, cudaMemcpyHostToDevice, pipe->in));
    LTR_CUDA(cudaEventRecord(pipe->ev_in[c], pipe->in));
    LTR_CUDA(cudaStreamWaitEvent(st, pipe->ev_in[c], 0));
    rc = launch_family(family, mode, d_scores + e0, d_rel + e0 * rel_bytes, rel_bytes,
                       d_n + static_cast<size_t>(b0) * n_bytes, n_bytes, nb, L, sigma, d_loss + b0,
                       want_grad ? d_grad + e0 : nullptr, d_sched, schedule_bytes(nb), stream);
    if (rc != LTR_OK) return rc;
    LTR_CUDA(cudaEventRecord(pipe->ev_k[c], st));
    LTR_CUDA(cudaStreamWaitEvent(pipe->out, pipe->ev_k[c], 0));
    LTR_CUDA(cudaMemcpyAsync(h_loss_out + b0, d_loss + b0, static_cast<size_t>(nb) * 4, cudaMemcpyDeviceToHost,
                             pipe->out));
    if (h_dscores_out)
      LTR_CUDA(cudaMemcpyAsync(h_dscores_out + e0, d_grad + e0, en * 4, cudaMemcpyDeviceToHost, pipe->out));
  }
  LTR_CUDA(cudaEventRecord(pipe->done, pipe->out));
  LTR_CUDA(cudaStreamWaitEvent(st, pipe->done, 0));
  return LTR_OK;
}

int ltr_loss_host(int family, int mode, const float* h_scores, const int64_t* h_rel,
                  const int64_t* h_n, int B, int L, float sigma, float* h_loss_out,
                  float* h_dscores_out, void* workspace, size_t workspace_bytes, void* stream) {
  return ltr_loss_host_ex(family, mode, h_scores, h_rel, 8, h_n, 8, B, L, sigma, h_loss_out, h_dscores_out,
                          h_dscores_out != nullptr, workspace, workspace_bytes, stream);
}

int ltr_scale_rows_host(float g, const float* d_dscores, float* h_out, int B, int L, float* d_scratch,
                        void* stream) {
  if (B < 0 || L < 1) return LTR_EINVAL;
  if (B == 0) return LTR_OK;
  if (!d_dscores || !h_out) return LTR_EINVAL;
  const bool scale = g != 1.0f;
  if (scale && !d_scratch) return LTR_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bl = static_cast<size_t>(B) * L;
  const float* src = scale ? d_scratch : d_dscores;
  const int bc = pipe_chunk_queries(B, 4u * static_cast<size_t>(L));
  if (bc <= 0 || bc >= B) {
    if (scale) {
      int rc = launch_scale_rows(nullptr, 0, g, d_dscores, d_scratch, B, L, st);
      if (rc != LTR_OK) return rc;
    }
    LTR_CUDA(cudaMemcpyAsync(h_out, src, bl * 4, cudaMemcpyDeviceToHost, st));
    return LTR_OK;
  }
  HostPipe* pipe = nullptr;
  int rc = host_pipe(&pipe);
  if (rc != LTR_OK) return rc;
  PipeLock lock(pipe);
  LTR_CUDA(cudaEventRecord(pipe->start, st));
  LTR_CUDA(cudaStreamWaitEvent(pipe->out, pipe->start, 0));
  int c = 0;
  for (int b0 = 0; b0 < B; b0 += bc, ++c) {
    const int nb = B - b0 < bc ? B - b0 : bc;
    const size_t e0 = static_cast<size_t>(b0) * L, en = static_cast<size_t>(nb) * L;
    if (scale) {
      rc = launch_scale_rows(nullptr, 0, g, d_dscores + e0, d_scratch + e0, nb, L, st);
      if (rc != LTR_OK) return rc;
      LTR_CUDA(cudaEventRecord(pipe->ev_k[c], st));
      LTR_CUDA(cudaStreamWaitEvent(pipe->out, pipe->ev_k[c], 0));
    }
    LTR_CUDA(cudaMemcpyAsync(h_out + e0, src + e0, en * 4, cudaMemcpyDeviceToHost, pipe->out));
  }
  LTR_CUDA(cudaEventRecord(pipe->done, pipe->out));
  LTR_CUDA(cudaStreamWaitEvent(st, pipe->done, 0));
  return LTR_OK;
}

}  // extern "C"
