// ltr_metrics_warp.cuh -- dcg / ndcg / arp (evaluation/dcg.py, evaluation/arp.py) and the
// standalone rank_by_score (utils/tensor_operations.py:48-64) with one WARP per query for list
// sizes up to 256: the row is read once with coalesced loads, ranked by an in-register bitonic
// network, and the metric comes out of a shuffle scan / reduction.  HBM traffic is the
// algorithmic 12 L + 12 bytes per query (8 L + 8 for rank_by_score's output).
#pragma once

#include "ltr_pair_warp.cuh"

namespace ltr {

constexpr int kMetricWarps = 4;

template <int E>
struct MetricScratch {
  float raw_s[32 * E];
  int raw_y[32 * E];
};

// Ranks the valid documents of one row by descending score (ties: lowest index first).
// On return position p = lane * E + r < nb holds document doc[r]; positions >= nb are the padded
// documents in index order, i.e. document p itself (not produced here).
template <int E>
__device__ __forceinline__ void warp_rank_by_score(const float (&sv)[E], int nb, int lane,
                                                   const float* __restrict__ raw_s, int (&doc)[E]) {
  constexpr uint32_t kIdxMask = 32u * E - 1u;
  uint32_t ekey[E], pk[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int j = lane * E + r;
    ekey[r] = j < nb ? desc_key_f32(sv[r]) : kPadKey;
    pk[r] = (ekey[r] & ~kIdxMask) | static_cast<uint32_t>(j);
  }
  warp_bitonic_sort32<E>(pk, lane);
  uint64_t xk[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    doc[r] = static_cast<int>(pk[r] & kIdxMask);
    const int p = lane * E + r;
    xk[r] = pack_key(p < nb ? desc_key_f32(raw_s[doc[r]]) : kPadKey, doc[r]);
  }
  const uint64_t next0 = __shfl_down_sync(0xffffffffu, xk[0], 1);
  bool bad = lane < 31 && xk[E - 1] > next0;
#pragma unroll
  for (int r = 0; r + 1 < E; ++r) bad = bad || (xk[r] > xk[r + 1]);
  if (__any_sync(0xffffffffu, bad)) {
    // two scores closer than the packed key can resolve: exact 64-bit network
#pragma unroll
    for (int r = 0; r < E; ++r) xk[r] = pack_key(ekey[r], lane * E + r);
    warp_bitonic_sort64<E>(xk, lane);
#pragma unroll
    for (int r = 0; r < E; ++r) doc[r] = static_cast<int>(xk[r] & 0xffffffffu);
  }
}

// Inclusive scan over positions p = lane * E + r of per-position values.
template <int E>
__device__ __forceinline__ void warp_inclusive_scan(float (&v)[E], int lane) {
#pragma unroll
  for (int r = 1; r < E; ++r) v[r] += v[r - 1];
  float tot = v[E - 1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, tot, o);
    if (lane >= o) tot += t;
  }
  const float excl = tot - v[E - 1];
#pragma unroll
  for (int r = 0; r < E; ++r) v[r] += excl;
}

template <int E>
__global__ void __launch_bounds__(kMetricWarps * 32)
rank_metrics_warp_kernel(int metric, const float* __restrict__ scores, const void* __restrict__ rel,
                         int rel_bytes, const void* __restrict__ n, int n_bytes, int B, int L, int k,
                         int exp_gain, float* __restrict__ out, int out_ld,
                         const PairTables* __restrict__ tabs) {
  __shared__ MetricScratch<E> scratch[kMetricWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  MetricScratch<E>& ws = scratch[warp];
  const float* __restrict__ inv_disc = tabs->inv_disc;
  for (int b = blockIdx.x * kMetricWarps + warp; b < B; b += gridDim.x * kMetricWarps) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    // coalesced loads (document j = k * 32 + lane) staged through shared memory into the blocked
    // layout of the sorting network (document j = lane * E + r)
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int j = q * 32 + lane;
      if (j < L) {
        ws.raw_s[j] = scores[base + j];
        ws.raw_y[j] = rel_bytes == 8
                          ? clamp_i64_to_i32(reinterpret_cast<const long long*>(rel)[base + j])
                          : reinterpret_cast<const int*>(rel)[base + j];
      }
    }
    __syncwarp();
    float sv[E];
    int doc[E];
#pragma unroll
    for (int r = 0; r < E; ++r) sv[r] = lane * E + r < nb ? ws.raw_s[lane * E + r] : 0.0f;
    warp_rank_by_score<E>(sv, nb, lane, ws.raw_s, doc);

    // relevance per rank; ranks >= nb are the padded documents in index order, whose relevance
    // the reference does not mask (dcg.py:85)
    int iy[E];
    float ry[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int p = lane * E + r;
      iy[r] = p < L ? ws.raw_y[p < nb ? doc[r] : p] : 0;
      ry[r] = static_cast<float>(iy[r]);
    }

    if (metric == LTR_METRIC_ARP) {
      // arp.py:32-42
      float srp = 0.0f, nrp = 0.0f;
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int p = lane * E + r;
        if (p < nb) { srp = fmaf(static_cast<float>(p + 1), ry[r], srp); nrp += ry[r]; }
      }
      srp = warp_sum(srp);
      nrp = warp_sum(nrp);
      if (nrp == 0.0f) nrp = 1.0f;
      if (lane == 0) out[static_cast<size_t>(b) * out_ld] = srp / nrp;
      __syncwarp();
      continue;
    }

    // dcg.py:91-93: gain / log2(rank + 2)  (as gain * (1 / log2(rank + 2)): <= 1 ulp apart)
    float term[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int p = lane * E + r;
      const float g = exp_gain ? gain_of_grade(iy[r]) : ry[r];
      term[r] = p < L ? g * __ldg(inv_disc + p) : 0.0f;
    }
    float iterm[E];
    if (metric == LTR_METRIC_NDCG) {
      // ideal ranking (dcg.py:36): valid documents by descending relevance, then the padding
      uint32_t yk[E];
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int j = lane * E + r;
        yk[r] = j < nb ? desc_key_i32(ws.raw_y[j]) : kPadKey;
      }
      warp_bitonic_sort32<E>(yk, lane);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int p = lane * E + r;
        int gy = 0;
        if (p < nb) gy = static_cast<int>(~yk[r] ^ 0x80000000u);
        else if (p < L) gy = ws.raw_y[p];
        const float g = exp_gain ? gain_of_grade(gy) : static_cast<float>(gy);
        iterm[r] = p < L ? g * __ldg(inv_disc + p) : 0.0f;
      }
    }

    const int kk = k > 0 ? (k < L ? k : L) : 0;
    if (kk > 0) {
      float part = 0.0f, ipart = 0.0f;
#pragma unroll
      for (int r = 0; r < E; ++r) {
        if (lane * E + r < kk) {
          part += term[r];
          if (metric == LTR_METRIC_NDCG) ipart += iterm[r];
        }
      }
      float v = warp_sum(part);
      if (metric == LTR_METRIC_NDCG) {
        float iv = warp_sum(ipart);
        if (iv == 0.0f) iv = 1.0f;   // dcg.py:37
        v = v / iv;
      }
      if (lane == 0) out[static_cast<size_t>(b) * out_ld] = v;
    } else {
      warp_inclusive_scan<E>(term, lane);   // dcg.py:94 cumsum
      if (metric == LTR_METRIC_NDCG) {
        warp_inclusive_scan<E>(iterm, lane);
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const float iv = iterm[r] == 0.0f ? 1.0f : iterm[r];
          term[r] = term[r] / iv;
        }
      }
      // stage through shared memory for coalesced stores
      float* o = ws.raw_s;
      __syncwarp();
#pragma unroll
      for (int r = 0; r < E; ++r) o[lane * E + r] = term[r];
      __syncwarp();
      float* __restrict__ go = out + static_cast<size_t>(b) * out_ld;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        if (j < L) go[j] = o[j];
      }
    }
    __syncwarp();
  }
}

template <int E>
__global__ void __launch_bounds__(kMetricWarps * 32)
rank_by_score_warp_kernel(const float* __restrict__ scores, const void* __restrict__ n, int n_bytes,
                          int B, int L, int64_t* __restrict__ ranking_out) {
  __shared__ MetricScratch<E> scratch[kMetricWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  MetricScratch<E>& ws = scratch[warp];
  for (int b = blockIdx.x * kMetricWarps + warp; b < B; b += gridDim.x * kMetricWarps) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int j = q * 32 + lane;
      if (j < L) ws.raw_s[j] = scores[base + j];
    }
    __syncwarp();
    float sv[E];
    int doc[E];
#pragma unroll
    for (int r = 0; r < E; ++r) sv[r] = lane * E + r < nb ? ws.raw_s[lane * E + r] : 0.0f;
    warp_rank_by_score<E>(sv, nb, lane, ws.raw_s, doc);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int p = lane * E + r;
      ws.raw_y[p] = p < nb ? doc[r] : p;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int j = q * 32 + lane;
      if (j < L) ranking_out[base + j] = ws.raw_y[j];
    }
    __syncwarp();
  }
}

}  // namespace ltr
