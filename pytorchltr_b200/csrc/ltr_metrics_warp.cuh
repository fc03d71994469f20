// ltr_metrics_warp.cuh -- dcg / ndcg / arp (evaluation/dcg.py, evaluation/arp.py) and the
// standalone rank_by_score (utils/tensor_operations.py:48-64) with one WARP per query for list
// sizes up to 1024 (metrics; rank_by_score up to 256): the row is read once with coalesced loads, ranked by an in-register bitonic
// network, and the metric comes out of a shuffle scan / reduction.  HBM traffic is the
// algorithmic 12 L + 12 bytes per query (8 L + 8 for rank_by_score's output).
#pragma once

#include "ltr_pair_warp.cuh"

namespace ltr {

constexpr int kMetricWarps = 4;

// Position p = lane * E + r of a row is read / written by lane `lane`: in a plain array the lanes of a warp
// are E words apart and collide on gcd(E, 32) banks.  One pad word per 32 (pidx) makes both the blocked
// access (lane * E + r) and the coalesced one (q * 32 + lane) conflict free.
__device__ __forceinline__ int pidx(int i) { return i + (i >> 5); }

template <int E>
struct MetricScratch {
  float raw_s[33 * E];    // padded (pidx)
  int raw_y[33 * E];      // padded (pidx)
};

// Exact (32-bit key, document) order of two neighbouring ranks: true when (ka, da) must come after (kb, db).
__device__ __forceinline__ bool key_doc_after(uint32_t ka, int da, uint32_t kb, int db) {
  return ka > kb || (ka == kb && da > db);
}

// Ranks the valid documents of one row by descending score (ties: lowest index first).
// On return position p = lane * E + r < nb holds document doc[r]; positions >= nb are the padded
// documents in index order, i.e. document p itself (not produced here).
// The network sorts one 32-bit word per document: the top bits of the score key with the document index
// in the low log2(32 E) bits.  That order is exact unless two neighbouring ranks share their key bits;
// only then the exact keys are fetched again (raw_s) and the neighbours are put right by odd-even
// transposition rounds (runs of equal key bits are short: scores closer than 2^-15 relative), with the
// exact 64-bit network as the last resort for pathological inputs.
template <int E, bool PADDED = false>
__device__ __forceinline__ void warp_rank_by_score(const float (&sv)[E], int nb, int lane,
                                                   const float* __restrict__ raw_s, int (&doc)[E]) {
  constexpr uint32_t kIdxMask = 32u * E - 1u;
  uint32_t pk[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int j = lane * E + r;
    const uint32_t ekey = j < nb ? desc_key_f32(sv[r]) : kPadKey;
    pk[r] = (ekey & ~kIdxMask) | static_cast<uint32_t>(j);
  }
  warp_bitonic_sort32<E>(pk, lane);
  const uint32_t pk_next = __shfl_down_sync(0xffffffffu, pk[0], 1);
  bool ambiguous = false;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    doc[r] = static_cast<int>(pk[r] & kIdxMask);
    const uint32_t nxt = r + 1 < E ? pk[(r + 1) % E] : pk_next;
    ambiguous = ambiguous || (((pk[r] ^ nxt) <= kIdxMask) && (lane * E + r + 1 < nb));
  }
  if (!__any_sync(0xffffffffu, ambiguous)) return;

  uint32_t ek[E];
#pragma unroll
  for (int r = 0; r < E; ++r)
    ek[r] = lane * E + r < nb ? desc_key_f32(raw_s[PADDED ? pidx(doc[r]) : doc[r]]) : kPadKey;
  for (int round = 0; round < 8; ++round) {
    bool swapped = false;
    // even phase: (0,1), (2,3), ... inside the lane
#pragma unroll
    for (int r = 0; r + 1 < E; r += 2) {
      if (key_doc_after(ek[r], doc[r], ek[r + 1], doc[r + 1])) {
        const uint32_t tk = ek[r]; ek[r] = ek[r + 1]; ek[r + 1] = tk;
        const int td = doc[r]; doc[r] = doc[r + 1]; doc[r + 1] = td;
        swapped = true;
      }
    }
    // odd phase: (1,2), (3,4), ... inside the lane and (E-1 of this lane, 0 of the next)
#pragma unroll
    for (int r = 1; r + 1 < E; r += 2) {
      if (key_doc_after(ek[r], doc[r], ek[r + 1], doc[r + 1])) {
        const uint32_t tk = ek[r]; ek[r] = ek[r + 1]; ek[r + 1] = tk;
        const int td = doc[r]; doc[r] = doc[r + 1]; doc[r + 1] = td;
        swapped = true;
      }
    }
    if constexpr (E == 1) {
      // one element per lane: the pairs (l, l + 1) with even l, then those with odd l
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const uint32_t nk = __shfl_down_sync(0xffffffffu, ek[0], 1);
        const int nd = __shfl_down_sync(0xffffffffu, doc[0], 1);
        const uint32_t pk_ = __shfl_up_sync(0xffffffffu, ek[0], 1);
        const int pd = __shfl_up_sync(0xffffffffu, doc[0], 1);
        const bool take_up = lane < 31 && (lane & 1) == par && key_doc_after(ek[0], doc[0], nk, nd);
        const bool take_dn = lane > 0 && ((lane - 1) & 1) == par && key_doc_after(pk_, pd, ek[0], doc[0]);
        if (take_up) { ek[0] = nk; doc[0] = nd; }
        else if (take_dn) { ek[0] = pk_; doc[0] = pd; }
        swapped = swapped || take_up || take_dn;
      }
    } else {
      const uint32_t nk = __shfl_down_sync(0xffffffffu, ek[0], 1);
      const int nd = __shfl_down_sync(0xffffffffu, doc[0], 1);
      const uint32_t pk_ = __shfl_up_sync(0xffffffffu, ek[E - 1], 1);
      const int pd = __shfl_up_sync(0xffffffffu, doc[E - 1], 1);
      const bool up_swap = lane < 31 && key_doc_after(ek[E - 1], doc[E - 1], nk, nd);     // my last <-> next first
      const bool dn_swap = lane > 0 && key_doc_after(pk_, pd, ek[0], doc[0]);             // previous last <-> my first
      if (up_swap) { ek[E - 1] = nk; doc[E - 1] = nd; }
      if (dn_swap) { ek[0] = pk_; doc[0] = pd; }
      swapped = swapped || up_swap || dn_swap;
    }
    if (!__any_sync(0xffffffffu, swapped)) return;
  }
  // still moving after 8 rounds: exact 64-bit network over the original documents
  uint64_t xk[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int j = lane * E + r;
    xk[r] = pack_key(j < nb ? desc_key_f32(sv[r]) : kPadKey, j);
  }
  warp_bitonic_sort64<E>(xk, lane);
#pragma unroll
  for (int r = 0; r < E; ++r) doc[r] = static_cast<int>(xk[r] & 0xffffffffu);
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

// Rows arrive by TMA bulk copies (cp.async.bulk + mbarrier) into one of TWO staging buffers per warp:
// the row of the warp's next query is in flight while the current one is ranked (the kernel is
// otherwise bound by the latency of its own row loads).  Rows that cannot be staged (not 16-byte
// aligned) are read with coalesced loads into the same buffers.
// WPB warps per CTA: 4 up to 256 documents (E <= 8), fewer for the longer lists, whose staging buffers and
// registers (E keys, documents and terms per lane) are larger.
#ifndef LTR_RM16_BLOCKS
#define LTR_RM16_BLOCKS 1
#endif
#ifndef LTR_RM32_BLOCKS
#define LTR_RM32_BLOCKS 1
#endif
template <int E, int WPB>
__global__ void __launch_bounds__(WPB * 32, E <= 8 ? 8 : (E == 16 ? LTR_RM16_BLOCKS : LTR_RM32_BLOCKS))
rank_metrics_warp_kernel(int metric, const float* __restrict__ scores, const void* __restrict__ rel,
                         int rel_bytes, const void* __restrict__ n, int n_bytes, int B, int L, int k,
                         int exp_gain, int tma, float* __restrict__ out, int out_ld,
                         const PairTables* __restrict__ tabs) {
  __shared__ __align__(16) unsigned char s_stage[WPB][2][32 * E * 12];
  __shared__ int s_y[WPB][33 * E];                // grades, padded (pidx)
  __shared__ float s_pad[WPB][E >= 16 ? 33 * E : 1];   // E >= 16: the scores once more, padded (pidx)
  __shared__ uint64_t s_bar[WPB][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* raw_y = s_y[warp];
  // padding pays from 16 documents per lane on (16- / 32-way conflicts); below, the extra copy costs more
  constexpr bool kPad = E >= 16;
  auto yi = [](int i) { return kPad ? pidx(i) : i; };
  const float* __restrict__ inv_disc = tabs->inv_disc;
  const uint32_t row_s_bytes = 4u * L, row_y_bytes = static_cast<uint32_t>(rel_bytes) * L;
  const int stride = gridDim.x * WPB;
  int b = blockIdx.x * WPB + warp;
  auto issue_row = [&](int q, int buf) {
    uint64_t* bar = &s_bar[warp][buf];
    unsigned char* stage = s_stage[warp][buf];
    mbar_arrive_expect_tx(bar, row_s_bytes + row_y_bytes);
    tma_load_1d(stage, scores + static_cast<size_t>(q) * L, row_s_bytes, bar);
    tma_load_1d(stage + row_s_bytes, static_cast<const unsigned char*>(rel) + static_cast<size_t>(q) * row_y_bytes,
                row_y_bytes, bar);
  };
  if (tma) {
    if (lane == 0) {
      mbar_init(&s_bar[warp][0], 1);
      mbar_init(&s_bar[warp][1], 1);
      fence_mbar_init();
      if (b < B) issue_row(b, 0);
    }
    __syncwarp();
  }

  for (int iter = 0; b < B; b += stride, ++iter) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    const int buf = iter & 1;
    unsigned char* stage = s_stage[warp][buf];
    float* raw_s = reinterpret_cast<float*>(stage);
    if (tma) {
      if (lane == 0 && b + stride < B) {
        fence_proxy_async();   // the other buffer was last read through the generic proxy
        issue_row(b + stride, buf ^ 1);
      }
      mbar_wait(&s_bar[warp][buf], (iter >> 1) & 1);
      const unsigned char* sy = stage + row_s_bytes;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        if (j < L) raw_y[yi(j)] = load_int_clamped(sy, rel_bytes, j);
      }
    } else {
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        if (j < L) {
          raw_s[j] = scores[base + j];
          raw_y[yi(j)] = load_int_clamped(rel, rel_bytes, base + j);
        }
      }
    }
    float* pad = s_pad[warp];
    __syncwarp();
    if constexpr (kPad) {
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        pad[j + q] = j < L ? raw_s[j] : 0.0f;          // pidx(j) = j + q
      }
      __syncwarp();
    }
    float sv[E];
    int doc[E];
#pragma unroll
    for (int r = 0; r < E; ++r)
      sv[r] = lane * E + r < nb ? (kPad ? pad[pidx(lane * E + r)] : raw_s[lane * E + r]) : 0.0f;
    warp_rank_by_score<E>(sv, nb, lane, raw_s, doc);

    // relevance per rank; ranks >= nb are the padded documents in index order, whose relevance
    // the reference does not mask (dcg.py:85)
    int iy[E];
    float ry[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int p = lane * E + r;
      iy[r] = p < L ? raw_y[yi(p < nb ? doc[r] : p)] : 0;
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
        yk[r] = j < nb ? desc_key_i32(raw_y[yi(j)]) : kPadKey;
      }
      warp_bitonic_sort32<E>(yk, lane);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int p = lane * E + r;
        int gy = 0;
        if (p < nb) gy = static_cast<int>(~yk[r] ^ 0x80000000u);
        else if (p < L) gy = raw_y[yi(p)];
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
      // stage through shared memory for coalesced stores: the padded array (its scores are dead), or, for
      // short lists, the grades' array (dead too)
      float* o = kPad ? pad : reinterpret_cast<float*>(raw_y);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < E; ++r) o[yi(lane * E + r)] = term[r];
      __syncwarp();
      float* __restrict__ go = out + static_cast<size_t>(b) * out_ld;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        if (j < L) go[j] = o[yi(j)];
      }
    }
    __syncwarp();
  }
}

template <int E, int WPB>
__global__ void __launch_bounds__(WPB * 32)
rank_by_score_warp_kernel(const float* __restrict__ scores, const void* __restrict__ n, int n_bytes,
                          int B, int L, int64_t* __restrict__ ranking_out) {
  __shared__ MetricScratch<E> scratch[WPB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  MetricScratch<E>& ws = scratch[warp];
  for (int b = blockIdx.x * WPB + warp; b < B; b += gridDim.x * WPB) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int j = q * 32 + lane;
      if (j < L) ws.raw_s[j + q] = scores[base + j];   // pidx(j) = j + q
    }
    __syncwarp();
    float sv[E];
    int doc[E];
#pragma unroll
    for (int r = 0; r < E; ++r) sv[r] = lane * E + r < nb ? ws.raw_s[pidx(lane * E + r)] : 0.0f;
    warp_rank_by_score<E, true>(sv, nb, lane, ws.raw_s, doc);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int p = lane * E + r;
      ws.raw_y[pidx(p)] = p < nb ? doc[r] : p;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int j = q * 32 + lane;
      if (j < L) ranking_out[base + j] = ws.raw_y[j + q];
    }
    __syncwarp();
  }
}


// ---- dcg@k / ndcg@k for small cut-offs (0 < k <= 32): top-k selection instead of a full sort ----------
//
// evaluation/dcg.py:85-98 needs the relevance at the first k ranks only, and the ideal DCG of
// ndcg (:36) only the k largest grades.  One warp per query, documents j = q * 32 + lane in
// registers straight from coalesced loads:
//   top-k by score : every lane finds its best key; the k-th best of those 32 lane minima (one
//                    32-key bitonic sort) is a threshold T that at least k documents reach; the
//                    documents with key <= T -- k plus a few -- are compacted by ballots into
//                    <= 32 candidates and ordered exactly by (key, index) in one more 32-element
//                    network.  More than 32 candidates (massive ties): one warp argmin per rank.
//   ideal DCG      : 16-bit grade counters packed in two 64-bit words and summed by shuffles
//                    (grades 0..7), grade g then owns the ideal ranks [start, start + count), cut at
//                    k, and contributes gain(g) * (S[end] - S[start]) with S the prefix sums of
//                    1 / log2(2 + r); other grades: one warp max + count per distinct grade.
// Ranks >= n hold the padded documents in index order with their relevance NOT masked, as in the
// reference (dcg.py:85); they enter dcg and ideal dcg alike.
__device__ __forceinline__ void sort32_u32(uint32_t& k, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool keep_min = ((lane & size) == 0) == ((lane & stride) == 0);
      const uint32_t other = __shfl_xor_sync(0xffffffffu, k, stride);
      k = keep_min ? min(k, other) : max(k, other);
    }
  }
}
__device__ __forceinline__ void sort32_u64(uint64_t& k, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool keep_min = ((lane & size) == 0) == ((lane & stride) == 0);
      const uint64_t other = __shfl_xor_sync(0xffffffffu, k, stride);
      k = (keep_min == (other < k)) ? other : k;
    }
  }
}

constexpr int kTopkWarps = 4;

// Rows are staged by TMA bulk copies (cp.async.bulk + one mbarrier per warp) when they are 16-byte
// aligned: a warp unpacks its row into registers / int32 grades at once and immediately re-arms
// the buffer with the row of its NEXT query, which is in flight while the current one is ranked.
template <int E>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_metrics_warp_kernel(int metric, const float* __restrict__ scores, const void* __restrict__ rel,
                         int rel_bytes, const void* __restrict__ n, int n_bytes, int B, int L, int k,
                         int exp_gain, int tma, float* __restrict__ out, int out_ld,
                         const PairTables* __restrict__ tabs) {
  constexpr bool kCanStage = E <= 16;     // 12 bytes per document and warp of static shared memory
  __shared__ int s_rel[kTopkWarps][32 * E];
  __shared__ uint64_t s_cand[kTopkWarps][32];
  __shared__ __align__(16) unsigned char s_stage[kCanStage ? kTopkWarps : 1][kCanStage ? 32 * E * 12 : 16];
  __shared__ uint64_t s_bar[kTopkWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* raw_y = s_rel[warp];
  uint64_t* cand = s_cand[warp];
  const float* __restrict__ inv_disc = tabs->inv_disc;
  const double* __restrict__ prefix = tabs->inv_disc_prefix;
  const int kk = k < L ? k : L;
  const bool staged = kCanStage && tma != 0;
  unsigned char* stage = s_stage[kCanStage ? warp : 0];
  uint64_t* bar = &s_bar[warp];
  const uint32_t row_s_bytes = 4u * L, row_y_bytes = static_cast<uint32_t>(rel_bytes) * L;
  const int stride = gridDim.x * kTopkWarps;
  int b = blockIdx.x * kTopkWarps + warp;
  auto issue_row = [&](int q) {
    mbar_arrive_expect_tx(bar, row_s_bytes + row_y_bytes);
    tma_load_1d(stage, scores + static_cast<size_t>(q) * L, row_s_bytes, bar);
    tma_load_1d(stage + row_s_bytes, static_cast<const unsigned char*>(rel) + static_cast<size_t>(q) * row_y_bytes,
                row_y_bytes, bar);
  };
  if (staged) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      if (b < B) issue_row(b);
    }
    __syncwarp();
  }

  for (int iter = 0; b < B; b += stride, ++iter) {
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    const int kv = kk < nb ? kk : nb;     // ranks held by valid documents

    // ---- the row: document j = q * 32 + lane -----------------------------------------------------------
    uint32_t key[E];
    // grade counters for the ideal DCG, filled while the row is unpacked: eight 4-bit counters per
    // lane (<= 8 documents at a time), widened to 16-bit fields (word f: grades f and f + 4)
    bool wide = false;
    uint32_t hacc = 0u;
    uint32_t wcnt[4] = {0u, 0u, 0u, 0u};
    auto count_grade = [&](int q, int j, int y) {
      const bool valid = j < nb;
      wide = wide || (valid && static_cast<unsigned int>(y) > 7u);
      hacc += valid ? (1u << (4 * (y & 7))) : 0u;
      if ((q & 7) == 7 || q == E - 1) {
#pragma unroll
        for (int f = 0; f < 4; ++f) wcnt[f] += (hacc >> (4 * f)) & 0x000f000fu;
        hacc = 0u;
      }
    };
    if (staged) {
      mbar_wait(bar, iter & 1);
      const float* ss = reinterpret_cast<const float*>(stage);
      const unsigned char* sy = stage + row_s_bytes;
      if (rel_bytes == 8) {
        const int2* sy2 = reinterpret_cast<const int2*>(sy);
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int j = q * 32 + lane;
          float s = 0.0f;
          int y = 0;
          if (j < L) {
            s = ss[j];
            const int2 w = sy2[j];   // little endian: x = low word
            y = w.y == (w.x >> 31) ? w.x : (w.y < 0 ? -2147483647 : 2147483647);
          }
          key[q] = j < nb ? desc_key_f32(s) : kPadKey;
          raw_y[j] = y;
          count_grade(q, j, y);
        }
      } else {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int j = q * 32 + lane;
          float s = 0.0f;
          int y = 0;
          if (j < L) {
            s = ss[j];
            y = load_int_clamped(sy, rel_bytes, j);
          }
          key[q] = j < nb ? desc_key_f32(s) : kPadKey;
          raw_y[j] = y;
          count_grade(q, j, y);
        }
      }
      __syncwarp();
      if (lane == 0 && b + stride < B) {
        fence_proxy_async();   // the buffer was just read through the generic proxy
        issue_row(b + stride);
      }
    } else {
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = q * 32 + lane;
        float s = 0.0f;
        int y = 0;
        if (j < L) {
          s = scores[base + j];
          y = rel_bytes == 8 ? clamp_i64_to_i32(reinterpret_cast<const long long*>(rel)[base + j])
                             : load_int_clamped(rel, rel_bytes, base + j);
        }
        key[q] = j < nb ? desc_key_f32(s) : kPadKey;
        raw_y[j] = y;   // grades wait in shared memory (keeps the registers for loads in flight)
        count_grade(q, j, y);
      }
      __syncwarp();
    }

    // ---- top-kv by (key, index) ---------------------------------------------------------------------------
    int my_doc = lane;                    // rank lane holds document my_doc (ranks >= nb: the padding itself)
    if (kv > 0) {
      uint32_t lmin = key[0];
#pragma unroll
      for (int q = 1; q < E; ++q) lmin = min(lmin, key[q]);
      uint32_t sorted_min = lmin;
      sort32_u32(sorted_min, lane);
      const uint32_t T = __shfl_sync(0xffffffffu, sorted_min, kv - 1);
      // candidates (key <= T): per-lane count, warp scan, each lane appends its own (any order:
      // they are sorted next)
      int mine = 0;
#pragma unroll
      for (int q = 0; q < E; ++q) mine += key[q] <= T ? 1 : 0;
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total <= 32) {
        int pos = incl - mine;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          if (key[q] <= T) cand[pos++] = pack_key(key[q], q * 32 + lane);
        }
        __syncwarp();
        uint64_t c = lane < total ? cand[lane] : ~0ull;
        sort32_u64(c, lane);
        if (lane < kv) my_doc = static_cast<int>(c & 0xffffffffu);
      } else {
        // massive ties at the threshold: one exact warp argmin per rank
        for (int p = 0; p < kv; ++p) {
          uint64_t best = ~0ull;
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const uint64_t c = pack_key(key[q], q * 32 + lane);
            best = c < best ? c : best;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
          }
          const int j = static_cast<int>(best & 0xffffffffu);
          if (lane == p) my_doc = j;
#pragma unroll
          for (int q = 0; q < E; ++q)
            if (q * 32 + lane == j) key[q] = kPadKey;   // taken (kPadKey never beats a valid key)
        }
      }
    }

    // ---- dcg@kk: gain / log2(rank + 2), dcg.py:91-93 -----------------------------------------------------
    float term = 0.0f, pad_term = 0.0f;
    if (lane < kk) {
      const int y = raw_y[my_doc];
      const float g = exp_gain ? gain_of_grade(y) : static_cast<float>(y);
      term = g * __ldg(inv_disc + lane);
      if (lane >= nb) pad_term = term;
    }
    float v = warp_sum(term);

    if (metric == LTR_METRIC_NDCG) {
      // ---- ideal dcg@kk: the kv largest grades of the valid documents, then the same padding ----------
      float iv = warp_sum(pad_term);
      if (kv > 0) {
        double acc = 0.0;
        if (!__any_sync(0xffffffffu, wide)) {
          // 16-bit fields, at most 32 x 32 documents per field: one warp-wide integer add per word (REDUX)
#pragma unroll
          for (int f = 0; f < 4; ++f) wcnt[f] = __reduce_add_sync(0xffffffffu, wcnt[f]);
          int start = 0;
#pragma unroll
          for (int g = 7; g >= 1; --g) {
            const int cnt = static_cast<int>((wcnt[g & 3] >> (16 * (g >> 2))) & 0xffffu);
            const int end = min(start + cnt, kv);
            if (end > start) {
              const float gain = exp_gain ? gain_of_grade(g) : static_cast<float>(g);
              acc += static_cast<double>(gain) * (prefix[end] - prefix[start]);
            }
            start = end;
          }
        } else {
          // arbitrary integer grades: peel off one distinct grade per round, largest first
          uint32_t yk[E];
#pragma unroll
          for (int q = 0; q < E; ++q) yk[q] = q * 32 + lane < nb ? desc_key_i32(raw_y[q * 32 + lane]) : kPadKey;
          int start = 0;
          while (start < kv) {
            uint32_t best = yk[0];
#pragma unroll
            for (int q = 1; q < E; ++q) best = min(best, yk[q]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < E; ++q) {
              const bool hit = yk[q] == best;
              cnt += __popc(__ballot_sync(0xffffffffu, hit));
              if (hit) yk[q] = kPadKey;
            }
            const int g = static_cast<int>(~best ^ 0x80000000u);
            const int end = min(start + cnt, kv);
            const float gain = exp_gain ? gain_of_grade(g) : static_cast<float>(g);
            acc += static_cast<double>(gain) * (prefix[end] - prefix[start]);
            start = end;
          }
        }
        iv += static_cast<float>(acc);
      }
      if (iv == 0.0f) iv = 1.0f;   // dcg.py:37
      v = v / iv;
    }
    if (lane == 0) out[static_cast<size_t>(b) * out_ld] = v;
    __syncwarp();   // raw_y / cand are reused by the next query
  }
}

}  // namespace ltr
