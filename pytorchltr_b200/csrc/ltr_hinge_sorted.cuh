// ltr_hinge_sorted.cuh -- PairwiseHingeLoss / PairwiseDCGHingeLoss (loss/pairwise_additive.py:93-133)
// in O(n log n + n G) per query instead of O(n^2), for list sizes 129 .. 4096 (256 threads per
// query up to 1024 documents, 1024 threads beyond).
//
// The hinge pair term is piecewise linear: with d = fl(s_i - s_j) the pair (i wins on relevance) is
// active iff !(fl(1 - d) < 0), i.e. iff d <= 1 (1 - d is exact around 1), and then contributes
// 1 - d to the loss, -1 to d/ds_i and +1 to d/ds_j.  fl(s_i - s_j) is monotone in s_j, so in
// ASCENDING score order the active partners of a document form a suffix (as the winner: every lower
// grade document from position lo_i on) and a prefix (as the loser: every higher grade document up to
// position hi_i).  Per query, one CTA:
//   1. sort the valid documents by score (cta_block_sort, ltr_pair_ring.cuh);
//   2. lo_i / hi_i by binary search with the exact float32 predicate fl(s_i - s_p) <= 1;
//   3. grades become dense classes 0 .. G-1 (shared-memory histogram of the grades 0 .. 31); for
//      every class boundary c = 1 .. G-1 one CTA-wide scan gives A_c[p] = #{q < p : class_q < c} and
//      S_c[p] = the sum of their scores (double); documents of class c read their winner-side count
//      and score sum off A_c / S_c at lo_i, documents of class c-1 their loser-side count at hi_i;
//   4. loss_b = sum_i count_i (1 - s_i) + scoresum_i (double), gradient = count differences: the
//      integer-valued gradients are exact, as the reference's are.
// Grades outside 0 .. 31 (no histogram): the query falls back to the plain O(n^2) loop in place.
#pragma once

#include "ltr_pair_ring.cuh"

namespace ltr {

constexpr int kHingePer = 4;          // documents per thread: L <= 4 * THREADS
constexpr int kHingeMaxWarps = 32;

struct HingeSmem {
  uint64_t* keys;     // [P]
  float* raw_s;       // [L]   document order; reused: document-order gradient
  int* raw_y;         // [L]
  float* ss;          // [Lp]  ascending scores
  int* cls;           // [Lp]  dense class of the document at each position
  uint16_t* doc;      // [Lp]
  uint16_t* A;        // [Lp + 1]
  double* S;          // [Lp + 1]
  double* wsum_d;     // [32]
  int* wsum_i;        // [32]
  int* hist;          // [40]
  double* red_d;      // [32]
};

__host__ __device__ inline size_t hinge_smem_bytes(int L, int P) {
  const size_t Lp = (static_cast<size_t>(L) + 127) / 128 * 128;
  return 8u * P + 4u * Lp * 2 + 4u * Lp * 2 + 2u * Lp + 2u * (Lp + 8) + 8u * (Lp + 2) + 8u * kHingeMaxWarps +
         4u * kHingeMaxWarps + 4u * 40 + 8u * kHingeMaxWarps;
}

__device__ __forceinline__ HingeSmem hinge_carve(unsigned char* base, int L, int P) {
  const size_t Lp = (static_cast<size_t>(L) + 127) / 128 * 128;
  HingeSmem m;
  m.keys = reinterpret_cast<uint64_t*>(base);      base += 8u * P;
  m.S = reinterpret_cast<double*>(base);           base += 8u * (Lp + 2);
  m.wsum_d = reinterpret_cast<double*>(base);      base += 8u * kHingeMaxWarps;
  m.red_d = reinterpret_cast<double*>(base);       base += 8u * kHingeMaxWarps;
  m.raw_s = reinterpret_cast<float*>(base);        base += 4u * Lp;
  m.raw_y = reinterpret_cast<int*>(base);          base += 4u * Lp;
  m.ss = reinterpret_cast<float*>(base);           base += 4u * Lp;
  m.cls = reinterpret_cast<int*>(base);            base += 4u * Lp;
  m.wsum_i = reinterpret_cast<int*>(base);         base += 4u * kHingeMaxWarps;
  m.hist = reinterpret_cast<int*>(base);           base += 4u * 40;
  m.doc = reinterpret_cast<uint16_t*>(base);       base += 2u * Lp;
  m.A = reinterpret_cast<uint16_t*>(base);
  return m;
}

// ascending score key: complement of the descending key (ltr_common.cuh)
__device__ __forceinline__ uint32_t asc_key_f32(float x) { return ~desc_key_f32(x); }

template <int kHingeThreads>
__global__ void __launch_bounds__(kHingeThreads)
hinge_sorted_kernel(const float* __restrict__ scores, const void* __restrict__ rel, int rel_bytes,
                    const void* __restrict__ n, int n_bytes, int B, int L, int P, int variant,
                    float* __restrict__ loss_out, float* __restrict__ grad_out, float* __restrict__ loss_sum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const HingeSmem m = hinge_carve(smem_raw, L, P);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kHingeThreads / 32;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();   // previous query fully consumed
    const int nb = load_n(n, n_bytes, b, L);
    const size_t base = static_cast<size_t>(b) * L;
    for (int j = tid; j < P; j += kHingeThreads) {
      uint64_t key = ~0ull;
      if (j < L) {
        const float s = scores[base + j];
        m.raw_s[j] = s;
        m.raw_y[j] = load_int_clamped(rel, rel_bytes, base + j);
        // asc_key never equals 0xffffffff for a valid document?  desc_key == 0 only for the
        // canonical positive NaN; fold that case one step down so that padding stays last
        uint32_t k = asc_key_f32(s);
        k = k == 0xffffffffu ? 0xfffffffeu : k;
        if (j < nb) key = pack_key(k, j);
      }
      m.keys[j] = key;
    }
    if (tid < 36) m.hist[tid] = 0;
    cta_block_sort(m.keys, P, lane, warp, kWarps);

    // ---- sorted scores, grade histogram ----------------------------------------------------------------
    for (int p = tid; p < nb; p += kHingeThreads) {
      const int d = static_cast<int>(m.keys[p] & 0xffffffffu);
      const int y = m.raw_y[d];
      m.doc[p] = static_cast<uint16_t>(d);
      m.ss[p] = m.raw_s[d];
      m.cls[p] = y;   // grade for now
      if (y < 0 || y > 31) m.hist[32] = 1;
      else atomicAdd(&m.hist[y], 1);
    }
    __syncthreads();
    const bool wide = m.hist[32] != 0;
    double lacc = 0.0;
    int gcnt[kHingePer];   // d loss / d s at position p = tid + 256 k (integer valued)
#pragma unroll
    for (int k = 0; k < kHingePer; ++k) gcnt[k] = 0;

    if (!wide) {
      // ---- dense classes: class = number of present lower grades ---------------------------------------
      unsigned int present = 0u;
      {
        const unsigned int bit = (lane < 32 && m.hist[lane] > 0) ? 1u : 0u;
        present = __ballot_sync(0xffffffffu, bit != 0u);
      }
      const int G = __popc(present);
      int lo[kHingePer], hi[kHingePer], cl[kHingePer];
      float sp[kHingePer];
#pragma unroll
      for (int k = 0; k < kHingePer; ++k) {
        const int p = tid + kHingeThreads * k;
        lo[k] = 0; hi[k] = -1; cl[k] = 0; sp[k] = 0.0f;
        if (p < nb) {
          const int y = m.cls[p];
          cl[k] = __popc(present & ((1u << y) - 1u));
          sp[k] = m.ss[p];
          // lo: first position q with fl(s_p - s_q) <= 1 (true from some q on: scores ascend)
          int a = 0, c = nb;
          while (a < c) {
            const int mid = (a + c) >> 1;
            if (sp[k] - m.ss[mid] <= 1.0f) c = mid; else a = mid + 1;
          }
          lo[k] = a;
          // hi: last position q with fl(s_q - s_p) <= 1 (true up to some q)
          a = 0; c = nb;
          while (a < c) {
            const int mid = (a + c) >> 1;
            if (m.ss[mid] - sp[k] <= 1.0f) a = mid + 1; else c = mid;
          }
          hi[k] = a - 1;
        }
      }
      __syncthreads();   // every thread has read its grades: cls[] becomes the class array
#pragma unroll
      for (int k = 0; k < kHingePer; ++k) {
        const int p = tid + kHingeThreads * k;
        if (p < nb) m.cls[p] = cl[k];
      }
      __syncthreads();

      // ---- one scan per class boundary ---------------------------------------------------------------------
      for (int c = 1; c < G; ++c) {
        // thread t owns positions [4 t, 4 t + 4): inclusive scan of (class < c) counts and score sums
        int ci[kHingePer];
        double si[kHingePer];
        int run_i = 0;
        double run_d = 0.0;
#pragma unroll
        for (int k = 0; k < kHingePer; ++k) {
          const int p = tid * kHingePer + k;
          const bool in = p < nb && m.cls[p] < c;
          run_i += in ? 1 : 0;
          run_d += in ? static_cast<double>(m.ss[p]) : 0.0;
          ci[k] = run_i; si[k] = run_d;
        }
        int inc_i = run_i;
        double inc_d = run_d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int ti = __shfl_up_sync(0xffffffffu, inc_i, o);
          const double td = __shfl_up_sync(0xffffffffu, inc_d, o);
          if (lane >= o) { inc_i += ti; inc_d += td; }
        }
        if (lane == 31) { m.wsum_i[warp] = inc_i; m.wsum_d[warp] = inc_d; }
        __syncthreads();
        int off_i = inc_i - run_i;
        double off_d = inc_d - run_d;
        for (int w = 0; w < warp; ++w) { off_i += m.wsum_i[w]; off_d += m.wsum_d[w]; }
        // A[p + 1] = count over positions <= p; A[0] = 0
#pragma unroll
        for (int k = 0; k < kHingePer; ++k) {
          const int p = tid * kHingePer + k;
          if (p < nb) {
            m.A[p + 1] = static_cast<uint16_t>(off_i + ci[k]);
            m.S[p + 1] = off_d + si[k];
          }
        }
        if (tid == 0) { m.A[0] = 0; m.S[0] = 0.0; }
        __syncthreads();
        const int tot_i = m.A[nb];
        const double tot_d = m.S[nb];
#pragma unroll
        for (int k = 0; k < kHingePer; ++k) {
          const int p = tid + kHingeThreads * k;
          if (p < nb) {
            if (cl[k] == c) {
              // winner side: lower classes at positions >= lo
              const int cnt = tot_i - m.A[lo[k]];
              const double ssum = tot_d - m.S[lo[k]];
              lacc += static_cast<double>(cnt) * (1.0 - static_cast<double>(sp[k])) + ssum;
              gcnt[k] -= cnt;
            } else if (cl[k] == c - 1) {
              // loser side: higher classes (class >= c) at positions <= hi
              gcnt[k] += (hi[k] + 1) - m.A[hi[k] + 1];
            }
          }
        }
        __syncthreads();   // A / S are rewritten by the next boundary
      }
      // classes above c need every boundary below them too: a document of class k wins against all
      // classes < k, which boundary k alone covers (A_k counts class < k); it loses against all
      // classes > k, which boundary k + 1 alone covers (positions minus class <= k).
    } else {
      // ---- grades outside 0 .. 31: plain O(n^2) loop (rare) ------------------------------------------------
#pragma unroll
      for (int k = 0; k < kHingePer; ++k) {
        const int p = tid + kHingeThreads * k;
        if (p < nb) {
          const float si = m.ss[p];
          const int yi = m.cls[p];
          int g = 0;
          for (int q = 0; q < nb; ++q) {
            const int yq = m.cls[q];
            const float sq = m.ss[q];
            if (yi > yq) {
              const float l = 1.0f - (si - sq);
              if (!(l < 0.0f)) { lacc += static_cast<double>(l); g -= 1; }
            } else if (yq > yi) {
              const float l = 1.0f - (sq - si);
              if (!(l < 0.0f)) g += 1;
            }
          }
          gcnt[k] = g;
        }
      }
    }

    // ---- loss, modifier, gradient back to document order ------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lacc += __shfl_xor_sync(0xffffffffu, lacc, o);
    __syncthreads();
    if (lane == 0) m.red_d[warp] = lacc;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < kWarps; ++w) tot += m.red_d[w];
    float loss = static_cast<float>(tot);
    float gmul = 1.0f;
    if (variant) {
      // pairwise_additive.py:132-133: -1 / ln(2 + h); d/dh = 1 / ((2 + h) ln^2(2 + h))
      const float lg = logf(2.0f + loss);
      gmul = 1.0f / ((2.0f + loss) * lg * lg);
      loss = -1.0f / lg;
    }
    if (tid == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }
    if (grad_out) {
      float* gdoc = m.raw_s;
      for (int j = tid; j < L; j += kHingeThreads) gdoc[j] = 0.0f;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kHingePer; ++k) {
        const int p = tid + kHingeThreads * k;
        if (p < nb) gdoc[m.doc[p]] = static_cast<float>(gcnt[k]) * gmul;
      }
      __syncthreads();
      float* __restrict__ go = grad_out + base;
      for (int j = tid; j < L; j += kHingeThreads) go[j] = gdoc[j];
    }
  }
}

}  // namespace ltr
