// ltr_collate.cuh -- ragged -> padded batches on the GPU: the step right before
// the loss path.  Reference: SVMRankDataset.collate_fn()._collate_fn with the default ListSampler
// (datasets/svmrank/svmrank.py:135-205, datasets/list_sampler.py:5-19): a Python loop over the samples
// of a batch that copies each query's (n_i, F) feature block and relevance vector into zero-filled
// (B, list_size, F) / (B, list_size) tensors, truncated to the first list_size documents.
//
// Here the dataset lives in device memory as one ragged (CSR) block -- features [N, F], relevance
// [N], offsets [Q + 1] -- and a batch is a list of query indices.  The documents of a query are
// contiguous, so one CTA per (query, 32 KB slab) streams a contiguous range of floats into the padded
// row block with 128-bit loads / stores and zero-fills the rest: pure HBM traffic,
// 4 F (n_b + L) + 8 (n_b + L) + 24 bytes per query.
#pragma once

#include "ltr_common.cuh"

namespace ltr {

constexpr int kCollateThreads = 256;
constexpr int kCollateSlab = 8192;    // floats of the padded feature block handled by one CTA

__global__ void __launch_bounds__(kCollateThreads)
collate_kernel(const float* __restrict__ features, const int64_t* __restrict__ relevance,
               const int64_t* __restrict__ offsets, const int64_t* __restrict__ qidx, int B, int L, int F,
               int slabs, int vec_ok, float* __restrict__ feat_out, int64_t* __restrict__ rel_out,
               int64_t* __restrict__ n_out, int64_t* __restrict__ cnt_out) {
  const int b = blockIdx.x / slabs, slab = blockIdx.x - b * slabs;
  if (b >= B) return;
  const int64_t q = qidx[b];
  const int64_t off = offsets[q];
  const int64_t cnt = offsets[q + 1] - off;
  const int n = static_cast<int>(cnt < L ? cnt : L);           // ListSampler: the first list_size documents
  const size_t row_floats = static_cast<size_t>(L) * F;
  const size_t valid = static_cast<size_t>(n) * F;
  const size_t lo = static_cast<size_t>(slab) * kCollateSlab;
  const size_t hi = lo + kCollateSlab < row_floats ? lo + kCollateSlab : row_floats;
  const float* __restrict__ src = features + static_cast<size_t>(off) * F;
  float* __restrict__ dst = feat_out + static_cast<size_t>(b) * row_floats;
  if (vec_ok) {
    const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
    float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
    const size_t v4 = valid >> 2;
    for (size_t i = (lo >> 2) + threadIdx.x; i < (hi >> 2); i += kCollateThreads)
      d4[i] = i < v4 ? s4[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  } else {
    for (size_t i = lo + threadIdx.x; i < hi; i += kCollateThreads) dst[i] = i < valid ? src[i] : 0.0f;
  }
  if (slab == 0) {
    for (int l = threadIdx.x; l < L; l += kCollateThreads)
      rel_out[static_cast<size_t>(b) * L + l] = l < n ? relevance[off + l] : 0;
    if (threadIdx.x == 0) {
      n_out[b] = n;                      // min(sample.n, list_size), svmrank.py:194
      if (cnt_out) cnt_out[b] = cnt;
    }
  }
}

// Sampled collation (list samplers, datasets/list_sampler.py:19-61): query b of the batch takes the
// documents sel[b, 0 .. n_b) of its list (indices inside the query, e.g. a random permutation prefix)
// when it holds more than L documents, and all of its documents in order otherwise, exactly like
// collate_fn (svmrank.py:159-190: the sampler is only consulted when xs.shape[0] > list_size).
// One warp per output row: a row is F contiguous floats of one source document.
constexpr int kGatherRowsPerCta = 32;

__global__ void __launch_bounds__(kCollateThreads)
collate_gather_kernel(const float* __restrict__ features, const int64_t* __restrict__ relevance,
                      const int64_t* __restrict__ offsets, const int64_t* __restrict__ qidx,
                      const int64_t* __restrict__ sel, int sel_ld, int B, int L, int F, int row_groups, int vec_ok,
                      float* __restrict__ feat_out, int64_t* __restrict__ rel_out, int64_t* __restrict__ n_out,
                      int64_t* __restrict__ cnt_out) {
  const int b = blockIdx.x / row_groups, grp = blockIdx.x - b * row_groups;
  if (b >= B) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kCollateThreads / 32;
  const int64_t q = qidx[b];
  const int64_t off = offsets[q];
  const int64_t cnt = offsets[q + 1] - off;
  const bool sampled = cnt > L;
  const int n = static_cast<int>(sampled ? L : cnt);
  const int l0 = grp * kGatherRowsPerCta;
  const int l1 = l0 + kGatherRowsPerCta < L ? l0 + kGatherRowsPerCta : L;
  for (int l = l0 + warp; l < l1; l += nwarps) {
    float* __restrict__ dst = feat_out + (static_cast<size_t>(b) * L + l) * F;
    int64_t d = -1;
    if (l < n) {
      d = (sampled && sel) ? sel[static_cast<size_t>(b) * sel_ld + l] : l;
      d = d < 0 ? 0 : (d >= cnt ? cnt - 1 : d);
    }
    if (d >= 0) {
      const float* __restrict__ src = features + static_cast<size_t>(off + d) * F;
      if (vec_ok) {
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
        float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
        for (int i = lane; i < (F >> 2); i += 32) d4[i] = s4[i];
      } else {
        for (int i = lane; i < F; i += 32) dst[i] = src[i];
      }
      if (lane == 0) rel_out[static_cast<size_t>(b) * L + l] = relevance[off + d];
    } else {
      if (vec_ok) {
        float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
        for (int i = lane; i < (F >> 2); i += 32) d4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      } else {
        for (int i = lane; i < F; i += 32) dst[i] = 0.0f;
      }
      if (lane == 0) rel_out[static_cast<size_t>(b) * L + l] = 0;
    }
  }
  if (grp == 0 && threadIdx.x == 0) {
    n_out[b] = n;
    if (cnt_out) cnt_out[b] = cnt;
  }
}

// Sparse collation (svmrank.py:163-177, 198-203): the dataset holds its features in CSR form over the
// documents (indptr [N + 1], indices [nnz], values [nnz]); the batch is returned as the COO triplets
// (batch row, list position, feature) + values of a torch sparse tensor (B, L, F).  out_ptr [B * L + 1]
// is the exclusive scan of the per-row non-zero counts (computed by the caller from indptr), so that
// every (b, l) row knows where its entries go; one warp per row.
__global__ void __launch_bounds__(kCollateThreads)
collate_sparse_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices,
                      const float* __restrict__ values, const int64_t* __restrict__ relevance,
                      const int64_t* __restrict__ offsets, const int64_t* __restrict__ qidx,
                      const int64_t* __restrict__ sel, int sel_ld, const int64_t* __restrict__ out_ptr, int B, int L,
                      int64_t nnz_out, int64_t* __restrict__ coo_out, float* __restrict__ val_out,
                      int64_t* __restrict__ rel_out, int64_t* __restrict__ n_out) {
  const int lane = threadIdx.x & 31, nwarps = kCollateThreads / 32;
  const size_t rows = static_cast<size_t>(B) * L;
  for (size_t row = static_cast<size_t>(blockIdx.x) * nwarps + (threadIdx.x >> 5); row < rows;
       row += static_cast<size_t>(gridDim.x) * nwarps) {
    const int b = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<size_t>(b) * L);
    const int64_t q = qidx[b];
    const int64_t off = offsets[q];
    const int64_t cnt = offsets[q + 1] - off;
    const bool sampled = cnt > L;
    const int n = static_cast<int>(sampled ? L : cnt);
    int64_t rel = 0;
    if (l < n) {
      int64_t d = (sampled && sel) ? sel[static_cast<size_t>(b) * sel_ld + l] : l;
      d = d < 0 ? 0 : (d >= cnt ? cnt - 1 : d);
      const int64_t doc = off + d;
      rel = relevance[doc];
      const int64_t s0 = indptr[doc], s1 = indptr[doc + 1];
      const int64_t o0 = out_ptr[row];
      for (int64_t i = s0 + lane; i < s1; i += 32) {
        const int64_t o = o0 + (i - s0);
        coo_out[o] = b;
        coo_out[nnz_out + o] = l;
        coo_out[2 * nnz_out + o] = indices[i];
        val_out[o] = values[i];
      }
    }
    if (lane == 0) {
      rel_out[row] = rel;
      if (l == 0) n_out[b] = n;
    }
  }
}

}  // namespace ltr
