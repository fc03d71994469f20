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

}  // namespace ltr
