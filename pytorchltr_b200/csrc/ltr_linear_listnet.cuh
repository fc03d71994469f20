// ltr_linear_listnet.cuh -- fused linear scorer + ListNet loss + gradients.
//
// The caller side of the loss path (examples/01-basic-usage.py:44,72, getting-started.rst:42-51) is
//     scores = torch.nn.Linear(F, 1)(features)          # (B, L, F) -> (B, L, 1)
//     loss_fn(scores, relevance, n).mean().backward()   # + d/dweight = features^T dscores
// which reads the (B, L, F) feature tensor twice (forward GEMV, backward GEMV): 2 x 891 MB at the
// MSLR-WEB30K shape (8192, 200, 136), next to 26 MB for the loss itself.  This kernel reads it ONCE:
//
//   one persistent CTA per SM; the L x F feature block of a query (109 KB at (200, 136)) and its
//   relevance row are staged in shared memory by TMA bulk copies (cp.async.bulk + mbarrier), double
//   buffered so that the next query's block streams in while the current one is processed;
//     scores   s_l = w . x_l + b        two threads per row, 128-bit loads (conflict free at F = 136)
//     ListNet  loss, d_l = softmax(s)_l - softmax(rel)_l over the valid documents (ltr_listnet)
//     weight gradient  += sum_l d_l x_l  out of the SAME shared-memory block: 4 features per
//                      thread (128-bit loads), the rows dealt to 512 / (F / 4) thread slices,
//                      summed over the slices in a fixed order (deterministic)
//   and writes loss, dscores (6.5 MB) plus the per-query weight gradient G_b = sum_l d_l x_l (B x (F+1)
//   floats, 4.5 MB): the backward pass for ANY upstream gradient g is then the small weighted column
//   sum sum_b g_b G_b (ltr_linear_listnet_backward), never a second pass over the features.
//
// HBM traffic: 4 L F + 8 L + 8 read, 8 L + 4 written per query -- the roofline of the kernel.
#pragma once

#include "ltr_common.cuh"
#include "ltr_sm100.h"

namespace ltr {

constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxF = 1024;         // F / 4 feature groups must not exceed the CTA
constexpr int kFusedBwdCtas = 296;       // rows of the backward pass's partial-sum workspace

__host__ __device__ inline size_t fused_align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }
__host__ __device__ inline size_t fused_stage_bytes(int L, int F, int rel_bytes) {
  return fused_align16(4u * static_cast<size_t>(L) * F) + fused_align16(static_cast<size_t>(rel_bytes) * L);
}
// [slices][F + 4] scratch for the per-query reduction of the row slices
__host__ __device__ inline size_t fused_part_bytes(int F) {
  const size_t slices = kFusedThreads / (static_cast<size_t>(F) / 4);
  return fused_align16(4u * slices * (static_cast<size_t>(F) + 4));
}
__host__ __device__ inline size_t fused_front_bytes(int L, int F, int rel_bytes, int nbuf) {
  return nbuf * fused_stage_bytes(L, F, rel_bytes);
}
__host__ __device__ inline size_t fused_smem_bytes(int L, int F, int rel_bytes, int nbuf) {
  return fused_front_bytes(L, F, rel_bytes, nbuf) + fused_part_bytes(F) + fused_align16(4u * F) +
         2u * fused_align16(4u * L) + 4u * 128 + 16u;
}

// CTA-wide reductions with ONE barrier each: every warp publishes its partial results, and after
// the barrier every warp folds the kFusedWarps partials itself (3 loads + a 16-lane butterfly).
// The two functions use disjoint halves of `red` (128 floats) and alternate, so neither needs a
// barrier before overwriting its slots: the other function's barrier lies in between.
__device__ __forceinline__ void cta_max2(float& a, float& b, float* red) {
  a = warp_max(a); b = warp_max(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp] = a; red[16 + warp] = b; }
  __syncthreads();
  float x = lane < kFusedWarps ? red[lane] : -INFINITY;
  float y = lane < kFusedWarps ? red[16 + lane] : -INFINITY;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    y = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, o));
  }
  a = __shfl_sync(0xffffffffu, x, 0);
  b = __shfl_sync(0xffffffffu, y, 0);
}
__device__ __forceinline__ void cta_sum3(float& a, float& b, float& c, float* red) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* r = red + 64;
  if (lane == 0) { r[warp] = a; r[16 + warp] = b; r[32 + warp] = c; }
  __syncthreads();
  float x = lane < kFusedWarps ? r[lane] : 0.0f;
  float y = lane < kFusedWarps ? r[16 + lane] : 0.0f;
  float z = lane < kFusedWarps ? r[32 + lane] : 0.0f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
    z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  a = __shfl_sync(0xffffffffu, x, 0);
  b = __shfl_sync(0xffffffffu, y, 0);
  c = __shfl_sync(0xffffffffu, z, 0);
}

template <int NBUF>
__global__ void __launch_bounds__(kFusedThreads, 1)
linear_listnet_kernel(const float* __restrict__ feat, const float* __restrict__ weight,
                      const float* __restrict__ bias, const void* __restrict__ rel, int rel_bytes,
                      const void* __restrict__ n, int n_bytes, int B, int L, int F,
                      float* __restrict__ scores_out, float* __restrict__ loss_out,
                      float* __restrict__ dscores_out, float* __restrict__ qgrad,
                      float* __restrict__ loss_sum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t stage_bytes = fused_stage_bytes(L, F, rel_bytes);
  const size_t x_bytes = fused_align16(4u * static_cast<size_t>(L) * F);
  unsigned char* p = smem_raw + fused_front_bytes(L, F, rel_bytes, NBUF);
  float* part = reinterpret_cast<float*>(p);           p += fused_part_bytes(F);
  float* w_s = reinterpret_cast<float*>(p);            p += fused_align16(4u * F);
  float* sc = reinterpret_cast<float*>(p);             p += fused_align16(4u * L);
  float* dd = reinterpret_cast<float*>(p);             p += fused_align16(4u * L);
  float* red = reinterpret_cast<float*>(p);            p += 4u * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p);

  const int tid = threadIdx.x;
  const int grid = gridDim.x;
  const uint32_t row_y_bytes = static_cast<uint32_t>(rel_bytes) * L;
  const size_t feat_row_bytes = 4u * static_cast<size_t>(L) * F;

  auto issue_query = [&](int q, int buf) {
    unsigned char* dst = smem_raw + buf * stage_bytes;
    uint64_t* bar = bars + buf;
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(feat_row_bytes) + row_y_bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(feat) + static_cast<size_t>(q) * feat_row_bytes;
    constexpr uint32_t kChunk = 32768u;
    for (size_t off = 0; off < feat_row_bytes; off += kChunk) {
      const uint32_t bytes = static_cast<uint32_t>(feat_row_bytes - off < kChunk ? feat_row_bytes - off : kChunk);
      tma_load_1d(dst + off, src + off, bytes, bar);
    }
    tma_load_1d(dst + x_bytes, static_cast<const unsigned char*>(rel) + static_cast<size_t>(q) * row_y_bytes,
                row_y_bytes, bar);
  };

  if (tid == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  for (int f = tid; f < F; f += kFusedThreads) w_s[f] = weight[f];
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < NBUF; ++i) {
      const int q = blockIdx.x + i * grid;
      if (q < B) issue_query(q, i);
    }
  }
  const float b0 = bias ? bias[0] : 0.0f;

  // gradient phase: thread = (feature group of 4, row slice)
  const int ngroups = F >> 2;
  const int slices = kFusedThreads / ngroups;
  const int my_g = tid % ngroups, my_slice = tid / ngroups;
  const bool grad_thread = my_slice < slices;
  const int ldp = F + 4;

  int nb_next = static_cast<int>(blockIdx.x) < B ? load_n(n, n_bytes, blockIdx.x, L) : 0;
  int it = 0;
  for (int b = blockIdx.x; b < B; b += grid, ++it) {
    const int buf = it % NBUF;
    const int nb = nb_next;
    if (b + grid < B) nb_next = load_n(n, n_bytes, b + grid, L);
    const float* Xs = reinterpret_cast<const float*>(smem_raw + buf * stage_bytes);
    const unsigned char* Ys = smem_raw + buf * stage_bytes + x_bytes;
    mbar_wait(bars + buf, (it / NBUF) & 1);

    // ---- scores: two threads per row (feature groups split in halves), 128-bit loads.  Rows are F
    // floats apart (8 l mod 32 banks for F = 136) and the second half starts (ngroups + 1) / 2 groups
    // further (4 banks further for F = 136): the 4 rows x 2 halves of a quarter warp hit 8 different
    // 4-bank groups -----------------------------------------------------------------------------------
    const int rows = scores_out ? L : nb;
    {
      const float4* w4 = reinterpret_cast<const float4*>(w_s);
      const int half = tid & 1;
      const int gsplit = (ngroups + 1) >> 1;
      const int g0 = half ? gsplit : 0, g1 = half ? ngroups : gsplit;
      const int rows2 = (rows + 1) & ~1;   // both lanes of a pair run the loop (shuffle below)
      for (int l = tid >> 1; l < rows2; l += kFusedThreads >> 1) {
        const int lr = l < rows ? l : rows - 1;
        const float4* xr = reinterpret_cast<const float4*>(Xs + static_cast<size_t>(lr) * F);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
        for (int g = g0; g < g1; ++g) {
          const float4 x = xr[g], w = w4[g];
          a0 = fmaf(x.x, w.x, a0); a1 = fmaf(x.y, w.y, a1);
          a2 = fmaf(x.z, w.z, a2); a3 = fmaf(x.w, w.w, a3);
        }
        float sum = (a0 + a1) + (a2 + a3);
        sum += __shfl_xor_sync(0x3u << (threadIdx.x & 30), sum, 1);   // the pair only: other lanes may have left
        if (half == 0 && l < rows) sc[l] = sum + b0;
      }
    }
    __syncthreads();

    // ---- ListNet over the valid documents (same arithmetic as listnet_reg_kernel); the grades wait
    // as floats in dd[], which the last pass overwrites with d loss / d score ------------------------
    float ms = -INFINITY, my = -INFINITY;
    for (int l = tid; l < nb; l += kFusedThreads) {
      const float y = static_cast<float>(load_int_clamped(Ys, rel_bytes, l));
      dd[l] = y;
      ms = fmaxf(ms, sc[l]);
      my = fmaxf(my, y);
    }
    cta_max2(ms, my, red);
    float zs = 0.0f, zy = 0.0f, a = 0.0f;
    for (int l = tid; l < nb; l += kFusedThreads) {
      const float ds = sc[l] - ms;
      const float es = ex2_approx(ds * kLog2e);
      const float ey = ex2_approx((dd[l] - my) * kLog2e);
      zs += es; zy += ey;
      a = fmaf(ey, ds, a);
    }
    cta_sum3(zs, zy, a, red);
    const float loss = nb > 0 ? logf(zs) - a / zy : 0.0f;
    const float izs = nb > 0 ? 1.0f / zs : 0.0f, izy = nb > 0 ? 1.0f / zy : 0.0f;
    const size_t base = static_cast<size_t>(b) * L;
    for (int l = tid; l < L; l += kFusedThreads) {
      float d = 0.0f;
      const float s = l < rows ? sc[l] : 0.0f;
      if (l < nb) {
        const float es = ex2_approx((s - ms) * kLog2e);
        const float ey = ex2_approx((dd[l] - my) * kLog2e);
        d = es * izs - ey * izy;
      }
      dd[l] = d;
      if (scores_out) scores_out[base + l] = s;
      if (dscores_out) dscores_out[base + l] = d;
    }
    if (tid == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
    }
    __syncthreads();

    // ---- per-query weight gradient out of the same block: G_b[f] = sum_l d_l x_l[f] --------------------
    if (grad_thread) {
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float accb = 0.0f;
      const float4* X4 = reinterpret_cast<const float4*>(Xs);
#pragma unroll 4
      for (int l = my_slice; l < nb; l += slices) {
        const float d = dd[l];
        const float4 x = X4[static_cast<size_t>(l) * ngroups + my_g];
        acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y);
        acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
        accb += d;
      }
      *reinterpret_cast<float4*>(part + my_slice * ldp + 4 * my_g) = acc;
      if (my_g == 0) part[my_slice * ldp + F] = accb;
    }
    __syncthreads();   // the block is consumed: its buffer takes the query NBUF rounds ahead
    if (tid == 0) {
      const int q = b + NBUF * grid;
      if (q < B) {
        fence_proxy_async();
        issue_query(q, buf);
      }
    }
    // the slices summed in a fixed order (deterministic); column F is d loss_b / d bias
    for (int f = tid; f <= F; f += kFusedThreads) {
      float s = 0.0f;
      for (int sl = 0; sl < slices; ++sl) s += part[sl * ldp + f];
      qgrad[static_cast<size_t>(b) * (F + 1) + f] = s;
    }
  }
}

// ---- any shape: feature blocks larger than shared memory, F % 4 != 0, unaligned buffers ----------------
// Same outputs, same arithmetic for the ListNet part; the L x F block of a query is streamed from global
// memory twice inside ONE kernel -- scores (one warp per row, coalesced), then, after the loss, the weight
// gradient (thread = (feature, row slice)) -- and the second pass finds the block in L2 (a query's block is
// at most a few MB), so the HBM traffic stays one pass over the features.  One CTA per query, grid-stride.
__host__ __device__ inline size_t fused_tiled_smem_bytes(int L, int F) {
  const size_t fc = F < kFusedThreads ? F : kFusedThreads;
  const size_t slices = kFusedThreads / fc;
  return fused_align16(4u * static_cast<size_t>(F)) + 2u * fused_align16(4u * static_cast<size_t>(L)) +
         fused_align16(4u * slices * fc) + 4u * 128;
}

__global__ void __launch_bounds__(kFusedThreads, 2)
linear_listnet_tiled_kernel(const float* __restrict__ feat, const float* __restrict__ weight,
                            const float* __restrict__ bias, const void* __restrict__ rel, int rel_bytes,
                            const void* __restrict__ n, int n_bytes, int B, int L, int F, int vec_ok,
                            float* __restrict__ scores_out, float* __restrict__ loss_out,
                            float* __restrict__ dscores_out, float* __restrict__ qgrad,
                            float* __restrict__ loss_sum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int fc = F < kFusedThreads ? F : kFusedThreads;      // features handled per gradient pass
  const int slices = kFusedThreads / fc;
  unsigned char* p = smem_raw;
  float* w_s = reinterpret_cast<float*>(p);            p += fused_align16(4u * static_cast<size_t>(F));
  float* sc = reinterpret_cast<float*>(p);             p += fused_align16(4u * static_cast<size_t>(L));
  float* dd = reinterpret_cast<float*>(p);             p += fused_align16(4u * static_cast<size_t>(L));
  float* part = reinterpret_cast<float*>(p);           p += fused_align16(4u * static_cast<size_t>(slices) * fc);
  float* red = reinterpret_cast<float*>(p);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int f = tid; f < F; f += kFusedThreads) w_s[f] = weight[f];
  const float b0 = bias ? bias[0] : 0.0f;
  __syncthreads();

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int nb = load_n(n, n_bytes, b, L);
    const float* __restrict__ X = feat + static_cast<size_t>(b) * L * F;
    const size_t base = static_cast<size_t>(b) * L;
    const int rows = scores_out ? L : nb;
    // ---- scores: one warp per row ----------------------------------------------------------------------
    for (int l = warp; l < rows; l += kFusedWarps) {
      const float* __restrict__ xr = X + static_cast<size_t>(l) * F;
      float acc = 0.0f;
      if (vec_ok) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        const float4* w4 = reinterpret_cast<const float4*>(w_s);
        for (int g = lane; g < (F >> 2); g += 32) {
          const float4 x = x4[g], w = w4[g];
          acc = fmaf(x.x, w.x, acc); acc = fmaf(x.y, w.y, acc);
          acc = fmaf(x.z, w.z, acc); acc = fmaf(x.w, w.w, acc);
        }
      } else {
        for (int f = lane; f < F; f += 32) acc = fmaf(xr[f], w_s[f], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) sc[l] = acc + b0;
    }
    __syncthreads();

    // ---- ListNet (same arithmetic as linear_listnet_kernel) ------------------------------------------
    float ms = -INFINITY, my = -INFINITY;
    for (int l = tid; l < nb; l += kFusedThreads) {
      const float y = static_cast<float>(load_int_clamped(rel, rel_bytes, base + l));
      dd[l] = y;
      ms = fmaxf(ms, sc[l]);
      my = fmaxf(my, y);
    }
    cta_max2(ms, my, red);
    float zs = 0.0f, zy = 0.0f, a = 0.0f;
    for (int l = tid; l < nb; l += kFusedThreads) {
      const float ds = sc[l] - ms;
      const float es = ex2_approx(ds * kLog2e);
      const float ey = ex2_approx((dd[l] - my) * kLog2e);
      zs += es; zy += ey;
      a = fmaf(ey, ds, a);
    }
    cta_sum3(zs, zy, a, red);
    const float loss = nb > 0 ? logf(zs) - a / zy : 0.0f;
    const float izs = nb > 0 ? 1.0f / zs : 0.0f, izy = nb > 0 ? 1.0f / zy : 0.0f;
    float dsum = 0.0f, dummy1 = 0.0f, dummy2 = 0.0f;
    for (int l = tid; l < L; l += kFusedThreads) {
      float d = 0.0f;
      const float s = l < rows ? sc[l] : 0.0f;
      if (l < nb) {
        const float es = ex2_approx((s - ms) * kLog2e);
        const float ey = ex2_approx((dd[l] - my) * kLog2e);
        d = es * izs - ey * izy;
      }
      dd[l] = d;
      dsum += d;
      if (scores_out) scores_out[base + l] = s;
      if (dscores_out) dscores_out[base + l] = d;
    }
    __syncthreads();                    // cta_max2's slots may be rewritten: everyone is past cta_sum3
    cta_max2(dummy1, dummy2, red);      // (a barrier with the reduction layout the next call expects)
    cta_sum3(dsum, dummy1, dummy2, red);
    if (tid == 0) {
      loss_out[b] = loss;
      if (loss_sum) atomicAdd(loss_sum, loss);
      qgrad[static_cast<size_t>(b) * (F + 1) + F] = dsum;    // d loss_b / d bias
    }

    // ---- weight gradient G_b[f] = sum_l d_l x_l[f]: the block comes from L2 this time -------------------
    for (int c0 = 0; c0 < F; c0 += fc) {
      const int f = tid % fc, slice = tid / fc;
      float acc = 0.0f;
      if (slice < slices && c0 + f < F) {
        const float* __restrict__ col = X + c0 + f;
#pragma unroll 4
        for (int l = slice; l < nb; l += slices) acc = fmaf(dd[l], col[static_cast<size_t>(l) * F], acc);
      }
      if (slice < slices) part[slice * fc + f] = acc;
      __syncthreads();
      if (tid < fc && c0 + tid < F) {
        float s = 0.0f;
        for (int sl = 0; sl < slices; ++sl) s += part[sl * fc + tid];     // fixed order: deterministic
        qgrad[static_cast<size_t>(b) * (F + 1) + c0 + tid] = s;
      }
      __syncthreads();
    }
  }
}

// out[f] = sum_b g[b * g_stride] * qgrad[b, f] for a chunk of queries per CTA -> partials[cta, f]
__global__ void __launch_bounds__(256)
weighted_colsum_kernel(const float* __restrict__ qgrad, const float* __restrict__ g, int g_stride, int B,
                       int cols, float* __restrict__ partials) {
  const int per = (B + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * per, b1 = min(B, b0 + per);
  for (int f = threadIdx.x; f < cols; f += blockDim.x) {
    float s = 0.0f;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) s = fmaf(g[static_cast<size_t>(b) * g_stride], qgrad[static_cast<size_t>(b) * cols + f], s);
    partials[static_cast<size_t>(blockIdx.x) * cols + f] = s;
  }
}

// dweight[f] (f < F) and dbias (f == F): sum of the per-CTA partials in a fixed order
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int rows, int cols, float* __restrict__ dweight,
                       float* __restrict__ dbias) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < cols; f += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += static_cast<double>(partials[static_cast<size_t>(r) * cols + f]);
    if (f < cols - 1) dweight[f] = static_cast<float>(s);
    else if (dbias) dbias[0] = static_cast<float>(s);
  }
}

}  // namespace ltr
