// ltr_common.cuh -- device helpers shared by the sm_100a kernels of libltr_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ltr {

constexpr float kLog2e = 1.4426950408889634f;   // 1 / ln 2
constexpr float kLn2 = 0.6931471805599453f;

// ---- MUFU wrappers (ex2 / lg2 / rcp run on the SFU pipe: 16 lanes / clk / SM) ------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- integer inputs of any width: int64 (what the reference passes), int32, int16, uint8 ----
__device__ __forceinline__ int load_int_clamped(const void* p, int bytes, size_t i) {
  if (bytes == 8) {
    long long v = reinterpret_cast<const long long*>(p)[i];
    v = v < -2147483647LL ? -2147483647LL : (v > 2147483647LL ? 2147483647LL : v);
    return static_cast<int>(v);
  }
  if (bytes == 4) return reinterpret_cast<const int*>(p)[i];
  if (bytes == 2) return reinterpret_cast<const short*>(p)[i];
  return reinterpret_cast<const unsigned char*>(p)[i];
}

__device__ __forceinline__ int load_n(const void* n, int n_bytes, int b, int L) {
  int v = load_int_clamped(n, n_bytes, b);
  return v < 0 ? 0 : (v > L ? L : v);
}

// ---- sort keys ------------------------------------------------------------------------
// 32-bit key whose ASCENDING unsigned order is DESCENDING score order.  -0 is folded into
// +0 and every NaN into the canonical positive NaN, which torch.argsort(descending=True)
// also places first.  0xFFFFFFFF is reserved for padding (sorts last).
__device__ __forceinline__ uint32_t desc_key_f32(float x) {
  x = x + 0.0f;
  uint32_t u = __float_as_uint(x);
  if (x != x) u = 0x7fc00000u;
  // ascending key = u ^ (sign ? 0xffffffff : 0x80000000); descending = its complement
  return u ^ (~static_cast<uint32_t>(static_cast<int>(u) >> 31) & 0x7fffffffu);
}
// Same for an int32 relevance grade (used for the ideal ranking of _max_dcg / ndcg).
__device__ __forceinline__ uint32_t desc_key_i32(int v) {
  return ~(static_cast<uint32_t>(v) ^ 0x80000000u);
}
constexpr uint32_t kPadKey = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t pack_key(uint32_t key, int idx) {
  return (static_cast<uint64_t>(key) << 32) | static_cast<uint32_t>(idx);
}

// ---- CTA-wide bitonic sort of P (power of two) 64-bit keys in shared memory ----------
// Ascending.  All threads of the CTA must call it; ends with a barrier.
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* keys, int P) {
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const bool up = (i & k) == 0;
        const uint64_t a = keys[i], b = keys[l];
        if ((a > b) == up) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
  }
}

// ---- reductions -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Sum over the CTA; `red` is >= 33 floats of shared memory.  Result returned to all threads.
__device__ __forceinline__ float cta_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    float x = lane < nw ? red[lane] : 0.0f;
    x = warp_sum(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}

// Minimum over the CTA; `red` is >= 33 floats of shared memory.  Result returned to all threads.
__device__ __forceinline__ float cta_min(float v, float* red) {
  v = -warp_max(-v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    float x = lane < nw ? red[lane] : INFINITY;
    x = -warp_max(-x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}

// gain of a relevance grade: 2^rel - 1 in float32 (pairwise_lambda.py:224-225, dcg.py:91-92)
__device__ __forceinline__ float exp_gain_f32(int rel) {
  return exp2f(static_cast<float>(rel)) - 1.0f;
}

}  // namespace ltr
