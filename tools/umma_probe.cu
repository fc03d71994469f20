// umma_probe.cu -- one-CTA experiment harness for tcgen05.mma kind::tf32 operand layouts.
// The host passes raw shared-memory images of A and B and the descriptor fields; the kernel copies the
// images to 1024-byte aligned shared memory, issues `ksteps` MMAs (descriptor start addresses advanced by
// a_step / b_step bytes per step) and dumps the accumulator: out[lane * 256 + col] for 128 TMEM lanes.
// nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/libumma_probe.so tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include "../pytorchltr_b200/csrc/ltr_mlp_scorer.cuh"

using namespace ltr;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const unsigned char* a_img, int a_bytes, const unsigned char* b_img, int b_bytes, uint32_t idesc,
                  uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout, uint32_t b_lbo, uint32_t b_sbo,
                  uint32_t b_layout, int ksteps, int a_step, int b_step, int ncols, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  unsigned char* as = base;
  unsigned char* bs = base + ((a_bytes + 1023) / 1024) * 1024;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < a_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(as)[i] = reinterpret_cast<const uint32_t*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(bs)[i] = reinterpret_cast<const uint32_t*>(b_img)[i];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // clear the accumulator region first so that untouched lanes read as a marker
  if (threadIdx.x == 0) {
    for (int k = 0; k < ksteps; ++k)
      umma_tf32(tmem, umma_desc(smem_u32(as) + k * a_step, a_lbo, a_sbo, a_layout),
                umma_desc(smem_u32(bs) + k * b_step, b_lbo, b_sbo, b_layout), idesc, k > 0);
    umma_commit(&bar);
  }
  mbar_wait_guarded(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) out[threadIdx.x * 256 + c0 + k] = __uint_as_float(v[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

extern "C" int umma_probe(const unsigned char* a_img, int a_bytes, const unsigned char* b_img, int b_bytes,
                          uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout, uint32_t b_lbo,
                          uint32_t b_sbo, uint32_t b_layout, int ksteps, int a_step, int b_step, int ncols,
                          float* out) {
  const int smem = 1024 + ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  umma_probe_kernel<<<1, 128, smem>>>(a_img, a_bytes, b_img, b_bytes, idesc, a_lbo, a_sbo, a_layout, b_lbo, b_sbo,
                                      b_layout, ksteps, a_step, b_step, ncols, out);
  e = cudaDeviceSynchronize();
  return static_cast<int>(e);
}
