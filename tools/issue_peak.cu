// issue_peak.cu -- measured issue / pipe ceilings of one B200 for the O(L^2) pair kernels
// ("verify with a micro-benchmark the way MEASURED_PEAKS did for HBM").
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/issue_peak tools/issue_peak.cu
//   tools/issue_peak > profiles/issue_peaks.json        (on the GPU box)
//
// Every stream is a dependent-free set of 8 register chains per thread, 1024 threads per SM x 2
// CTAs per SM, timed with CUDA events after a warm-up launch; the result is lane-operations per
// second over the whole device (148 SMs) and, next to it, the value SM count x clock x pipe width
// would predict.  Streams:
//   mufu_rcp / mufu_lg2 / mufu_ex2 / mufu_rcp_lg2   XU pipe (the pair kernels spend rcp + lg2 per pair)
//   ffma / fadd / fmul / fmnmx                      scalar FP32 on the fma / alu pipes
//   ffma2 / fadd2 / fmul2                           packed f32x2 (two lane-ops per issue slot)
//   mix_mufu2_ffmaK / mix_mufu2_ffma2K              2 MUFU + K FP32 per "pair": how much FP32 issues under a
//                                                   saturated XU pipe (the co-bound regime of the kernels)
//   body_*                                          the pair bodies themselves on register operands:
//     body_scalar   LambdaNDCGLoss2 pair as shipped in round 1 (2 MUFU + 11 FP32)
//     body_packed   two columns per instruction with f32x2 arithmetic
//     body_packed_rcp2   + one reciprocal for two pairs (1 / (p0 p1) and two multiplies)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e__ = (x);                                                           \
    if (e__ != cudaSuccess) {                                                        \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ float rcp_(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fma_(float a, float b, float c) { float y; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float add_(float a, float b) { float y; asm volatile("add.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float mul_(float a, float b) { float y; asm volatile("mul.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float max_(float a, float b) { float y; asm volatile("max.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float2 fma2_(float2 a, float2 b, float2 c) {
  unsigned long long y;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y)
               : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
                 "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&y);
}
__device__ __forceinline__ float2 add2_(float2 a, float2 b) {
  unsigned long long y;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(y)
               : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&y);
}
__device__ __forceinline__ float2 mul2_(float2 a, float2 b) {
  unsigned long long y;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(y)
               : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&y);
}

enum Op : int {
  OP_RCP, OP_LG2, OP_EX2, OP_RCP_LG2, OP_FFMA, OP_FADD, OP_FMUL, OP_FMNMX, OP_FFMA2, OP_FADD2, OP_FMUL2,
  OP_FFMA2_FMNMX, OP_FFMA2_MUFU, OP_MIX_S4, OP_MIX_S8, OP_MIX_S12, OP_MIX_P2, OP_MIX_P4, OP_MIX_P6, OP_MIX_P8,
  OP_BODY_SCALAR, OP_BODY_PACKED, OP_BODY_PACKED_RCP2, OP_COUNT
};

constexpr int kChains = 8;

// one "unit" of a stream = what is counted once per chain per iteration
template <int OP>
__global__ void __launch_bounds__(512, 2) stream_kernel(float* out, int iters, float seed) {
  float x[kChains];
  float2 x2[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) {
    x[k] = seed + 0.001f * static_cast<float>(threadIdx.x + k);
    x2[k] = make_float2(x[k], x[k] + 0.5f);
  }
  const float b = 1.0f + seed * 1e-6f, c = seed * 1e-7f;
  const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
  // operands of the pair bodies (rows stay in registers, columns change every step in the kernels)
  const float ra = 0.7f + seed * 1e-3f, re = 0.1f * seed, rg = 0.3f * seed;
  const float2 ra2 = make_float2(ra, ra), re2 = make_float2(re, re), rg2 = make_float2(rg, rg);
  const float2 one2 = make_float2(1.0f, 1.0f), two2 = make_float2(2.0f, 2.0f), neg1 = make_float2(-1.0f, -1.0f);
  float lacc = 0.0f, racc = 0.0f;
  float2 lacc2 = make_float2(0.0f, 0.0f), racc2 = make_float2(0.0f, 0.0f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
      if constexpr (OP == OP_RCP) x[k] = add_(rcp_(x[k]), c);   // (+ c: ptxas folds rcp(rcp(x)) to x)
      else if constexpr (OP == OP_LG2) x[k] = lg2_(x[k]);
      else if constexpr (OP == OP_EX2) x[k] = ex2_(x[k]);
      else if constexpr (OP == OP_RCP_LG2) { x[k] = add_(rcp_(x[k]), c); x2[k].x = lg2_(x2[k].x); }
      else if constexpr (OP == OP_FFMA) x[k] = fma_(x[k], b, c);
      else if constexpr (OP == OP_FADD) x[k] = add_(x[k], c);
      else if constexpr (OP == OP_FMUL) x[k] = mul_(x[k], b);
      else if constexpr (OP == OP_FMNMX) x[k] = max_(x[k], c);
      else if constexpr (OP == OP_FFMA2) x2[k] = fma2_(x2[k], b2, c2);
      else if constexpr (OP == OP_FADD2) x2[k] = add2_(x2[k], c2);
      else if constexpr (OP == OP_FMUL2) x2[k] = mul2_(x2[k], b2);
      else if constexpr (OP == OP_FFMA2_FMNMX) { x2[k] = fma2_(x2[k], b2, c2); x[k] = max_(x[k], c); }
      else if constexpr (OP == OP_FFMA2_MUFU) { x2[k] = fma2_(x2[k], b2, c2); x2[k] = fma2_(x2[k], b2, c2); x2[k] = fma2_(x2[k], b2, c2); x2[k] = fma2_(x2[k], b2, c2); x[k] = add_(lg2_(x[k]), c); }
      else if constexpr (OP == OP_MIX_S4 || OP == OP_MIX_S8 || OP == OP_MIX_S12) {
        constexpr int K = OP == OP_MIX_S4 ? 4 : (OP == OP_MIX_S8 ? 8 : 12);
        const float r = rcp_(x[k]);
        const float l = lg2_(x[k]);
        float t = r;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) t = fma_(t, b, l);
        x[k] = fma_(t, b, c);
      } else if constexpr (OP == OP_MIX_P2 || OP == OP_MIX_P4 || OP == OP_MIX_P6 || OP == OP_MIX_P8) {
        // two pairs per unit: 4 MUFU + K packed instructions
        constexpr int K = OP == OP_MIX_P2 ? 2 : (OP == OP_MIX_P4 ? 4 : (OP == OP_MIX_P6 ? 6 : 8));
        float2 r, l;
        r.x = rcp_(x2[k].x); r.y = rcp_(x2[k].y);
        l.x = lg2_(x2[k].x); l.y = lg2_(x2[k].y);
        float2 t = r;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) t = fma2_(t, b2, l);
        x2[k] = fma2_(t, b2, c2);
      } else if constexpr (OP == OP_BODY_SCALAR) {
        // round-1 pair_once<TW_DELTA, FACTORED>: column (cx, ce, cg) and the window value dw vary per pair
        const float cx = x[k], ce = x2[k].x, cg = x2[k].y, dw = b;
        const float gd = add_(rg, -cg);
        const float ws = mul_(dw, gd);
        const float p = fma_(ra, cx, 1.0f);
        const float r = rcp_(p);
        const float lg = lg2_(p);
        const float K = max_(ws, 0.0f);
        lacc = fma_(fabsf(ws), lg, lacc);
        lacc = fma_(add_(K, -ws), add_(re, -ce), lacc);
        const float v = fma_(fabsf(ws), r, -K);
        racc = add_(racc, v);
        x[k] = add_(x[k], -v * 1e-9f);          // the column accumulator (keeps the chain alive)
      } else if constexpr (OP == OP_BODY_PACKED || OP == OP_BODY_PACKED_RCP2) {
        // two columns of one row per unit; halved weights h = 0.5 ws, a = |h|:
        //   loss / 2 += a lg + (a - h) (e_i - e_j) / 2,   v = a (2 r - 1) - h
        const float2 cx = x2[k], ce = make_float2(x[k], x[k] + c), cg = make_float2(x2[k].y, x2[k].x), dwh = b2;
        const float2 gd = __fadd2_rn(rg2, make_float2(-cg.x, -cg.y));
        const float2 h = __fmul2_rn(dwh, gd);
        const float2 a = make_float2(fabsf(h.x), fabsf(h.y));
        const float2 p = __ffma2_rn(ra2, cx, one2);
        float2 r, lg;
        if constexpr (OP == OP_BODY_PACKED_RCP2) {
          const float R = rcp_(p.x * p.y);
          r.x = R * p.y; r.y = R * p.x;
        } else {
          r.x = rcp_(p.x); r.y = rcp_(p.y);
        }
        lg.x = lg2_(p.x); lg.y = lg2_(p.y);
        lacc2 = __ffma2_rn(a, lg, lacc2);
        const float2 d = __fadd2_rn(a, make_float2(-h.x, -h.y));
        const float2 ed = __fadd2_rn(re2, make_float2(-ce.x, -ce.y));
        lacc2 = __ffma2_rn(d, ed, lacc2);
        const float2 t = __ffma2_rn(r, two2, neg1);
        const float2 v = __ffma2_rn(a, t, make_float2(-h.x, -h.y));
        racc2 = __fadd2_rn(racc2, v);
        x2[k] = fma2_(v, make_float2(-1e-9f, -1e-9f), x2[k]);   // the column accumulator
      }
    }
  }
  float s = lacc + racc + lacc2.x + lacc2.y + racc2.x + racc2.y;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += x[k] + x2[k].x + x2[k].y;
  if (s == 12345.678f) out[threadIdx.x] = s;   // never true: defeats dead-code elimination
}

struct Stream {
  const char* name;
  const char* what;
  double lane_ops_per_unit;     // lane-operations counted per unit (for the rate column)
  const char* counted;
  double predicted_per_clk_sm;  // lane-ops / clk / SM from the nominal pipe width (0 = none)
};

template <int OP>
double time_stream(int sms, int iters, float* dout) {
  const int grid = sms * 2, threads = 512;
  stream_kernel<OP><<<grid, threads>>>(dout, iters / 8, 1.5f);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    stream_kernel<OP><<<grid, threads>>>(dout, iters, 1.5f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
  // units executed: grid * threads * chains * iters
  return static_cast<double>(grid) * threads * kChains * static_cast<double>(iters) / (best * 1e-3);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  float* dout = nullptr;
  CK(cudaMalloc(&dout, 4096));
  const int iters = 4096;
  const double clk = khz * 1e3;

  static const Stream streams[OP_COUNT] = {
      {"mufu_rcp", "rcp.approx.ftz.f32", 1, "MUFU lane-ops", 16},
      {"mufu_lg2", "lg2.approx.ftz.f32", 1, "MUFU lane-ops", 16},
      {"mufu_ex2", "ex2.approx.ftz.f32", 1, "MUFU lane-ops", 16},
      {"mufu_rcp_lg2", "rcp + lg2 (the two MUFU of one factored pair)", 2, "MUFU lane-ops", 16},
      {"ffma", "fma.rn.f32 (3 register operands)", 1, "FP32 lane-ops", 128},
      {"fadd", "add.f32", 1, "FP32 lane-ops", 128},
      {"fmul", "mul.f32", 1, "FP32 lane-ops", 128},
      {"fmnmx", "max.f32", 1, "FP32 lane-ops", 64},
      {"ffma2", "fma.rn.f32x2", 2, "FP32 lane-ops", 128},
      {"fadd2", "add.rn.f32x2", 2, "FP32 lane-ops", 128},
      {"fmul2", "mul.rn.f32x2", 2, "FP32 lane-ops", 128},
      {"ffma2_plus_fmnmx", "1 fma.rn.f32x2 + 1 max.f32 (does the alu pipe run beside the packed FP32 datapath?)", 3, "FP32 lane-ops", 128},
      {"ffma2x4_plus_mufu", "4 fma.rn.f32x2 + 1 lg2 + 1 add (8 FP32 lane-ops per MUFU: both pipes saturated?)", 9, "FP32 lane-ops", 128},
      {"mix_mufu2_ffma4", "per pair: rcp + lg2 + 4 fma.f32", 1, "pairs", 8},
      {"mix_mufu2_ffma8", "per pair: rcp + lg2 + 8 fma.f32", 1, "pairs", 8},
      {"mix_mufu2_ffma12", "per pair: rcp + lg2 + 12 fma.f32", 1, "pairs", 8},
      {"mix_mufu2_ffma2x2", "per 2 pairs: 2 rcp + 2 lg2 + 2 fma.f32x2", 2, "pairs", 8},
      {"mix_mufu2_ffma2x4", "per 2 pairs: 2 rcp + 2 lg2 + 4 fma.f32x2", 2, "pairs", 8},
      {"mix_mufu2_ffma2x6", "per 2 pairs: 2 rcp + 2 lg2 + 6 fma.f32x2", 2, "pairs", 8},
      {"mix_mufu2_ffma2x8", "per 2 pairs: 2 rcp + 2 lg2 + 8 fma.f32x2", 2, "pairs", 8},
      {"body_scalar", "LambdaNDCGLoss2 pair body of round 1: 2 MUFU + 11 scalar FP32", 1, "pairs", 8},
      {"body_packed", "two columns per instruction: 4 MUFU + 11 f32x2 per 2 pairs (abs / neg / broadcast are operand modifiers)", 2, "pairs", 8},
      {"body_packed_rcp2", "same with one rcp per 2 pairs: 3 MUFU + 11 f32x2 + 3 scalar mul per 2 pairs", 2, "pairs", 32.0 / 3.0},
  };
  double rate[OP_COUNT];
#define RUN(OP) rate[OP] = time_stream<OP>(sms, iters, dout) * streams[OP].lane_ops_per_unit
  RUN(OP_RCP); RUN(OP_LG2); RUN(OP_EX2); RUN(OP_RCP_LG2); RUN(OP_FFMA); RUN(OP_FADD); RUN(OP_FMUL); RUN(OP_FMNMX);
  RUN(OP_FFMA2); RUN(OP_FADD2); RUN(OP_FMUL2); RUN(OP_FFMA2_FMNMX); RUN(OP_FFMA2_MUFU); RUN(OP_MIX_S4); RUN(OP_MIX_S8); RUN(OP_MIX_S12); RUN(OP_MIX_P2);
  RUN(OP_MIX_P4); RUN(OP_MIX_P6); RUN(OP_MIX_P8); RUN(OP_BODY_SCALAR); RUN(OP_BODY_PACKED); RUN(OP_BODY_PACKED_RCP2);
#undef RUN

  printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"max_sm_mhz\": %.1f,\n", prop.name, sms, khz / 1e3);
  printf(" \"how\": \"tools/issue_peak.cu: 2 CTAs x 512 threads per SM, 8 independent register chains per thread, "
         "%d iterations, best of 5 launches, CUDA events; rates are device-wide lane-operations (or pairs) per second; "
         "per_clk_sm = rate / (SMs x max SM clock); the clock under load can sit below the max, so per_clk_sm is a lower bound\",\n",
         iters);
  printf(" \"streams\": {\n");
  for (int i = 0; i < OP_COUNT; ++i) {
    printf("  \"%s\": {\"what\": \"%s\", \"counted\": \"%s\", \"per_s\": %.6e, \"per_clk_sm\": %.3f, "
           "\"nominal_per_clk_sm\": %.3f}%s\n",
           streams[i].name, streams[i].what, streams[i].counted, rate[i], rate[i] / (sms * clk),
           streams[i].predicted_per_clk_sm, i + 1 < OP_COUNT ? "," : "");
  }
  printf(" }\n}\n");
  CK(cudaFree(dout));
  return 0;
}
