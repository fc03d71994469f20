// ltr_mlp_scorer.cuh -- the reference's documented ranker, a small ReLU MLP F -> H1 -> H2 -> 1
// (docs/source/getting-started.rst:42-51: 136 -> 50 -> 10 -> 1), scored over the flat (B*L, F) feature block
// in front of the loss kernels.
//
// This is the one contraction on the path (rows x F x H1), so layer 1 runs on the 5th-generation tensor cores:
//   * feature tiles of 128 documents arrive by TMA tensor copies (cp.async.bulk.tensor, 128-byte swizzle)
//     as K-major chunks of 32 floats; the first-layer weights sit in shared memory in the same form;
//   * one thread issues tcgen05.mma.kind::tf32 (M = 128 documents, N = 64 hidden units, K = 8 per
//     instruction), accumulating Z1 = X W1^T in tensor memory (TMEM), two accumulator tiles in flight;
//   * epilogue warps read their document's row of Z1 with tcgen05.ld (one thread per document) and run
//     bias + ReLU, layer 2 (H1 x H2) and layer 3 in registers: only the score leaves the SM.
// Precision: layer 1 multiplies TF32 operands (the tensor core reads the upper 19 bits of each float32)
// with float32 accumulation, as torch does under torch.backends.cuda.matmul.allow_tf32; layers 2-3 are
// float32 FMA.
#pragma once

#include <cuda.h>

#include "ltr_common.cuh"

namespace ltr {

constexpr int kMlpTileDocs = 128;                   // documents per tile = UMMA M
constexpr int kMlpN1 = 64;                          // hidden units per tile = UMMA N (H1 padded)
constexpr int kMlpChunkX = kMlpTileDocs * 128;      // bytes of one 32-float chunk of a feature tile
constexpr int kMlpChunkW = kMlpN1 * 128;            // bytes of one 32-float chunk of W1
constexpr int kMlpFwdThreads = 320;                 // 8 epilogue warps + TMA warp + MMA warp
constexpr int kMlpMaxH1 = 64;
constexpr int kMlpMaxH2 = 16;

// how the F feature columns are cut into K-major shared-memory chunks
struct MlpGeom {
  int F;
  int nfull;         // chunks of 32 floats (128-byte swizzle)
  int tail_pitch;    // bytes per row of the last, narrower chunk: 0 (none), 32, 64 or 128
  int tail_ksteps;   // K = 8 steps in it
  int stage_bytes;   // one feature tile
  int w1_bytes;      // W1 in the same form (64 rows)
  int stages;
};

// ---- tcgen05 / TMA-tensor wrappers ------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait (2 s of wall clock): a protocol error traps instead of hanging the device
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// shared-memory matrix descriptor (sm_100 format): start address, leading / stride byte offsets in
// 16-byte units, version 1, swizzle mode in bits 61..63
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout) << 61);
}
__host__ __device__ constexpr uint32_t umma_layout_code(int pitch) {   // bytes per swizzled row
  return pitch == 128 ? 2u : (pitch == 64 ? 4u : 6u);
}
// instruction descriptor of kind::tf32 with float32 accumulation
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(
          d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier receives one arrival once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- shared-memory plan ----------------------------------------------------------------------------------
// [W1 chunks][feature stages][small parameters][barriers]; every chunk starts on a 1024-byte boundary
struct MlpSmallParams {
  float b1[kMlpMaxH1];
  float w2t[kMlpMaxH1 * kMlpMaxH2];    // [j][i], row pitch H2P
  float b2[kMlpMaxH2];
  float w3[kMlpMaxH2];
  float b3;
  uint32_t tmem_base;
  uint64_t bar_w;
  uint64_t bar_full[4];
  uint64_t bar_empty[4];
  uint64_t bar_tfull[2];
  uint64_t bar_tempty[2];
  uint64_t bar_aux[4];
};

__device__ __forceinline__ void mlp_load_small(MlpSmallParams* sp, const float* __restrict__ b1,
                                               const float* __restrict__ w2, const float* __restrict__ b2,
                                               const float* __restrict__ w3, const float* __restrict__ b3, int H1,
                                               int H2, int H2P) {
  for (int t = threadIdx.x; t < kMlpMaxH1; t += blockDim.x) sp->b1[t] = (t < H1 && b1) ? b1[t] : 0.0f;
  for (int t = threadIdx.x; t < kMlpMaxH1 * kMlpMaxH2; t += blockDim.x) {
    const int j = t / H2P, i = t - j * H2P;
    sp->w2t[t] = (j < H1 && i < H2) ? w2[i * H1 + j] : 0.0f;
  }
  for (int t = threadIdx.x; t < kMlpMaxH2; t += blockDim.x) {
    sp->b2[t] = (t < H2 && b2) ? b2[t] : 0.0f;
    sp->w3[t] = t < H2 ? w3[t] : 0.0f;
  }
  if (threadIdx.x == 0) sp->b3 = b3 ? b3[0] : 0.0f;
}

// TMA producer: one feature tile = nfull full chunks + the tail chunk, all on `bar`
__device__ __forceinline__ void mlp_load_tile(unsigned char* stage, const CUtensorMap* map_main,
                                              const CUtensorMap* map_tail, const MlpGeom& g, int row0, int chunk_bytes,
                                              uint64_t* bar) {
  for (int c = 0; c < g.nfull; ++c) tma_load_2d(stage + c * chunk_bytes, map_main, c * 32, row0, bar);
  if (g.tail_pitch) tma_load_2d(stage + g.nfull * chunk_bytes, map_tail, g.nfull * 32, row0, bar);
}

// MMA issuer: Z1 tile (128 x 64) = X tile (128 x F) . W1^T, K-major operands
__device__ __forceinline__ void mlp_issue_layer1(uint32_t d_tmem, uint32_t xs, uint32_t ws, const MlpGeom& g) {
  constexpr uint32_t idesc = umma_idesc_tf32(kMlpTileDocs, kMlpN1, 0, 0);
  uint32_t acc = 0;
  for (int c = 0; c < g.nfull; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      umma_tf32(d_tmem, umma_desc(xs + c * kMlpChunkX + k * 32, 16, 1024, 2),
                umma_desc(ws + c * kMlpChunkW + k * 32, 16, 1024, 2), idesc, acc);
      acc = 1;
    }
  }
  if (g.tail_pitch) {
    const uint32_t code = umma_layout_code(g.tail_pitch);
    const uint32_t sbo = 8u * g.tail_pitch;
    for (int k = 0; k < g.tail_ksteps; ++k) {
      umma_tf32(d_tmem, umma_desc(xs + g.nfull * kMlpChunkX + k * 32, 16, sbo, code),
                umma_desc(ws + g.nfull * kMlpChunkW + k * 32, 16, sbo, code), idesc, acc);
      acc = 1;
    }
  }
}

// One document's hidden layers from its row of Z1 (TMEM -> registers).  H2P = H2 rounded up to 4.
template <int H1, int H2>
struct MlpRow {
  static constexpr int H1C = (H1 + 15) / 16;
  static constexpr int H2P = (H2 + 3) / 4 * 4;
  float h1[H1];      // relu(z1 + b1)
  float z2[H2P];     // pre-activation of layer 2

  __device__ __forceinline__ void load(uint32_t taddr) {
    uint32_t r[16 * H1C];
#pragma unroll
    for (int q = 0; q < H1C; ++q) tmem_ld16(taddr + 16 * q, r + 16 * q);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < H1; ++j) h1[j] = __uint_as_float(r[j]);
  }
  __device__ __forceinline__ void layers(const MlpSmallParams* sp) {
#pragma unroll
    for (int i = 0; i < H2P; ++i) z2[i] = sp->b2[i];
#pragma unroll
    for (int j = 0; j < H1; ++j) {
      h1[j] = fmaxf(h1[j] + sp->b1[j], 0.0f);
      const float4* w = reinterpret_cast<const float4*>(sp->w2t + j * H2P);
#pragma unroll
      for (int q = 0; q < H2P / 4; ++q) {
        const float4 wv = w[q];
        z2[4 * q + 0] = fmaf(wv.x, h1[j], z2[4 * q + 0]);
        z2[4 * q + 1] = fmaf(wv.y, h1[j], z2[4 * q + 1]);
        z2[4 * q + 2] = fmaf(wv.z, h1[j], z2[4 * q + 2]);
        z2[4 * q + 3] = fmaf(wv.w, h1[j], z2[4 * q + 3]);
      }
    }
  }
  __device__ __forceinline__ float score(const MlpSmallParams* sp) const {
    float s = sp->b3;
#pragma unroll
    for (int i = 0; i < H2; ++i) s = fmaf(sp->w3[i], fmaxf(z2[i], 0.0f), s);
    return s;
  }
};

// ---- forward: scores[r] = MLP(features[r, :]) -------------------------------------------------------
template <int H1, int H2>
__global__ void __launch_bounds__(kMlpFwdThreads, 1)
mlp_scores_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x_tail,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_tail,
                  const MlpGeom g, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
                  int h1, int h2, long long rows, int ntiles, float* __restrict__ scores_out) {
  extern __shared__ __align__(1024) unsigned char mlp_smem[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mlp_smem) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  unsigned char* w1s = base;
  unsigned char* xs = base + g.w1_bytes;
  MlpSmallParams* sp = reinterpret_cast<MlpSmallParams*>(xs + static_cast<size_t>(g.stages) * g.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  using Row = MlpRow<H1, H2>;

  mlp_load_small(sp, b1, w2, b2, w3, b3, h1, h2, Row::H2P);
  if (threadIdx.x == 0) {
    mbar_init(&sp->bar_w, 1);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&sp->bar_full[s], 1);
      mbar_init(&sp->bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sp->bar_tfull[b], 1);
      mbar_init(&sp->bar_tempty[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&sp->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sp->tmem_base;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&sp->bar_w, g.w1_bytes);
      mlp_load_tile(w1s, &map_w, &map_w_tail, g, 0, kMlpChunkW, &sp->bar_w);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        mbar_wait_guarded(&sp->bar_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&sp->bar_full[s], g.stage_bytes);
        mlp_load_tile(xs + static_cast<size_t>(s) * g.stage_bytes, &map_x, &map_x_tail, g, tile * kMlpTileDocs,
                      kMlpChunkX, &sp->bar_full[s]);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      mbar_wait_guarded(&sp->bar_w, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        const int b = it & 1;
        const uint32_t tph = (it >> 1) & 1;
        mbar_wait_guarded(&sp->bar_tempty[b], tph ^ 1u);
        mbar_wait_guarded(&sp->bar_full[s], ph);
        tc_fence_after();
        mlp_issue_layer1(tmem + b * kMlpN1, smem_u32(xs + static_cast<size_t>(s) * g.stage_bytes), smem_u32(w1s), g);
        umma_commit(&sp->bar_empty[s]);
        umma_commit(&sp->bar_tfull[b]);
      }
    }
  } else {
    const int grp = warp >> 2;                     // epilogue group: tiles with (it & 1) == grp
    const int quarter = warp & 3;                  // TMEM lanes 32 * quarter ..
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const uint32_t tph = (it >> 1) & 1;
      mbar_wait_guarded(&sp->bar_tfull[grp], tph);
      tc_fence_after();
      Row row;
      row.load(tmem + grp * kMlpN1 + (static_cast<uint32_t>(quarter * 32) << 16));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sp->bar_tempty[grp]);
      row.layers(sp);
      const long long r = static_cast<long long>(tile) * kMlpTileDocs + quarter * 32 + lane;
      if (r < rows) scores_out[r] = row.score(sp);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 128);
}

// ---- backward: parameter gradients for an upstream d loss / d scores --------------------------------------
// Second pass over the features (the loss needs every score of a query before any gradient exists, and a
// query's features do not stay on chip between the two).  Per tile of 128 documents:
//   MMA1   Z1 = X W1^T again (recomputing beats storing H1: 200 B per document against 544 B of features);
//   warps  one thread per document: h1, z2, h2 as in the forward pass, then
//            dz2 = ds * w3 * [z2 > 0],  dh1 = W2^T dz2,  dz1 = dh1 * [z1 > 0]
//          dW3 / db3 / db2 / db1 accumulate in registers over all the tiles of the CTA; dW2 = dZ2^T H1 goes
//          through a shared-memory transpose, two half tiles at a time, every thread owning a 2 x NI block of it;
//          dZ1^T is written to shared memory as the K-major A operand (128-byte swizzle) of
//   MMA2   dW1 (64 x F) += dZ1^T (64 x 128 documents) . X (128 documents x F): M = 64, X is the MN-major B
//          operand; the accumulator stays in TMEM for the whole launch.
// kind::tf32 takes an MN-major operand only in the "128-byte swizzle, 32-byte atom" layout and a K-major one
// only in the 16-byte-atom layouts (tools/umma_probe.py: every other combination returns zeros or faults),
// so the tile is fetched twice, once per layout; the second TMA copy is served by L2.  One buffer per layout
// is enough: the K-major copy is dead as soon as MMA1 has run (the next tile streams in during the
// epilogue), the MN-major copy is only needed by MMA2 at the end of the epilogue.
// Each CTA writes one partial gradient vector [dW1 | db1 | dW2 | db2 | dW3 | db3]; mlp_reduce_kernel sums
// them in CTA order (bit-reproducible).
constexpr int kMlpBwdThreads = 192;                 // 4 epilogue warps + TMA warp + MMA warp
constexpr int kMlpExPitch = 68;                     // floats per row of the transposed exchange arrays
constexpr int kMlpA2Bytes = kMlpN1 * kMlpTileDocs * 4;
constexpr int kMlpD2Col = 2 * kMlpN1;               // first TMEM column of the dW1 accumulator
constexpr uint32_t kUmmaLayout32BAtom = 1;          // SWIZZLE_128B_BASE32B

struct MlpBwdSmem {
  MlpSmallParams sp;
  float red[4][kMlpMaxH1 + 3 * kMlpMaxH2 + 4];      // per-warp sums of db1 | db2 | dW3 | db3
};
static_assert((kMlpMaxH1 + 4 + kMlpMaxH2 + 4) * kMlpExPitch * 4 <= kMlpA2Bytes, "exchange arrays live in the A2 buffer");

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int H1, int H2>
__global__ void __launch_bounds__(kMlpBwdThreads, 1)
mlp_backward_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x_tail,
                    const __grid_constant__ CUtensorMap map_x_mn, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_w_tail, const MlpGeom g, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                    const float* __restrict__ b3, int h1n, int h2n, const float* __restrict__ dscores, long long rows,
                    int ntiles, int tmem_cols, float* __restrict__ partials, int partial_len) {
  extern __shared__ __align__(1024) unsigned char mlp_smem[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mlp_smem) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  const int mn_chunks = (g.F + 31) / 32;            // MN-major copy: full 32-feature boxes, zero filled
  unsigned char* w1s = base;
  unsigned char* a2s = base + g.w1_bytes;
  unsigned char* xk = a2s + kMlpA2Bytes;            // K-major copy (MMA1)
  unsigned char* xmn = xk + g.stage_bytes;          // MN-major copy (MMA2)
  MlpBwdSmem* sm = reinterpret_cast<MlpBwdSmem*>(xmn + mn_chunks * kMlpChunkX);
  MlpSmallParams* sp = &sm->sp;
  float* ex_h1 = reinterpret_cast<float*>(a2s);     // [j][doc of the half tile], between two uses of A2
  float* ex_dz2 = ex_h1 + (kMlpMaxH1 + 4) * kMlpExPitch;
  uint64_t* bar_a2_full = &sp->bar_aux[0];
  uint64_t* bar_a2_free = &sp->bar_aux[1];
  uint64_t* bar_mn_full = &sp->bar_aux[2];
  uint64_t* bar_mn_empty = &sp->bar_aux[3];
  uint64_t* bar_k_full = &sp->bar_full[0];
  uint64_t* bar_k_empty = &sp->bar_empty[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  using Row = MlpRow<H1, H2>;
  constexpr int H2P = Row::H2P;
  constexpr int JP = (H1 + 1) / 2;                  // thread t of the dW2 phase owns rows jp and jp + JP of H1 ...
  constexpr int IG = 128 / JP;                      // ... and rows ig, ig + IG, ... of dZ2
  constexpr int NI = (H2 + IG - 1) / IG;
  static_assert(2 * JP <= kMlpMaxH1 + 4 && NI * IG <= kMlpMaxH2 + 4, "exchange rows");
  const int d2_n = mn_chunks * 32;                  // accumulator columns: whole 32-feature atoms

  mlp_load_small(sp, b1, w2, b2, w3, b3, h1n, h2n, H2P);
  for (int t = threadIdx.x; t < kMlpA2Bytes / 4; t += blockDim.x) reinterpret_cast<float*>(a2s)[t] = 0.0f;
  if (threadIdx.x == 0) {
    mbar_init(&sp->bar_w, 1);
    mbar_init(bar_k_full, 1);
    mbar_init(bar_k_empty, 1);
    mbar_init(bar_mn_full, 1);
    mbar_init(bar_mn_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sp->bar_tfull[b], 1);
      mbar_init(&sp->bar_tempty[b], 4);
    }
    mbar_init(bar_a2_full, 4);
    mbar_init(bar_a2_free, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&sp->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sp->tmem_base;
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 4) {
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(&sp->bar_w, g.w1_bytes);
      mlp_load_tile(w1s, &map_w, &map_w_tail, g, 0, kMlpChunkW, &sp->bar_w);
      auto load_k = [&](int it) {                     // needs MMA1 of tile it - 1 done
        mbar_wait_guarded(bar_k_empty, (it & 1) ^ 1u);
        mbar_arrive_expect_tx(bar_k_full, g.stage_bytes);
        mlp_load_tile(xk, &map_x, &map_x_tail, g, (blockIdx.x + it * gridDim.x) * kMlpTileDocs, kMlpChunkX,
                      bar_k_full);
      };
      load_k(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) load_k(it + 1);
        mbar_wait_guarded(bar_mn_empty, (it & 1) ^ 1u);   // MMA2 of tile it - 1 done
        mbar_arrive_expect_tx(bar_mn_full, mn_chunks * kMlpChunkX);
        for (int c = 0; c < mn_chunks; ++c)
          tma_load_2d(xmn + c * kMlpChunkX, &map_x_mn, c * 32, (blockIdx.x + it * gridDim.x) * kMlpTileDocs,
                      bar_mn_full);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc2 = umma_idesc_tf32(64, d2_n, 0, 1);
      const uint32_t a2 = smem_u32(a2s);
      const uint32_t x_k = smem_u32(xk), x_mn = smem_u32(xmn);
      mbar_wait_guarded(&sp->bar_w, 0);
      // layer 1 of tile `it` into accumulator it & 1; the K-major copy is released at once
      auto issue_mma1 = [&](int it) {
        const int b = it & 1;
        mbar_wait_guarded(&sp->bar_tempty[b], ((it >> 1) & 1) ^ 1u);
        mbar_wait_guarded(bar_k_full, it & 1);
        tc_fence_after();
        mlp_issue_layer1(tmem + b * kMlpN1, x_k, smem_u32(w1s), g);
        umma_commit(bar_k_empty);
        umma_commit(&sp->bar_tfull[b]);
      };
      issue_mma1(0);
      uint32_t acc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) issue_mma1(it + 1);    // one tile ahead: the epilogue warps never wait for layer 1
        mbar_wait_guarded(bar_mn_full, it & 1);
        mbar_wait_guarded(bar_a2_full, it & 1);
        tc_fence_after();
        for (int ks = 0; ks < kMlpTileDocs / 8; ++ks) {
          umma_tf32(tmem + kMlpD2Col, umma_desc(a2 + (ks >> 2) * (kMlpN1 * 128) + (ks & 3) * 32, 16, 1024, 2),
                    umma_desc(x_mn + ks * 1024, kMlpChunkX, 512, kUmmaLayout32BAtom), idesc2, acc);
          acc = 1;
        }
        umma_commit(bar_a2_free);
        umma_commit(bar_mn_empty);
      }
    }
  } else {
    const int t = threadIdx.x;                        // 0..127 = document of the tile = TMEM lane
    const int jp = t % JP, ig = t / JP;
    float db1acc[H1], db2acc[H2], dw3acc[H2], db3acc = 0.0f;
    float dw2acc[2][NI];
#pragma unroll
    for (int j = 0; j < H1; ++j) db1acc[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < H2; ++i) db2acc[i] = dw3acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < NI; ++k) dw2acc[0][k] = dw2acc[1][k] = 0.0f;

    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const long long r = static_cast<long long>(tile) * kMlpTileDocs + t;
      const float ds = r < rows ? dscores[r] : 0.0f;
      const int b = it & 1;
      mbar_wait_guarded(&sp->bar_tfull[b], (it >> 1) & 1);
      tc_fence_after();
      Row row;
      row.load(tmem + b * kMlpN1 + (static_cast<uint32_t>(warp * 32) << 16));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sp->bar_tempty[b]);
      row.layers(sp);
      float dz2[H2P];
      db3acc += ds;
#pragma unroll
      for (int i = 0; i < H2P; ++i) {
        if (i < H2) {
          dw3acc[i] = fmaf(ds, fmaxf(row.z2[i], 0.0f), dw3acc[i]);
          dz2[i] = row.z2[i] > 0.0f ? ds * sp->w3[i] : 0.0f;
          db2acc[i] += dz2[i];
        } else {
          dz2[i] = 0.0f;
        }
      }
      // the A2 buffer is free once MMA2 of the previous tile has read it
      mbar_wait_guarded(bar_a2_free, (it & 1) ^ 1u);
      // dW2 += dZ2^T H1, half a tile at a time through the transposed exchange arrays (inside A2)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        epi_bar();                                    // the previous readers are done
        if ((warp >> 1) == half) {
          const int d = t & 63;
#pragma unroll
          for (int j = 0; j < H1; ++j) ex_h1[j * kMlpExPitch + d] = row.h1[j];
#pragma unroll
          for (int i = 0; i < H2; ++i) ex_dz2[i * kMlpExPitch + d] = dz2[i];
        }
        epi_bar();
        if (ig < IG) {
          const float4* ha = reinterpret_cast<const float4*>(ex_h1 + jp * kMlpExPitch);
          const float4* hb = reinterpret_cast<const float4*>(ex_h1 + (jp + JP) * kMlpExPitch);
#pragma unroll 4
          for (int d4 = 0; d4 < 16; ++d4) {
            const float4 a = ha[d4], c = hb[d4];
#pragma unroll
            for (int k = 0; k < NI; ++k) {
              const float4 z = reinterpret_cast<const float4*>(ex_dz2 + (ig + k * IG) * kMlpExPitch)[d4];
              dw2acc[0][k] = fmaf(z.x, a.x, fmaf(z.y, a.y, fmaf(z.z, a.z, fmaf(z.w, a.w, dw2acc[0][k]))));
              dw2acc[1][k] = fmaf(z.x, c.x, fmaf(z.y, c.y, fmaf(z.z, c.z, fmaf(z.w, c.w, dw2acc[1][k]))));
            }
          }
        }
      }
      epi_bar();                                      // exchange reads done: A2 becomes the operand again
      // dZ1^T -> A operand of MMA2 (K-major, 128-byte swizzle: row j, 16-byte chunk (doc / 4) ^ (j % 8));
      // rows H1..63 are rewritten with zeros (the exchange arrays passed through them)
      {
        unsigned char* a2row = a2s + (t >> 5) * (kMlpN1 * 128) + (t & 3) * 4;
        const int chunk = (t & 31) >> 2;
#pragma unroll
        for (int j = 0; j < kMlpN1; ++j) {
          float dz1 = 0.0f;
          if (j < H1) {
            const float4* w = reinterpret_cast<const float4*>(sp->w2t + j * H2P);
            float dh = 0.0f;
#pragma unroll
            for (int q = 0; q < H2P / 4; ++q) {
              const float4 wv = w[q];
              dh = fmaf(wv.x, dz2[4 * q + 0], dh);
              dh = fmaf(wv.y, dz2[4 * q + 1], dh);
              dh = fmaf(wv.z, dz2[4 * q + 2], dh);
              dh = fmaf(wv.w, dz2[4 * q + 3], dh);
            }
            dz1 = row.h1[j < H1 ? j : 0] > 0.0f ? dh : 0.0f;
            db1acc[j < H1 ? j : 0] += dz1;
          }
          *reinterpret_cast<float*>(a2row + j * 128 + ((chunk ^ (j & 7)) << 4)) = dz1;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a2_full);
    }

    // ---- per-CTA partial gradient vector ----
    float* out = partials + static_cast<size_t>(blockIdx.x) * partial_len;
    const int off_b1 = h1n * g.F, off_w2 = off_b1 + h1n, off_b2 = off_w2 + h2n * h1n, off_w3 = off_b2 + h2n,
              off_b3 = off_w3 + h2n;
    // dW2: every entry has one owner
    if (ig < IG) {
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        const int i = ig + k * IG;
        if (i < h2n) {
          if (jp < h1n) out[off_w2 + i * h1n + jp] = dw2acc[0][k];
          if (jp + JP < h1n) out[off_w2 + i * h1n + jp + JP] = dw2acc[1][k];
        }
      }
    }
    // db1 | db2 | dW3 | db3: warp sums, then the four warps in order
    float* red = sm->red[warp];
#pragma unroll
    for (int j = 0; j < H1; ++j) {
      const float v = warp_sum(db1acc[j]);
      if (lane == 0) red[j] = v;
    }
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      const float v = warp_sum(db2acc[i]), u = warp_sum(dw3acc[i]);
      if (lane == 0) {
        red[kMlpMaxH1 + i] = v;
        red[kMlpMaxH1 + kMlpMaxH2 + i] = u;
      }
    }
    {
      const float v = warp_sum(db3acc);
      if (lane == 0) red[kMlpMaxH1 + 2 * kMlpMaxH2] = v;
    }
    epi_bar();
    auto red4 = [&](int k) { return ((sm->red[0][k] + sm->red[1][k]) + sm->red[2][k]) + sm->red[3][k]; };
    if (t < h1n) out[off_b1 + t] = red4(t);
    if (t < h2n) {
      out[off_b2 + t] = red4(kMlpMaxH1 + t);
      out[off_w3 + t] = red4(kMlpMaxH1 + kMlpMaxH2 + t);
    }
    if (t == 0) out[off_b3] = red4(kMlpMaxH1 + 2 * kMlpMaxH2);
    // dW1 out of TMEM: row j of the M = 64 accumulator lives in TMEM lane (j % 16) + 32 (j / 16)
    if (my_tiles > 0) {
      mbar_wait_guarded(bar_a2_free, (my_tiles - 1) & 1);
      tc_fence_after();
      const int j = warp * 16 + lane;
      for (int c0 = 0; c0 < d2_n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + kMlpD2Col + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_ld_wait();
        if (lane < 16 && j < h1n) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (c0 + k < g.F) out[static_cast<size_t>(j) * g.F + c0 + k] = __uint_as_float(v[k]);
        }
      }
    } else {
      for (int k = t; k < h1n * g.F; k += 128) out[k] = 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, tmem_cols);
}

// out[k] = sum over the CTAs' partial vectors, in CTA order
__global__ void __launch_bounds__(256)
mlp_reduce_kernel(const float* __restrict__ partials, int nparts, int len, float* __restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int p = 0; p < nparts; ++p) s += partials[static_cast<size_t>(p) * len + k];
    out[k] = s;
  }
}

}  // namespace ltr
