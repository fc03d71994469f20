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
#include "ltr_p2p.cuh"

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
  int slab_chunks;   // 0: W1 resident, one stage = one feature tile.  > 0 (wide rows, forward only): W1 does not
                     // fit beside a tile, so a stage holds `slab_chunks` chunks of the tile and the matching
                     // chunks of W1, and the tile's product accumulates over ceil(nchunks / slab_chunks) stages
  int nchunks;       // ceil(F / 32)
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
template <bool BACKOFF = false>
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
    if (BACKOFF) __nanosleep(64);                   // single-thread roles: leave the issue slots to the math warps
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
  float w2t[kMlpMaxH1 * kMlpMaxH2];    // [j][i], row pitch H2P (forward epilogue only)
  float b2[kMlpMaxH2];
  float w3[kMlpMaxH2];
  float b3;
  uint32_t tmem_base;
  uint64_t bar_w;
  uint64_t bar_full[8];
  uint64_t bar_empty[8];
  uint64_t bar_tfull[2];
  uint64_t bar_tempty[2];
  uint64_t bar_aux[8];
};

__device__ __forceinline__ void mlp_load_small(MlpSmallParams* sp, const float* __restrict__ b1,
                                               const float* __restrict__ w2, const float* __restrict__ b2,
                                               const float* __restrict__ w3, const float* __restrict__ b3, int H1,
                                               int H2, int H2P) {
  for (int t = threadIdx.x; t < kMlpMaxH1; t += blockDim.x) sp->b1[t] = (t < H1 && b1) ? b1[t] : 0.0f;
  if (H2P > 0) {
    for (int t = threadIdx.x; t < kMlpMaxH1 * kMlpMaxH2; t += blockDim.x) {
      const int j = t / H2P, i = t - j * H2P;
      sp->w2t[t] = (j < H1 && i < H2) ? w2[i * H1 + j] : 0.0f;
    }
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

// Activation row kept for the backward pass (ltr_mlp_scores with hz_out): P floats per document,
// H1 = relu(Z1 + b1) in columns [0, H1), the layer-2 pre-activation Z2 (bias included) in [Z0, Z0 + H2), 1.0 in
// column Z0 + H2 (the sums of dZ1 / dZ2 over documents then fall out of the same product as dW2), zeros elsewhere.  The backward kernel uses the row index of the dZ1 operand for both: rows [0, H1) carry dZ1,
// rows [Z0, Z0 + H2) carry dZ2 -- which needs Z0 + H2 <= 64 (true for 50-10 and 32-8).
template <int H1, int H2>
struct MlpHz {
  static constexpr int Z0 = (H1 + 3) / 4 * 4;
  static constexpr int ONE = Z0 + H2;               // a column of ones: its products with dZ1 / dZ2 are db1 / db2
  static constexpr int P = (ONE + 1 + 3) / 4 * 4;
  static constexpr bool kFits = Z0 + H2 <= kMlpN1;
};

// Forward epilogue: one document's hidden layers from its row of Z1 (TMEM -> registers); H2P = H2 rounded up
// to 4.  Layer 2 as packed f32x2 FMAs on pairs of units; the rows of W2^T (shared memory, one 128-bit
// broadcast load per four weights) are fetched PF hidden units ahead of their use.
template <int H1, int H2>
struct MlpRow {
  static constexpr int H1C = (H1 + 15) / 16;
  static constexpr int H2P = (H2 + 3) / 4 * 4;
  static constexpr int PF = H1 < 4 ? H1 : 4;
  float h1[H1];
  float2 z2[H2P / 2];

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
    for (int q = 0; q < H2P / 2; ++q) z2[q] = make_float2(sp->b2[2 * q], sp->b2[2 * q + 1]);
    float4 wbuf[PF][H2P / 4];
    float bbuf[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      bbuf[j] = sp->b1[j];
#pragma unroll
      for (int q = 0; q < H2P / 4; ++q) wbuf[j][q] = reinterpret_cast<const float4*>(sp->w2t + j * H2P)[q];
    }
#pragma unroll
    for (int j = 0; j < H1; ++j) {
      h1[j] = fmaxf(h1[j] + bbuf[j % PF], 0.0f);
      const float2 hh = make_float2(h1[j], h1[j]);
#pragma unroll
      for (int q = 0; q < H2P / 4; ++q) {
        const float4 wv = wbuf[j % PF][q];
        z2[2 * q + 0] = __ffma2_rn(make_float2(wv.x, wv.y), hh, z2[2 * q + 0]);
        z2[2 * q + 1] = __ffma2_rn(make_float2(wv.z, wv.w), hh, z2[2 * q + 1]);
      }
      if (j + PF < H1) {
        bbuf[j % PF] = sp->b1[j + PF];
#pragma unroll
        for (int q = 0; q < H2P / 4; ++q) wbuf[j % PF][q] = reinterpret_cast<const float4*>(sp->w2t + (j + PF) * H2P)[q];
      }
    }
  }
  __device__ __forceinline__ float score(const MlpSmallParams* sp) const {
    float s = sp->b3;
#pragma unroll
    for (int i = 0; i < H2; ++i) s = fmaf(sp->w3[i], fmaxf((i & 1) ? z2[i >> 1].y : z2[i >> 1].x, 0.0f), s);
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
                  int h1, int h2, long long rows, int ntiles, float* __restrict__ scores_out,
                  float* __restrict__ hz_out) {
  extern __shared__ __align__(1024) unsigned char mlp_smem[];
  // aligned up to 1024 bytes in the pointer domain, so the compiler keeps the shared address space (LDS / STS)
  unsigned char* base = mlp_smem + ((1024u - (smem_u32(mlp_smem) & 1023u)) & 1023u);
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
    if (lane == 0 && g.slab_chunks > 0) {
      // wide rows: (tile, slab) pairs through the stage ring; columns past F and hidden units past H1 arrive as
      // zeros (out-of-range elements of a tensor copy), so every chunk is a full 32-column box
      int v = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c0 = 0; c0 < g.nchunks; c0 += g.slab_chunks, ++v) {
          const int s = v % g.stages;
          const int nch = g.nchunks - c0 < g.slab_chunks ? g.nchunks - c0 : g.slab_chunks;
          unsigned char* stage = xs + static_cast<size_t>(s) * g.stage_bytes;
          mbar_wait_guarded<true>(&sp->bar_empty[s], ((v / g.stages) & 1) ^ 1u);
          mbar_arrive_expect_tx(&sp->bar_full[s], nch * (kMlpChunkX + kMlpChunkW));
          for (int c = 0; c < nch; ++c) {
            tma_load_2d(stage + c * kMlpChunkX, &map_x, (c0 + c) * 32, tile * kMlpTileDocs, &sp->bar_full[s]);
            tma_load_2d(stage + g.slab_chunks * kMlpChunkX + c * kMlpChunkW, &map_w, (c0 + c) * 32, 0,
                        &sp->bar_full[s]);
          }
        }
      }
    } else if (lane == 0) {
      mbar_arrive_expect_tx(&sp->bar_w, g.w1_bytes);
      mlp_load_tile(w1s, &map_w, &map_w_tail, g, 0, kMlpChunkW, &sp->bar_w);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        mbar_wait_guarded<true>(&sp->bar_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&sp->bar_full[s], g.stage_bytes);
        mlp_load_tile(xs + static_cast<size_t>(s) * g.stage_bytes, &map_x, &map_x_tail, g, tile * kMlpTileDocs,
                      kMlpChunkX, &sp->bar_full[s]);
      }
    }
  } else if (warp == 9) {
    if (lane == 0 && g.slab_chunks > 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kMlpTileDocs, kMlpN1, 0, 0);
      int it = 0, v = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait_guarded<true>(&sp->bar_tempty[b], ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        uint32_t acc = 0;
        for (int c0 = 0; c0 < g.nchunks; c0 += g.slab_chunks, ++v) {
          const int s = v % g.stages;
          const int nch = g.nchunks - c0 < g.slab_chunks ? g.nchunks - c0 : g.slab_chunks;
          const uint32_t xa = smem_u32(xs + static_cast<size_t>(s) * g.stage_bytes);
          const uint32_t wa = xa + g.slab_chunks * kMlpChunkX;
          mbar_wait_guarded<true>(&sp->bar_full[s], (v / g.stages) & 1);
          tc_fence_after();
          for (int c = 0; c < nch; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_tf32(tmem + b * kMlpN1, umma_desc(xa + c * kMlpChunkX + k * 32, 16, 1024, 2),
                        umma_desc(wa + c * kMlpChunkW + k * 32, 16, 1024, 2), idesc, acc);
              acc = 1;
            }
          }
          umma_commit(&sp->bar_empty[s]);
        }
        umma_commit(&sp->bar_tfull[b]);
      }
    } else if (lane == 0) {
      mbar_wait_guarded<true>(&sp->bar_w, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        const int b = it & 1;
        const uint32_t tph = (it >> 1) & 1;
        mbar_wait_guarded<true>(&sp->bar_tempty[b], tph ^ 1u);
        mbar_wait_guarded<true>(&sp->bar_full[s], ph);
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
      if (hz_out && r < rows) {
        // the activations the backward pass needs, one row per document: [H1 | 0 | Z2 | 0] (MlpHz)
        using Hz = MlpHz<H1, H2>;
        float4* o = reinterpret_cast<float4*>(hz_out + static_cast<size_t>(r) * Hz::P);
#pragma unroll
        for (int q = 0; q < Hz::P / 4; ++q) {
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = 4 * q + k;
            v[k] = j < H1 ? row.h1[j < H1 ? j : 0]
                          : (j >= Hz::Z0 && j < Hz::Z0 + H2
                                 ? (((j - Hz::Z0) & 1) ? row.z2[(j >= Hz::Z0 && j < Hz::Z0 + H2 ? j - Hz::Z0 : 0) >> 1].y
                                                       : row.z2[(j >= Hz::Z0 && j < Hz::Z0 + H2 ? j - Hz::Z0 : 0) >> 1].x)
                                 : (j == Hz::ONE ? 1.0f : 0.0f));
          }
          o[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 128);
}

// ---- backward: parameter gradients for an upstream d loss / d scores --------------------------------------
// Second pass over the features (the loss needs every score of a query before any gradient exists, and a
// query's features do not stay on chip between the two).  Every contraction of a tile of 128 documents runs
// on the tensor cores; the warps only apply the element-wise steps between them (one thread per document):
//   MMA1    Z1 = X W1^T again (recomputing beats storing H1: 200 B per document against 544 B of features)
//   warps   H1 = relu(Z1 + b1), written back over Z1 in tensor memory
//   MMA-L2  Z2 (128 x 16) = H1 W2^T        A operand straight from TMEM, B = W2 in shared memory (K-major)
//   warps   dZ2 = ds * w3 * [Z2 + b2 > 0] into TMEM; dW3 / db3 / db2 in registers
//   MMA-dH  dH1 (128 x 64) = dZ2 W2        A from TMEM, B = W2^T in shared memory (K-major, 64-byte swizzle)
//   warps   dZ1 = dH1 * [H1 > 0], db1 in registers, dZ1^T to shared memory as the K-major A operand of
//   MMA2    dW1 (64 x F) += dZ1^T (64 x 128 documents) . X (128 documents x F): M = 64, X is the MN-major B
//           operand; the accumulator stays in TMEM for the whole launch.
// dW2 = dZ2^T H1 (10 x 50, contraction over documents) has no operand layout left in shared memory: H1 and
// dZ2 pass through a document-major exchange array inside the dZ1 operand buffer, half a tile at a time,
// and eight warps sum BJ x 2 blocks of it with packed FMAs while MMA-dH runs.
// kind::tf32 takes an MN-major operand only in the "128-byte swizzle, 32-byte atom" layout and a K-major one
// only in the 16-byte-atom layouts (tools/umma_probe.py: every other combination returns zeros or faults),
// so the tile is fetched twice, once per layout; the second TMA copy is served by L2.  One buffer per layout
// is enough: the K-major copy is dead as soon as MMA1 has run (the next tile streams in during the
// epilogue), the MN-major copy is only needed by MMA2 at the end of the epilogue.
// Each CTA writes one partial gradient vector [dW1 | db1 | dW2 | db2 | dW3 | db3]; mlp_reduce_kernel sums
// them in CTA order (bit-reproducible).
constexpr int kMlpBwdThreads = 320;                 // 4 document warps + 4 dW2 helper warps + TMA warp + MMA warp
constexpr int kMlpBwdEpi = 256;
constexpr int kMlpHPitch = 68;                      // floats per document row of the H1 exchange array
constexpr int kMlpZPitch = 20;                      // ... of the dZ2 exchange array
constexpr int kMlpA2Bytes = kMlpN1 * kMlpTileDocs * 4;
constexpr int kMlpW2Bytes = 2 * kMlpMaxH2 * 128 + kMlpN1 * 64;   // W2 (2 chunks of 16 rows x 128 B) + W2^T (64 rows x 64 B)
constexpr int kMlpBufCols = 160;                    // TMEM columns of one tile: Z1/H1 64 | Z2 16 | dZ2 16 | dH1 64
constexpr int kMlpD2Col = 2 * kMlpBufCols;          // first TMEM column of the dW1 accumulator
constexpr uint32_t kUmmaLayout32BAtom = 1;          // SWIZZLE_128B_BASE32B
constexpr int kMlpRedLen = kMlpMaxH1 + 2 * kMlpMaxH2 + 4;

struct MlpBwdSmall {
  float b1[kMlpMaxH1];
  float b2[kMlpMaxH2];
  float w3[kMlpMaxH2];
  uint32_t tmem_base;
  uint32_t pad;
  uint64_t bar_w;
  uint64_t bar_kf[8];      // K-major feature chunk c of the next tile has landed / has been consumed by MMA1
  uint64_t bar_ke[8];
  uint64_t bar_mnf[4];     // MN-major copy, 32 documents at a time: landed / consumed by MMA2
  uint64_t bar_mne[4];
  uint64_t bar_tfull[2];
  uint64_t bar_tempty[2];
  uint64_t bar_aux[8];
};
struct MlpBwdSmem {
  MlpBwdSmall sp;
  float red[4][kMlpRedLen];                         // per-warp sums of db2 | dW3 | db3
};
constexpr int kMlpOnesBytes = 1024;                 // 8 rows x 128 B of 1.0f: the B operand that sums dZ1 over documents
static_assert(64 * (kMlpHPitch + kMlpZPitch) * 4 <= kMlpA2Bytes, "exchange arrays live in the A2 buffer");

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (M rows = TMEM lanes, K = 8 consecutive columns) from TMEM
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(
          d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// N consecutive floats to shared memory with the widest stores the (compile-time) offset allows
template <int N>
__device__ __forceinline__ void sts_run(float* dst, const float* v) {
#pragma unroll
  for (int k = 0; k + 4 <= N; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
#pragma unroll
  for (int k = N / 4 * 4; k < N; ++k) dst[k] = v[k];
}

template <int H1, int H2>
__global__ void __launch_bounds__(kMlpBwdThreads, 1)
mlp_backward_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x_tail,
                    const __grid_constant__ CUtensorMap map_x_mn, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_w_tail, const MlpGeom g, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                    const float* __restrict__ b3, int h1n, int h2n, const float* __restrict__ dscores, long long rows,
                    int ntiles, int tmem_cols, float* __restrict__ partials, int partial_len,
                    long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) unsigned char mlp_smem[];
  // aligned up to 1024 bytes in the pointer domain, so the compiler keeps the shared address space (LDS / STS)
  unsigned char* base = mlp_smem + ((1024u - (smem_u32(mlp_smem) & 1023u)) & 1023u);
  // optional event trace of CTA 0 (tools/mlp_trace.py): dbg[tile * 32 + slot] = clock
  auto stamp = [&](int it, int slot) {
    if (dbg && blockIdx.x == 0 && it < 24) dbg[it * 32 + slot] = clock64();
  };
  const int mn_chunks = (g.F + 31) / 32;            // MN-major copy: full 32-feature boxes, zero filled
  unsigned char* w1s = base;
  unsigned char* a2s = base + g.w1_bytes;
  unsigned char* xk = a2s + kMlpA2Bytes;            // K-major copy (MMA1)
  unsigned char* xmn = xk + g.stage_bytes;          // MN-major copy (MMA2)
  unsigned char* w2s = xmn + mn_chunks * kMlpChunkX;   // W2 [i][j] as two K-major chunks, 128-byte swizzle
  unsigned char* w2ts = w2s + 2 * kMlpMaxH2 * 128;     // W2^T [j][i], K-major, 64-byte swizzle
  unsigned char* ones = w2s + kMlpW2Bytes;
  MlpBwdSmem* sm = reinterpret_cast<MlpBwdSmem*>(ones + kMlpOnesBytes);
  MlpBwdSmall* sp = &sm->sp;
  uint64_t* bar_a2_full = &sp->bar_aux[0];
  uint64_t* bar_a2_free = &sp->bar_aux[1];
  uint64_t* bar_h1 = &sp->bar_aux[4];               // H1 of the tile is in TMEM
  uint64_t* bar_z2 = &sp->bar_aux[5];               // MMA-L2 done
  uint64_t* bar_dz2 = &sp->bar_aux[6];              // dZ2 of the tile is in TMEM
  uint64_t* bar_dh = &sp->bar_aux[7];               // MMA-dH done
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d2_n = mn_chunks * 32;                  // accumulator columns: whole 32-feature atoms
  const int nk = g.nfull + (g.tail_pitch ? 1 : 0);  // K-major chunks of a tile
  // dW2 phase: lane b of each of the eight warps owns the BJ x 2 block (j block b / IB, unit pair b % IB) of
  // dW2 and sums it over the documents its warp is dealt (8 per half tile)
  constexpr int IB = (H2 + 1) / 2;
  constexpr int BJ = ((H1 + 32 / IB - 1) / (32 / IB) + 3) / 4 * 4;
  constexpr int NJB = (H1 + BJ - 1) / BJ;
  static_assert(NJB * IB <= 32 && NJB * BJ <= kMlpHPitch && 2 * IB <= kMlpZPitch && H2 <= kMlpMaxH2 && H1 <= kMlpMaxH1,
                "dW2 blocking");
  static_assert(8 * 32 * BJ * 8 <= kMlpA2Bytes, "final dW2 partials live in the A2 buffer");

  for (int t = threadIdx.x; t < kMlpMaxH1; t += blockDim.x) sp->b1[t] = (t < h1n && b1) ? b1[t] : 0.0f;
  for (int t = threadIdx.x; t < kMlpMaxH2; t += blockDim.x) {
    sp->b2[t] = (t < h2n && b2) ? b2[t] : 0.0f;
    sp->w3[t] = t < h2n ? w3[t] : 0.0f;
  }
  for (int t = threadIdx.x; t < kMlpA2Bytes / 4; t += blockDim.x) reinterpret_cast<float*>(a2s)[t] = 0.0f;
  for (int t = threadIdx.x; t < kMlpOnesBytes / 4; t += blockDim.x) reinterpret_cast<float*>(ones)[t] = 1.0f;
  // B operands of the two small MMAs, written in the swizzled K-major layouts the tensor core reads
  for (int t = threadIdx.x; t < kMlpMaxH2 * kMlpMaxH1; t += blockDim.x) {
    const int i = t / kMlpMaxH1, j = t - i * kMlpMaxH1;
    const float v = (i < h2n && j < h1n) ? w2[i * h1n + j] : 0.0f;
    // W2: row i, K = j: chunk j / 32 of [16 rows][128 B], 16-byte unit ((j % 32) / 4) ^ (i % 8)
    *reinterpret_cast<float*>(w2s + (j >> 5) * (kMlpMaxH2 * 128) + i * 128 + (((((j & 31) >> 2) ^ (i & 7))) << 4) +
                              (j & 3) * 4) = v;
    // W2^T: row j, K = i: [64 rows][64 B], 16-byte unit (i / 4) ^ ((j / 2) % 4)
    *reinterpret_cast<float*>(w2ts + j * 64 + ((((i >> 2) ^ ((j >> 1) & 3))) << 4) + (i & 3) * 4) = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&sp->bar_w, 1);
    for (int c = 0; c < 8; ++c) {
      mbar_init(&sp->bar_kf[c], 1);
      mbar_init(&sp->bar_ke[c], 1);
    }
    for (int c = 0; c < 4; ++c) {
      mbar_init(&sp->bar_mnf[c], 1);
      mbar_init(&sp->bar_mne[c], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sp->bar_tfull[b], 1);
      mbar_init(&sp->bar_tempty[b], 4);
    }
    mbar_init(bar_a2_full, 4);
    mbar_init(bar_a2_free, 1);
    mbar_init(bar_h1, 4);
    mbar_init(bar_z2, 1);
    mbar_init(bar_dz2, 4);
    mbar_init(bar_dh, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&sp->tmem_base, tmem_cols);
  fence_proxy_async();                               // W2 / W2^T / zeroed A2 are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sp->tmem_base;
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 8) {
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(&sp->bar_w, g.w1_bytes);
      mlp_load_tile(w1s, &map_w, &map_w_tail, g, 0, kMlpChunkW, &sp->bar_w);
      // Both copies stream in pieces that are refilled as soon as the tensor core has consumed them (the
      // K-major copy chunk by chunk of 32 features, the MN-major copy 32 documents at a time): with one
      // buffer per layout the loads of the next tile still overlap the MMAs of this one.
      auto load_k = [&](int it) {
        const int row0 = (blockIdx.x + it * gridDim.x) * kMlpTileDocs;
        for (int c = 0; c < nk; ++c) {
          mbar_wait_guarded<true>(&sp->bar_ke[c], (it & 1) ^ 1u);   // MMA1 of tile it - 1 has read chunk c
          const bool tail = c >= g.nfull;
          mbar_arrive_expect_tx(&sp->bar_kf[c], tail ? kMlpTileDocs * g.tail_pitch : kMlpChunkX);
          tma_load_2d(xk + c * kMlpChunkX, tail ? &map_x_tail : &map_x, c * 32, row0, &sp->bar_kf[c]);
          if (c == 0) stamp(it, 0);                   // first K-major chunk of tile `it` requested
        }
        stamp(it, 1);                                 // last one requested
      };
      load_k(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) load_k(it + 1);
        const int row0 = (blockIdx.x + it * gridDim.x) * kMlpTileDocs;
        for (int dg = 0; dg < 4; ++dg) {
          mbar_wait_guarded<true>(&sp->bar_mne[dg], (it & 1) ^ 1u);  // MMA2 of tile it - 1 has read these documents
          mbar_arrive_expect_tx(&sp->bar_mnf[dg], mn_chunks * 4096);
          for (int c = 0; c < mn_chunks; ++c)
            tma_load_2d(xmn + c * kMlpChunkX + dg * 4096, &map_x_mn, c * 32, row0 + dg * 32, &sp->bar_mnf[dg]);
          if (dg == 0) stamp(it, 2);
          if (dg == 3) stamp(it, 3);                  // MN-major copy requested
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc_l2 = umma_idesc_tf32(kMlpTileDocs, kMlpMaxH2, 0, 0);
      constexpr uint32_t idesc_dh = umma_idesc_tf32(kMlpTileDocs, kMlpN1, 0, 0);
      const uint32_t idesc2 = umma_idesc_tf32(64, d2_n, 0, 1);
      constexpr uint32_t idesc_b1 = umma_idesc_tf32(64, 8, 0, 0);
      const uint64_t ones_desc = umma_desc(smem_u32(ones), 16, 1024, 2);
      const uint32_t a2 = smem_u32(a2s);
      const uint32_t x_k = smem_u32(xk), x_mn = smem_u32(xmn), w2a = smem_u32(w2s), w2ta = smem_u32(w2ts);
      mbar_wait_guarded<true>(&sp->bar_w, 0);
      // layer 1 of tile `it` into TMEM buffer it & 1, chunk by chunk as the features land; every chunk is
      // handed back to the TMA warp as soon as its MMAs have run
      auto issue_mma1 = [&](int it) {
        const int b = it & 1;
        constexpr uint32_t idesc1 = umma_idesc_tf32(kMlpTileDocs, kMlpN1, 0, 0);
        const uint32_t d = tmem + b * kMlpBufCols, ws = smem_u32(w1s);
        mbar_wait_guarded<true>(&sp->bar_tempty[b], ((it >> 1) & 1) ^ 1u);
        stamp(it, 4);                                 // MMA1: TMEM buffer free
        uint32_t first = 0;
        for (int c = 0; c < nk; ++c) {
          mbar_wait_guarded<true>(&sp->bar_kf[c], it & 1);
          if (c == 0) stamp(it, 5);                   // first chunk landed
          tc_fence_after();
          const bool tail = c >= g.nfull;
          const int steps = tail ? g.tail_ksteps : 4;
          const uint32_t code = tail ? umma_layout_code(g.tail_pitch) : 2u;
          const uint32_t sbo = tail ? 8u * g.tail_pitch : 1024u;
          for (int k = 0; k < steps; ++k) {
            umma_tf32(d, umma_desc(x_k + c * kMlpChunkX + k * 32, 16, sbo, code),
                      umma_desc(ws + c * kMlpChunkW + k * 32, 16, sbo, code), idesc1, first);
            first = 1;
          }
          umma_commit(&sp->bar_ke[c]);
        }
        umma_commit(&sp->bar_tfull[b]);
        stamp(it, 6);                                 // MMA1 issued (last chunk landed)
      };
      issue_mma1(0);
      uint32_t acc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t buf = tmem + (it & 1) * kMlpBufCols;
        // Z2 = H1 W2^T (K = 64 hidden units, 8 per instruction)
        mbar_wait_guarded<true>(bar_h1, it & 1);
        stamp(it, 7);                                 // H1 ready -> MMA-L2
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kMlpN1 / 8; ++k)
          umma_tf32_ts(buf + 64, buf + 8 * k, umma_desc(w2a + (k >> 2) * (kMlpMaxH2 * 128) + (k & 3) * 32, 16, 1024, 2),
                       idesc_l2, k > 0);
        umma_commit(bar_z2);
        // dH1 = dZ2 W2 (K = 16 layer-2 units)
        mbar_wait_guarded<true>(bar_dz2, it & 1);
        stamp(it, 8);                                 // dZ2 ready -> MMA-dH
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kMlpMaxH2 / 8; ++k)
          umma_tf32_ts(buf + 96, buf + 80 + 8 * k, umma_desc(w2ta + k * 32, 16, 512, 4), idesc_dh, k > 0);
        umma_commit(bar_dh);
        // layer 1 of the next tile: its features have been streaming in since MMA1 of this tile
        if (it + 1 < my_tiles) issue_mma1(it + 1);
        // dW1 += dZ1^T X, db1 += dZ1^T 1
        mbar_wait_guarded<true>(bar_a2_full, it & 1);
        stamp(it, 9);                                 // dZ1 operand ready -> MMA2
        for (int dg = 0; dg < 4; ++dg) {
          mbar_wait_guarded<true>(&sp->bar_mnf[dg], it & 1);
          if (dg == 3) stamp(it, 10);                 // whole MN-major copy landed
          tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = 4 * dg + k4;
            const uint64_t adesc = umma_desc(a2 + dg * (kMlpN1 * 128) + k4 * 32, 16, 1024, 2);
            umma_tf32(tmem + kMlpD2Col, adesc, umma_desc(x_mn + ks * 1024, kMlpChunkX, 512, kUmmaLayout32BAtom), idesc2,
                      acc);
            umma_tf32(tmem + kMlpD2Col + d2_n, adesc, ones_desc, idesc_b1, acc);
            acc = 1;
          }
          umma_commit(&sp->bar_mne[dg]);
        }
        umma_commit(bar_a2_free);
        stamp(it, 11);                                // MMA2 issued
      }
    }
  } else if (my_tiles > 0) {
    float* ex_h1 = reinterpret_cast<float*>(a2s);     // [doc of the half tile][j], between two uses of A2
    float* ex_dz2 = ex_h1 + 64 * kMlpHPitch;          // [doc of the half tile][i]
    const int t = threadIdx.x;
    const bool docwarp = warp < 4;                    // warps 0-3: one thread per document (TMEM lane = t)
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int jb = lane / IB, ib = lane - jb * IB;
    const bool owner = lane < NJB * IB;
    float db2acc[H2], dw3acc[H2], db3acc = 0.0f;
    float2 dw2acc[BJ];
#pragma unroll
    for (int i = 0; i < H2; ++i) db2acc[i] = dw3acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < BJ; ++k) dw2acc[k] = make_float2(0.0f, 0.0f);
    unsigned char* a2row = a2s + (t >> 5) * (kMlpN1 * 128) + (t & 3) * 4;   // document warps: this document's column
    const int chunk = (t & 31) >> 2;

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t buf = tmem + (it & 1) * kMlpBufCols + lane_addr;
      float dz2[H2];
      if (docwarp) {
        const long long r = static_cast<long long>(blockIdx.x + it * gridDim.x) * kMlpTileDocs + t;
        const float ds = r < rows ? dscores[r] : 0.0f;
        mbar_wait_guarded(&sp->bar_tfull[it & 1], (it >> 1) & 1);
        if (t == 0) stamp(it, 16);                    // Z1 ready
        tc_fence_after();
        // H1 = relu(Z1 + b1), back into the same TMEM columns (units >= H1: zero weights and bias -> 0)
#pragma unroll
        for (int q = 0; q < kMlpN1 / 16; ++q) {
          uint32_t v[16];
          tmem_ld16(buf + 16 * q, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int j = 16 * q + k;
            const float h = j < H1 ? fmaxf(__uint_as_float(v[k]) + sp->b1[j < H1 ? j : 0], 0.0f) : 0.0f;
            v[k] = __float_as_uint(h);
          }
          tmem_st16(buf + 16 * q, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_h1);
        if (t == 0) stamp(it, 17);                    // H1 stored
        // dZ2 = ds * w3 * [Z2 + b2 > 0]
        mbar_wait_guarded(bar_z2, it & 1);
        if (t == 0) stamp(it, 18);                    // Z2 ready
        tc_fence_after();
        {
          uint32_t v[16];
          tmem_ld16(buf + 64, v);
          tmem_ld_wait();
          db3acc += ds;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float d = 0.0f;
            if (i < H2) {
              const float z = __uint_as_float(v[i]) + sp->b2[i];
              d = z > 0.0f ? ds * sp->w3[i] : 0.0f;
              dw3acc[i < H2 ? i : 0] = fmaf(ds, fmaxf(z, 0.0f), dw3acc[i < H2 ? i : 0]);
              db2acc[i < H2 ? i : 0] += d;
              dz2[i < H2 ? i : 0] = d;
            }
            v[i] = __float_as_uint(d);
          }
          tmem_st16(buf + 80, v);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dz2);
        if (t == 0) stamp(it, 19);                    // dZ2 stored
        // the A2 buffer (exchange arrays now, dZ1 operand below) is free once MMA2 of the previous tile has read it
        mbar_wait_guarded(bar_a2_free, (it & 1) ^ 1u);
        if (t == 0) stamp(it, 20);                    // A2 free
      }
      // dW2 += dZ2^T H1, half a tile at a time through the document-major exchange arrays, while MMA-dH runs
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (docwarp && (warp >> 1) == half) {
          const int d = t & 63;
#pragma unroll
          for (int q = 0; q < (H1 + 15) / 16; ++q) {  // H1 of the document back from TMEM, 16 units at a time
            uint32_t v[16];
            tmem_ld16(buf + 16 * q, v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; k += 4)
              if (16 * q + k < H1)
                *reinterpret_cast<uint4*>(ex_h1 + d * kMlpHPitch + 16 * q + k) = make_uint4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          }
          sts_run<H2>(ex_dz2 + d * kMlpZPitch, dz2);
        }
        epi_bar();
        if (owner) {
#pragma unroll 2
          for (int dd = 0; dd < 8; ++dd) {
            const int d = warp * 8 + dd;
            const float4* hr = reinterpret_cast<const float4*>(ex_h1 + d * kMlpHPitch + jb * BJ);
            const float2 zz = *reinterpret_cast<const float2*>(ex_dz2 + d * kMlpZPitch + 2 * ib);
#pragma unroll
            for (int q = 0; q < BJ / 4; ++q) {
              const float4 h = hr[q];
              dw2acc[4 * q + 0] = __ffma2_rn(zz, make_float2(h.x, h.x), dw2acc[4 * q + 0]);
              dw2acc[4 * q + 1] = __ffma2_rn(zz, make_float2(h.y, h.y), dw2acc[4 * q + 1]);
              dw2acc[4 * q + 2] = __ffma2_rn(zz, make_float2(h.z, h.z), dw2acc[4 * q + 2]);
              dw2acc[4 * q + 3] = __ffma2_rn(zz, make_float2(h.w, h.w), dw2acc[4 * q + 3]);
            }
          }
        }
        epi_bar();                                    // readers done: the next half / the dZ1 operand may be written
      }
      if (docwarp) {
        // dZ1 = dH1 * [H1 > 0] -> A operand of MMA2 (K-major, 128-byte swizzle: row j, 16-byte chunk
        // (doc / 4) ^ (j % 8)); rows H1..63 are rewritten with zeros (the exchange arrays passed through them)
        if (t == 0) stamp(it, 21);                    // dW2 phase done
        mbar_wait_guarded(bar_dh, it & 1);
        if (t == 0) stamp(it, 22);                    // dH1 ready
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < kMlpN1 / 16; ++q) {
          uint32_t v[16], h[16];
          tmem_ld16(buf + 96 + 16 * q, v);
          tmem_ld16(buf + 16 * q, h);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int j = 16 * q + k;
            const float dz1 = (j < H1 && __uint_as_float(h[k]) > 0.0f) ? __uint_as_float(v[k]) : 0.0f;
            *reinterpret_cast<float*>(a2row + j * 128 + ((chunk ^ (j & 7)) << 4)) = dz1;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sp->bar_tempty[it & 1]);       // every TMEM column of this tile has been read
          mbar_arrive(bar_a2_full);
        }
        if (t == 0) stamp(it, 23);                    // dZ1 stored
      }
    }

    // ---- per-CTA partial gradient vector ----
    float* out = partials + static_cast<size_t>(blockIdx.x) * partial_len;
    const int off_b1 = h1n * g.F, off_w2 = off_b1 + h1n, off_b2 = off_w2 + h2n * h1n, off_w3 = off_b2 + h2n,
              off_b3 = off_w3 + h2n;
    // db1 | db2 | dW3 | db3: warp sums, then the four document warps in order
    if (docwarp) {
      float* red = sm->red[warp];
#pragma unroll
      for (int i = 0; i < H2; ++i) {
        const float v = warp_sum(db2acc[i]), u = warp_sum(dw3acc[i]);
        if (lane == 0) {
          red[kMlpMaxH1 + i] = v;
          red[kMlpMaxH1 + kMlpMaxH2 + i] = u;
        }
      }
      const float v = warp_sum(db3acc);
      if (lane == 0) red[kMlpMaxH1 + 2 * kMlpMaxH2] = v;
      // the last MMA2 has read A2: the buffer now collects the eight warps' dW2 blocks
      mbar_wait_guarded(bar_a2_free, (my_tiles - 1) & 1);
      tc_fence_after();
    }
    epi_bar();
    float2* part2 = reinterpret_cast<float2*>(a2s);
#pragma unroll
    for (int k = 0; k < BJ; ++k) part2[(warp * 32 + lane) * BJ + k] = dw2acc[k];
    epi_bar();
    auto red4 = [&](int k) { return ((sm->red[0][k] + sm->red[1][k]) + sm->red[2][k]) + sm->red[3][k]; };
    if (t < h2n) {
      out[off_b2 + t] = red4(kMlpMaxH1 + t);
      out[off_w3 + t] = red4(kMlpMaxH1 + kMlpMaxH2 + t);
    }
    if (t == 0) out[off_b3] = red4(kMlpMaxH1 + 2 * kMlpMaxH2);
    for (int o = t; o < NJB * IB * BJ; o += kMlpBwdEpi) {        // (block lane, row of the block), warps in order
      const int bl = o / BJ, k = o - bl * BJ;
      const int j = (bl / IB) * BJ + k, i = 2 * (bl % IB);
      float2 s2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float2 v = part2[(w * 32 + bl) * BJ + k];
        s2.x += v.x;
        s2.y += v.y;
      }
      if (j < h1n) {
        if (i < h2n) out[off_w2 + i * h1n + j] = s2.x;
        if (i + 1 < h2n) out[off_w2 + (i + 1) * h1n + j] = s2.y;
      }
    }
    // dW1 out of TMEM: row j of the M = 64 accumulator lives in TMEM lane (j % 16) + 32 (j / 16); the two
    // warps of a quarter of the lanes take alternate 16-column groups
    // (the columns behind the features hold db1: every one of them the row sum of dZ1^T)
    const int j = (warp & 3) * 16 + lane;
    for (int c0 = (warp >> 2) * 16; c0 < d2_n + 8; c0 += 32) {
      uint32_t v[16];
      tmem_ld16(tmem + kMlpD2Col + c0 + lane_addr, v);
      tmem_ld_wait();
      if (lane < 16 && j < h1n) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (c0 + k < g.F) out[static_cast<size_t>(j) * g.F + c0 + k] = __uint_as_float(v[k]);
          if (c0 + k == d2_n) out[off_b1 + j] = __uint_as_float(v[k]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, tmem_cols);
}

// ---- backward from kept activations (ltr_mlp_backward with hz != NULL) --------------------------------------
// When the forward pass kept [H1 | Z2] per document (256 B for 50-10, next to 544 B of features), the backward
// pass needs neither W1 nor layer 1 again, reads every byte ONCE, and every sum over documents becomes one
// tensor-core product with the same A operand.  Per tile of 128 documents:
//   warps   dZ2 = ds * w3 * [Z2 > 0] (Z2 read from the staged activation tile) -> TMEM; dW3 / db3 in registers
//   MMA-dH  dH1 (128 x 64) = dZ2 W2          A from TMEM, W2^T in shared memory
//   warps   dZ1 = dH1 * [H1 > 0]; rows [0, H1) of the operand buffer get dZ1^T, rows [Z0, Z0 + H2) get dZ2^T
//   MMA2    [dZ1^T ; dZ2^T] (64 x 128 documents) . [X | H1 Z2 1] (128 documents x (F + 64)): the rows of dZ1
//           against X give dW1, the rows of dZ2 against H1 give dW2, either against the column of ones db1 / db2
//           (the other blocks of the product are never read).  X and the activation tile are both MN-major
//           operands exactly as TMA leaves them (128-byte swizzle, 32-byte atom); accumulators stay in TMEM.
// The masks are the forward pass's own (same H1 and Z2 bits), so the result is the gradient of exactly the
// function the forward kernel computed, up to the TF32 operands of the three products.
constexpr int kMlpHzThreads = 192;                  // 4 document warps + TMA warp + MMA warp
constexpr int kMlpHzBufCols = 80;                   // TMEM columns of one tile: dZ2 16 | dH1 64

struct MlpHzSmall {
  float b1pad[4];
  float w3[kMlpMaxH2];
  uint32_t tmem_base;
  uint32_t pad;
  uint64_t bar_mnf[4];     // feature tile, 32 documents at a time: landed / consumed by MMA2
  uint64_t bar_mne[4];
  uint64_t bar_hzf[2];     // activation tile (two stages): landed / consumed by MMA2
  uint64_t bar_hze[2];
  uint64_t bar_tempty[2];  // TMEM columns of the tile read by the warps
  uint64_t bar_dz2, bar_dh, bar_a2_full, bar_a2_free;
  float red[4][2 * kMlpMaxH2 + 4];
};

template <int H1, int H2>
__global__ void __launch_bounds__(kMlpHzThreads, 1)
mlp_backward_hz_kernel(const __grid_constant__ CUtensorMap map_x_mn, const __grid_constant__ CUtensorMap map_hz,
                       const MlpGeom g, const float* __restrict__ w2, const float* __restrict__ w3, int h1n, int h2n,
                       const float* __restrict__ dscores, long long rows, int ntiles, float* __restrict__ partials,
                       int partial_len, int nslabs, int slab_cols) {
  // Wide rows: the dW1 accumulator of one CTA holds `slab_cols` feature columns (TMEM columns, shared memory of
  // one MN-major tile), so CTA b takes column slab b % nslabs of the tiles b / nslabs, b / nslabs + nparts, ...
  // Neighbouring CTAs walk the same tiles: the activation rows they share come out of L2.  mlp_reduce_kernel
  // sums a dW1 column over the CTAs of its slab and every other gradient over the CTAs of slab 0.
  const int slab = blockIdx.x % nslabs, part = blockIdx.x / nslabs, nparts = gridDim.x / nslabs;
  const int f_off = slab * slab_cols;
  const int f_cols = g.F - f_off < slab_cols ? g.F - f_off : slab_cols;
  using Hz = MlpHz<H1, H2>;
  static_assert(Hz::kFits, "dZ2 rows must fit behind the dZ1 rows of the 64-row operand");
  constexpr int NHZ = (Hz::P + 31) / 32;            // activation chunks of 32 columns
  constexpr int HZN = 32 * NHZ;                     // accumulator columns they take
  extern __shared__ __align__(1024) unsigned char mlp_smem[];
  unsigned char* base = mlp_smem + ((1024u - (smem_u32(mlp_smem) & 1023u)) & 1023u);
  const int nk = (f_cols + 31) / 32;
  const int xn = 32 * nk;
  unsigned char* a2s = base;                        // [dZ1^T ; dZ2^T], K-major, 128-byte swizzle
  unsigned char* xmn = a2s + kMlpA2Bytes;           // feature tile, MN-major form
  unsigned char* hzs = xmn + nk * kMlpChunkX;       // two stages of the activation tile, MN-major form
  unsigned char* w2ts = hzs + 2 * NHZ * kMlpChunkX; // W2^T [j][i], K-major, 64-byte swizzle
  MlpHzSmall* sp = reinterpret_cast<MlpHzSmall*>(w2ts + kMlpN1 * 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int tmem_cols = 512;
  const uint32_t d2col = 2 * kMlpHzBufCols;         // dW1 block | dW2 block | bias block
  for (int t = threadIdx.x; t < kMlpMaxH2; t += blockDim.x) sp->w3[t] = t < h2n ? w3[t] : 0.0f;
  for (int t = threadIdx.x; t < kMlpA2Bytes / 4; t += blockDim.x) reinterpret_cast<float*>(a2s)[t] = 0.0f;
  for (int t = threadIdx.x; t < kMlpMaxH2 * kMlpMaxH1; t += blockDim.x) {
    const int i = t / kMlpMaxH1, j = t - i * kMlpMaxH1;
    const float v = (i < h2n && j < h1n) ? w2[i * h1n + j] : 0.0f;
    *reinterpret_cast<float*>(w2ts + j * 64 + ((((i >> 2) ^ ((j >> 1) & 3))) << 4) + (i & 3) * 4) = v;
  }
  if (threadIdx.x == 0) {
    for (int c = 0; c < 4; ++c) {
      mbar_init(&sp->bar_mnf[c], 1);
      mbar_init(&sp->bar_mne[c], 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(&sp->bar_hzf[c], 1);
      mbar_init(&sp->bar_hze[c], 1);
      mbar_init(&sp->bar_tempty[c], 4);
    }
    mbar_init(&sp->bar_dz2, 4);
    mbar_init(&sp->bar_dh, 1);
    mbar_init(&sp->bar_a2_full, 4);
    mbar_init(&sp->bar_a2_free, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&sp->tmem_base, tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sp->tmem_base;
  const int my_tiles = part < ntiles ? (ntiles - part + nparts - 1) / nparts : 0;

  if (warp == 4) {
    if (lane == 0 && my_tiles > 0) {
      auto load_hz = [&](int it) {                    // stage it & 1, free once MMA2 of tile it - 2 has read it
        const int s = it & 1;
        mbar_wait_guarded<true>(&sp->bar_hze[s], ((it >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&sp->bar_hzf[s], NHZ * kMlpChunkX);
        for (int c = 0; c < NHZ; ++c)
          tma_load_2d(hzs + (s * NHZ + c) * kMlpChunkX, &map_hz, c * 32, (part + it * nparts) * kMlpTileDocs,
                      &sp->bar_hzf[s]);
      };
      load_hz(0);
      for (int it = 0; it < my_tiles; ++it) {
        const int row0 = (part + it * nparts) * kMlpTileDocs;
        for (int dg = 0; dg < 4; ++dg) {
          mbar_wait_guarded<true>(&sp->bar_mne[dg], (it & 1) ^ 1u);   // MMA2 of tile it - 1 has read these documents
          mbar_arrive_expect_tx(&sp->bar_mnf[dg], nk * 4096);
          for (int c = 0; c < nk; ++c)
            tma_load_2d(xmn + c * kMlpChunkX + dg * 4096, &map_x_mn, f_off + c * 32, row0 + dg * 32, &sp->bar_mnf[dg]);
        }
        if (it + 1 < my_tiles) load_hz(it + 1);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc_dh = umma_idesc_tf32(kMlpTileDocs, kMlpN1, 0, 0);
      constexpr uint32_t idesc_hz = umma_idesc_tf32(64, HZN, 0, 1);
      const uint32_t idesc_x = umma_idesc_tf32(64, xn, 0, 1);
      const uint32_t a2 = smem_u32(a2s), x_mn = smem_u32(xmn), hz0 = smem_u32(hzs), w2ta = smem_u32(w2ts);
      uint32_t acc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        const uint32_t buf = tmem + b * kMlpHzBufCols;
        // dH1 = dZ2 W2 (K = 16 layer-2 units); the dH1 columns are free once tile it - 2 has been read
        mbar_wait_guarded<true>(&sp->bar_tempty[b], ((it >> 1) & 1) ^ 1u);
        mbar_wait_guarded<true>(&sp->bar_dz2, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kMlpMaxH2 / 8; ++k)
          umma_tf32_ts(buf + 16, buf + 8 * k, umma_desc(w2ta + k * 32, 16, 512, 4), idesc_dh, k > 0);
        umma_commit(&sp->bar_dh);
        // [dW1 | dW2 | db] += [dZ1^T ; dZ2^T] [X | H1 Z2 | 1]
        mbar_wait_guarded<true>(&sp->bar_a2_full, it & 1);   // (the warps read the activation tile before: it has landed)
        // (the issuing thread is a serial resource: 32 MMAs per tile, descriptors advanced by one add each)
        const uint64_t a_desc0 = umma_desc(a2, 16, 1024, 2);
        const uint64_t x_desc0 = umma_desc(x_mn, kMlpChunkX, 512, kUmmaLayout32BAtom);
        const uint64_t h_desc0 = umma_desc(hz0 + b * NHZ * kMlpChunkX, kMlpChunkX, 512, kUmmaLayout32BAtom);
#pragma unroll
        for (int dg = 0; dg < 4; ++dg) {
          mbar_wait_guarded<true>(&sp->bar_mnf[dg], it & 1);
          tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = 4 * dg + k4;
            // start addresses live in the low 14 bits of the descriptors, in 16-byte units
            const uint64_t adesc = a_desc0 + static_cast<uint64_t>((dg * (kMlpN1 * 128) + k4 * 32) >> 4);
            const uint64_t koff = static_cast<uint64_t>((ks * 1024) >> 4);
            umma_tf32(tmem + d2col, adesc, x_desc0 + koff, idesc_x, acc);
            umma_tf32(tmem + d2col + xn, adesc, h_desc0 + koff, idesc_hz, acc);
            acc = 1;
          }
          umma_commit(&sp->bar_mne[dg]);
        }
        umma_commit(&sp->bar_hze[b]);
        umma_commit(&sp->bar_a2_free);
      }
    }
  } else if (my_tiles > 0) {
    const int t = threadIdx.x;                        // document of the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float dw3acc[H2], db3acc = 0.0f;
#pragma unroll
    for (int i = 0; i < H2; ++i) dw3acc[i] = 0.0f;
    unsigned char* a2row = a2s + (t >> 5) * (kMlpN1 * 128) + (t & 3) * 4;   // this document's column of the operand
    const int chunk = (t & 31) >> 2;
    auto a2_store = [&](int j, float v) {
      *reinterpret_cast<float*>(a2row + j * 128 + ((chunk ^ (j & 7)) << 4)) = v;
    };
    // this document's activation row: column c sits in chunk c / 32 at 32-byte unit ((c % 32) / 8) ^ (t % 4)
    auto hz_vec = [&](const unsigned char* stage, int c) {   // c a multiple of 4
      return *reinterpret_cast<const float4*>(stage + (c >> 5) * kMlpChunkX + t * 128 +
                                              (((((c & 31) >> 3) ^ (t & 3)) << 5) | ((c & 4) << 2)));
    };

    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const uint32_t buf = tmem + b * kMlpHzBufCols + lane_addr;
      const unsigned char* stage = hzs + b * NHZ * kMlpChunkX;
      const long long r = static_cast<long long>(part + it * nparts) * kMlpTileDocs + t;
      const float ds = r < rows ? dscores[r] : 0.0f;
      mbar_wait_guarded(&sp->bar_hzf[b], (it >> 1) & 1);
      // dZ2 = ds * w3 * [Z2 > 0]
      float dz2[H2];
      {
        float z[(H2 + 3) / 4 * 4 + 4];
        constexpr int zq0 = Hz::Z0 / 4 * 4;           // Z0 is a multiple of 4
#pragma unroll
        for (int q = 0; q < (H2 + 3) / 4; ++q) {
          const float4 v = hz_vec(stage, zq0 + 4 * q);
          z[4 * q] = v.x; z[4 * q + 1] = v.y; z[4 * q + 2] = v.z; z[4 * q + 3] = v.w;
        }
        db3acc += ds;
        uint32_t v16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float d = 0.0f;
          if (i < H2) {
            d = z[i] > 0.0f ? ds * sp->w3[i] : 0.0f;
            dw3acc[i < H2 ? i : 0] = fmaf(ds, fmaxf(z[i], 0.0f), dw3acc[i < H2 ? i : 0]);
            dz2[i < H2 ? i : 0] = d;
          }
          v16[i] = __float_as_uint(d);
        }
        // the dZ2 / dH1 columns of this TMEM buffer are free once this thread's warp has read tile it - 2 (program order)
        tmem_st16(buf, v16);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sp->bar_dz2);
      // H1 > 0 of the document as bit masks (two words)
      uint32_t m0 = 0, m1 = 0;
#pragma unroll
      for (int q = 0; q < (H1 + 3) / 4; ++q) {
        const float4 v = hz_vec(stage, 4 * q);
        const uint32_t bits = (v.x > 0.0f ? 1u : 0u) | (v.y > 0.0f ? 2u : 0u) | (v.z > 0.0f ? 4u : 0u) | (v.w > 0.0f ? 8u : 0u);
        if (q < 8) m0 |= bits << (4 * q);
        else m1 |= bits << (4 * (q - 8));
      }
      // the operand buffer is free once MMA2 of the previous tile has read it
      mbar_wait_guarded(&sp->bar_a2_free, (it & 1) ^ 1u);
#pragma unroll
      for (int i = 0; i < H2; ++i) a2_store(Hz::Z0 + i, dz2[i]);
      mbar_wait_guarded(&sp->bar_dh, it & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < (H1 + 15) / 16; ++q) {
        uint32_t v[16];
        tmem_ld16(buf + 16 + 16 * q, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int j = 16 * q + k;
          if (j < H1) {
            const bool on = ((j < 32 ? m0 >> j : m1 >> (j - 32)) & 1u) != 0;
            a2_store(j, on ? __uint_as_float(v[k]) : 0.0f);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sp->bar_tempty[b]);
        mbar_arrive(&sp->bar_a2_full);
      }
    }

    // ---- per-CTA partial gradient vector ----
    float* out = partials + static_cast<size_t>(blockIdx.x) * partial_len;
    const int off_b1 = h1n * g.F, off_w2 = off_b1 + h1n, off_b2 = off_w2 + h2n * h1n, off_w3 = off_b2 + h2n,
              off_b3 = off_w3 + h2n;
    float* red = sp->red[warp];
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      const float u = warp_sum(dw3acc[i]);
      if (lane == 0) red[i] = u;
    }
    {
      const float v = warp_sum(db3acc);
      if (lane == 0) red[kMlpMaxH2] = v;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    auto red4 = [&](int k) { return ((sp->red[0][k] + sp->red[1][k]) + sp->red[2][k]) + sp->red[3][k]; };
    if (t < h2n) out[off_w3 + t] = red4(t);
    if (t == 0) out[off_b3] = red4(kMlpMaxH2);
    // accumulator rows out of TMEM: row j lives in lane (j % 16) + 32 (j / 16); rows [0, H1) are dZ1 rows (dW1,
    // db1), rows [Z0, Z0 + H2) dZ2 rows (dW2, db2)
    mbar_wait_guarded(&sp->bar_a2_free, (my_tiles - 1) & 1);
    tc_fence_after();
    const int j = warp * 16 + lane;
    const int ncols = xn + HZN;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + d2col + c0 + lane_addr, v);
      tmem_ld_wait();
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int c = c0 + k;
          const float val = __uint_as_float(v[k]);
          if (j < h1n) {
            if (c < f_cols && f_off + c < g.F) out[static_cast<size_t>(j) * g.F + f_off + c] = val;
            if (c == xn + Hz::ONE) out[off_b1 + j] = val;
          } else if (j >= Hz::Z0 && j - Hz::Z0 < h2n) {
            if (c >= xn && c - xn < h1n) out[off_w2 + (j - Hz::Z0) * h1n + (c - xn)] = val;
            if (c == xn + Hz::ONE) out[off_b2 + (j - Hz::Z0)] = val;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, tmem_cols);
}

// out[k] = sum over the CTAs' partial vectors in a fixed order: four runs of consecutive CTAs per element (one
// thread each, coalesced 256-byte rows), then ((r0 + r1) + r2) + r3.  With column slabs (nslabs > 1, kept-activation
// backward of wide rows) partial vector p belongs to slab p % nslabs: dW1 element (j, c) (k < w1_len, c = k % F) is
// summed over the vectors of slab c / slab_cols, every other element over those of slab 0.
// With a mailbox (`mine`, ltr_p2p.cuh; len <= kP2PVecCapacity) the sum leaves the kernel already all-reduced over
// the ranks of the node: the thread that owns element k pushes its value into every peer's mailbox over NVLink and
// adds up what the peers pushed for k, in rank order -- the data-parallel gradient exchange fused into the
// reduction that produces the gradient (one launch, no library collective, bit-identical on every rank).
__global__ void __launch_bounds__(256)
mlp_reduce_kernel(const float* __restrict__ partials, int nparts, int len, float* __restrict__ out, int nslabs,
                  int slab_cols, int F, int w1_len, P2PMailbox* __restrict__ mine, int rank, int world) {
  __shared__ float run[4][64];
  const int kl = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + kl;
  const int per = (nparts + 3) / 4;
  const int p0 = sub * per, p1 = p0 + per < nparts ? p0 + per : nparts;
  const unsigned int seq = mine ? p2p_vec_begin(mine) : 0u;
  float s = 0.0f;
  if (k < len) {
    const int slab = (nslabs > 1 && k < w1_len) ? (k % F) / slab_cols : 0;
    for (int p = p0; p < p1; ++p) s += partials[static_cast<size_t>(p * nslabs + slab) * len + k];
  }
  run[sub][kl] = s;
  __syncthreads();
  if (sub == 0 && k < len) {
    float v = ((run[0][kl] + run[1][kl]) + run[2][kl]) + run[3][kl];
    if (mine) v = p2p_vec_element(mine, rank, world, seq, k, v);
    out[k] = v;
  }
  if (mine) p2p_vec_finish(mine, seq);
}

}  // namespace ltr
