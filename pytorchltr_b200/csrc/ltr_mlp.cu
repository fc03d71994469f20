// ltr_mlp.cu -- host side of the MLP scorer entry points of libltr_sm100.so (kernels: ltr_mlp_scorer.cuh).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "ltr_common.cuh"
#include "ltr_host.cuh"
#include "ltr_mlp_scorer.cuh"
#include "ltr_sm100.h"

using namespace ltr;

// ---- MLP scorer (tcgen05 layer 1): tensor maps and launches -------------------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query, so the library
// still links nothing but the static CUDA runtime.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_tiled_fn(EncodeTiledFn* out) {
  static std::atomic<void*> cached{nullptr};
  void* fn = cached.load(std::memory_order_acquire);
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    LTR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) return LTR_EUNSUPPORTED;
    cached.store(fn, std::memory_order_release);
  }
  *out = reinterpret_cast<EncodeTiledFn>(fn);
  return LTR_OK;
}

// a row-major float32 matrix [nrows, ncols] seen as boxes of `box_cols` x `box_rows`, rows `pitch` bytes
// apart in shared memory under the matching swizzle; out-of-range elements read as zero
static int make_map_2d(CUtensorMap* map, const float* base, long long nrows, int ncols, int box_cols, int box_rows,
                       int swizzle_override = -1) {
  EncodeTiledFn encode = nullptr;
  int rc = encode_tiled_fn(&encode);
  if (rc != LTR_OK) return rc;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ncols), static_cast<cuuint64_t>(nrows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ncols) * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const int pitch = box_cols * 4;
  CUtensorMapSwizzle sw = pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (swizzle_override >= 0) sw = static_cast<CUtensorMapSwizzle>(swizzle_override);
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tls_cuda_error = static_cast<int>(r);
    return LTR_ECUDA;
  }
  return LTR_OK;
}

static MlpGeom mlp_geometry(int F, int stages) {
  MlpGeom g;
  g.F = F;
  g.nfull = F / 32;
  const int w = F - 32 * g.nfull;
  g.tail_pitch = w == 0 ? 0 : (w <= 8 ? 32 : (w <= 16 ? 64 : 128));
  g.tail_ksteps = (w + 7) / 8;
  g.stage_bytes = g.nfull * kMlpChunkX + kMlpTileDocs * g.tail_pitch;
  g.w1_bytes = g.nfull * kMlpChunkW + kMlpN1 * g.tail_pitch;
  g.stages = stages;
  g.slab_chunks = 0;
  g.nchunks = (F + 31) / 32;
  return g;
}

// LTR_MLP_TRACE=1: event timestamps of CTA 0 of the backward kernel (tools/mlp_trace.py reads them back
// through ltr_mlp_trace_read)
static long long* mlp_trace_buffer() {
  static long long* buf = [] {
    const char* v = getenv("LTR_MLP_TRACE");
    long long* p = nullptr;
    if (v && v[0] == '1' && cudaMalloc(&p, 24 * 32 * sizeof(long long)) == cudaSuccess)
      cudaMemset(p, 0, 24 * 32 * sizeof(long long));
    return p;
  }();
  return buf;
}

constexpr int kMlpMaxCtas = 256;   // partial gradient vectors in the backward workspace

struct MlpMaps { CUtensorMap x, x_tail, w, w_tail; };

static int mlp_make_maps(MlpMaps* m, const MlpGeom& g, const float* features, long long rows, const float* w1,
                         int H1) {
  int rc = LTR_OK;
  const int tail_cols = g.tail_pitch ? g.tail_pitch / 4 : 32;
  // a geometry without full chunks (F < 32) or without a tail still gets valid (unused) maps
  rc = make_map_2d(&m->x, features, rows, g.F, g.F < 32 ? tail_cols : 32, kMlpTileDocs);
  if (rc != LTR_OK) return rc;
  rc = make_map_2d(&m->x_tail, features, rows, g.F, tail_cols, kMlpTileDocs);
  if (rc != LTR_OK) return rc;
  rc = make_map_2d(&m->w, w1, H1, g.F, g.F < 32 ? tail_cols : 32, kMlpN1);
  if (rc != LTR_OK) return rc;
  return make_map_2d(&m->w_tail, w1, H1, g.F, tail_cols, kMlpN1);
}

static int mlp_check(const float* features, long long rows, int F, const float* w1, int H1, const float* w2, int H2,
                     const float* w3) {
  if (rows < 0 || F < 1 || H1 < 1 || H2 < 1) return LTR_EINVAL;
  if (rows > 0 && (!features || !w1 || !w2 || !w3)) return LTR_EINVAL;
  // TMA tensor copies: rows of a multiple of 16 bytes on 16-byte boundaries; TMEM tile of 64 hidden units
  if (F % 4 != 0 || !aligned16(features) || !aligned16(w1)) return LTR_EUNSUPPORTED;
  if (H1 > kMlpMaxH1 || H2 > kMlpMaxH2 || rows > (1LL << 31) - kMlpTileDocs) return LTR_EUNSUPPORTED;
  return LTR_OK;
}

template <int H1, int H2>
static int launch_mlp_scores(const MlpMaps& m, const MlpGeom& g, const float* b1, const float* w2, const float* b2,
                             const float* w3, const float* b3, int h1, int h2, long long rows, float* scores_out,
                             float* hz_out, cudaStream_t st, const DeviceInfo& di) {
  const size_t smem = 1024 + static_cast<size_t>(g.w1_bytes) + static_cast<size_t>(g.stages) * g.stage_bytes +
                      sizeof(MlpSmallParams);
  LTR_CUDA(cudaFuncSetAttribute(mlp_scores_kernel<H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int ntiles = static_cast<int>((rows + kMlpTileDocs - 1) / kMlpTileDocs);
  const int grid = ntiles < di.sms ? ntiles : di.sms;
  mlp_scores_kernel<H1, H2><<<grid, kMlpFwdThreads, smem, st>>>(m.x, m.x_tail, m.w, m.w_tail, g, b1, w2, b2, w3, b3,
                                                               h1, h2, rows, ntiles, scores_out, hz_out);
  LTR_CUDA(cudaGetLastError());
  return LTR_OK;
}

static int mlp_hz_pitch(int H1, int H2) {
  if (H1 < 1 || H2 < 1) return 0;
  if (H1 <= 32 && H2 <= 8) return MlpHz<32, 8>::P;
  if (H1 <= 50 && H2 <= 10) return MlpHz<50, 10>::P;
  return 0;
}

template <int H1, int H2>
static int launch_mlp_backward_hz(const float* features, const float* hz, long long rows, int F, const float* w2,
                                  const float* w3, int h1, int h2, const float* dscores, float* partials, int len,
                                  int* nparts_out, int* nslabs_out, int* slab_cols_out, cudaStream_t st,
                                  const DeviceInfo& di) {
  using Hz = MlpHz<H1, H2>;
  constexpr int NHZ = (Hz::P + 31) / 32;
  // feature columns per CTA: bounded by the TMEM columns of the dW1 accumulator and by the shared memory of one
  // MN-major tile; wider rows are split into column slabs over the CTAs
  const int nk_all = (F + 31) / 32;
  int nk_max = (512 - 2 * kMlpHzBufCols - 32 * NHZ) / 32;
  const size_t fixed = 1024 + kMlpA2Bytes + static_cast<size_t>(2 * NHZ) * kMlpChunkX + kMlpN1 * 64 + sizeof(MlpHzSmall);
  const int nk_smem = static_cast<int>((227u * 1024u - fixed) / kMlpChunkX);
  if (nk_smem < nk_max) nk_max = nk_smem;
  if (nk_max < 1) return LTR_EUNSUPPORTED;
  const int nslabs = (nk_all + nk_max - 1) / nk_max;
  const int nk = (nk_all + nslabs - 1) / nslabs;
  const size_t smem = fixed + static_cast<size_t>(nk) * kMlpChunkX;
  CUtensorMap map_x, map_hz;
  int rc = make_map_2d(&map_x, features, rows, F, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != LTR_OK) return rc;
  rc = make_map_2d(&map_hz, hz, rows, Hz::P, 32, kMlpTileDocs, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != LTR_OK) return rc;
  MlpGeom g = mlp_geometry(F, 1);
  const int ntiles = static_cast<int>((rows + kMlpTileDocs - 1) / kMlpTileDocs);
  // CTA b: column slab b % nslabs of the tiles b / nslabs, b / nslabs + nparts, ...
  int cap = di.sms < kMlpMaxCtas ? di.sms : kMlpMaxCtas;
  if (cap < nslabs) return LTR_EUNSUPPORTED;
  int nparts = cap / nslabs;
  if (nparts > ntiles) nparts = ntiles;
  const int grid = nparts * nslabs;
  LTR_CUDA(cudaFuncSetAttribute(mlp_backward_hz_kernel<H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  mlp_backward_hz_kernel<H1, H2><<<grid, kMlpHzThreads, smem, st>>>(map_x, map_hz, g, w2, w3, h1, h2, dscores, rows,
                                                                    ntiles, partials, len, nslabs, 32 * nk);
  LTR_CUDA(cudaGetLastError());
  *nparts_out = nparts;
  *nslabs_out = nslabs;
  *slab_cols_out = 32 * nk;
  return LTR_OK;
}

extern "C" {

// floats per document of the activation rows ltr_mlp_scores can keep for ltr_mlp_backward (0: this shape has none)
int ltr_mlp_hz_pitch(int H1, int H2) { return mlp_hz_pitch(H1, H2); }

int ltr_mlp_scores(const float* features, long long rows, int F, const float* w1, const float* b1, int H1,
                              const float* w2, const float* b2, int H2, const float* w3, const float* b3,
                              float* scores_out, float* hz_out, void* stream) {
  int rc = mlp_check(features, rows, F, w1, H1, w2, H2, w3);
  if (rc != LTR_OK) return rc;
  if (rows > 0 && !scores_out) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  if (rows == 0) return LTR_OK;
  const size_t budget = 227u * 1024u - 1024u - sizeof(MlpSmallParams);
  MlpGeom g = mlp_geometry(F, 1);
  MlpMaps m;
  // W1 resident beside whole feature tiles while two of them fit (rows up to 192 floats); beyond that an
  // inference call streams W1 with the tile through a ring of four 2-chunk slabs (measured: 1.2-1.4x the
  // single-stage resident form at 200..288 features, tools/mlp_slab_ab.sh), a call that also writes the activation
  // rows stays resident until a tile no longer fits (288 features), then streams as well
  int slab_chunks = 2, slab_stages = 4;
  const size_t resident = static_cast<size_t>(g.w1_bytes) + g.stage_bytes;
  bool slabs = resident > budget || (!hz_out && resident + g.stage_bytes > budget && F >= 32);
  if (const char* v = getenv("LTR_MLP_SLAB")) {         // "chunks,stages" (A/B runs): force the streamed-W1 form
    if (sscanf(v, "%d,%d", &slab_chunks, &slab_stages) == 2 && slab_chunks >= 1 && slab_stages >= 1 &&
        slab_stages <= 8 && static_cast<size_t>(slab_chunks) * slab_stages * (kMlpChunkX + kMlpChunkW) <= budget)
      slabs = F >= 32;
    else
      slab_chunks = 2, slab_stages = 4;
  }
  if (slabs) {
    // rows too wide for W1 and a whole tile side by side: stream both in slabs of two chunks, four stages
    if (F < 32) return LTR_EUNSUPPORTED;
    g.slab_chunks = slab_chunks;
    g.stages = slab_stages;
    g.stage_bytes = g.slab_chunks * (kMlpChunkX + kMlpChunkW);
    g.w1_bytes = 0;
  } else {
    const int stages = static_cast<int>((budget - g.w1_bytes) / g.stage_bytes);
    int cap = 4;                                        // (8 barriers exist; deeper rings measured no faster)
    if (const char* v = getenv("LTR_MLP_STAGES")) cap = atoi(v) >= 1 && atoi(v) <= 8 ? atoi(v) : cap;
    g.stages = stages > cap ? cap : stages;
  }
  rc = mlp_make_maps(&m, g, features, rows, w1, H1);
  if (rc != LTR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (hz_out && (ltr_mlp_hz_pitch(H1, H2) == 0 || !aligned16(hz_out))) return LTR_EINVAL;
  // the smallest instantiation that holds the model (hidden units beyond H1 / H2 are zero weights)
  if (H1 <= 32 && H2 <= 8)
    return launch_mlp_scores<32, 8>(m, g, b1, w2, b2, w3, b3, H1, H2, rows, scores_out, hz_out, st, di);
  if (H1 <= 50 && H2 <= 10)
    return launch_mlp_scores<50, 10>(m, g, b1, w2, b2, w3, b3, H1, H2, rows, scores_out, hz_out, st, di);
  return launch_mlp_scores<64, 16>(m, g, b1, w2, b2, w3, b3, H1, H2, rows, scores_out, nullptr, st, di);
}

size_t ltr_mlp_grad_len(int F, int H1, int H2) {
  if (F < 1 || H1 < 1 || H2 < 1) return 0;
  return static_cast<size_t>(H1) * F + H1 + static_cast<size_t>(H2) * H1 + 2 * static_cast<size_t>(H2) + 1;
}

size_t ltr_mlp_workspace_bytes(int F, int H1, int H2) {
  return ltr_mlp_grad_len(F, H1, H2) * kMlpMaxCtas * sizeof(float);
}

}  // extern "C"

// p2p (or NULL): the ranks' gradients are summed over NVLink peer memory inside the final reduction
static int mlp_backward_impl(const float* features, long long rows, int F, const float* w1, const float* b1, int H1,
                             const float* w2, const float* b2, int H2, const float* w3, const float* b3,
                             const float* hz, const float* dscores, float* grads_out, void* workspace,
                             size_t workspace_bytes, void* stream, ltr_p2p* p2p) {
  int rc = mlp_check(features, rows, F, w1, H1, w2, H2, w3);
  if (rc != LTR_OK) return rc;
  if (!grads_out || !workspace || (rows > 0 && !dscores)) return LTR_EINVAL;
  if (workspace_bytes < ltr_mlp_workspace_bytes(F, H1, H2)) return LTR_EINVAL;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc != LTR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int len = static_cast<int>(ltr_mlp_grad_len(F, H1, H2));
  // the exchange rides in the reduction when the vector fits one mailbox piece, else it follows it
  const bool fused_x = p2p && len <= kP2PVecCapacity;
  P2PMailbox* box = fused_x ? p2p->mine : nullptr;
  const int xrank = p2p ? p2p->rank : 0, xworld = p2p ? p2p->world : 1;
  auto after = [&]() -> int {
    return (p2p && !fused_x) ? ltr_p2p_allreduce_vec(p2p, grads_out, len, stream) : LTR_OK;
  };
  if (rows == 0) {
    // a rank without documents still takes part in the exchange
    LTR_CUDA(cudaMemsetAsync(grads_out, 0, sizeof(float) * len, st));
    return p2p ? ltr_p2p_allreduce_vec(p2p, grads_out, len, stream) : LTR_OK;
  }
  if (hz) {
    // the forward pass kept [H1 | Z2]: every byte once, no layer 1 again
    if (ltr_mlp_hz_pitch(H1, H2) == 0 || !aligned16(hz)) return LTR_EINVAL;
    int nparts = 0, nslabs = 1, slab_cols = 0;
    rc = (H1 > 32 || H2 > 8) ? launch_mlp_backward_hz<50, 10>(features, hz, rows, F, w2, w3, H1, H2, dscores,
                                                   static_cast<float*>(workspace), len, &nparts, &nslabs, &slab_cols,
                                                   st, di)
                  : launch_mlp_backward_hz<32, 8>(features, hz, rows, F, w2, w3, H1, H2, dscores,
                                                  static_cast<float*>(workspace), len, &nparts, &nslabs, &slab_cols,
                                                  st, di);
    if (rc != LTR_OK) return rc;
    mlp_reduce_kernel<<<(len + 63) / 64, 256, 0, st>>>(static_cast<float*>(workspace), nparts, len, grads_out, nslabs,
                                                       slab_cols, F, H1 * F, box, xrank, xworld);
    LTR_CUDA(cudaGetLastError());
    return after();
  }
  MlpGeom g = mlp_geometry(F, 1);
  const int mn_chunks = (F + 31) / 32;
  const int d2_cols = mn_chunks * 32;
  if (kMlpD2Col + d2_cols + 16 > 512 || g.nfull + 1 > 8) return LTR_EUNSUPPORTED;   // TMEM columns: dW1 | db1
  // W1, the dZ1 operand and the two copies of a feature tile (K-major for MMA1, MN-major for MMA2)
  const size_t smem = 1024 + static_cast<size_t>(g.w1_bytes) + kMlpA2Bytes + g.stage_bytes +
                      static_cast<size_t>(mn_chunks) * kMlpChunkX + kMlpW2Bytes + kMlpOnesBytes + sizeof(MlpBwdSmem);
  if (smem > 227u * 1024u) return LTR_EUNSUPPORTED;
  MlpMaps m;
  rc = mlp_make_maps(&m, g, features, rows, w1, H1);
  if (rc != LTR_OK) return rc;
  CUtensorMap map_mn;
  rc = make_map_2d(&map_mn, features, rows, F, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);   // 32 documents a box
  if (rc != LTR_OK) return rc;
  const int tmem_cols = 512;
  const int ntiles = static_cast<int>((rows + kMlpTileDocs - 1) / kMlpTileDocs);
  int grid = ntiles < di.sms ? ntiles : di.sms;
  if (grid > kMlpMaxCtas) grid = kMlpMaxCtas;
  float* partials = static_cast<float*>(workspace);
#define LTR_MLP_BWD(A, B)                                                                                          \
  do {                                                                                                             \
    LTR_CUDA(cudaFuncSetAttribute(mlp_backward_kernel<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  static_cast<int>(smem)));                                                        \
    mlp_backward_kernel<A, B><<<grid, kMlpBwdThreads, smem, st>>>(m.x, m.x_tail, map_mn, m.w, m.w_tail, g, b1, w2,  \
                                                                  b2, w3, b3, H1, H2, dscores, rows, ntiles,       \
                                                                  tmem_cols, partials, len, mlp_trace_buffer());   \
  } while (0)
  if (H1 <= 32 && H2 <= 8) LTR_MLP_BWD(32, 8);
  else if (H1 <= 50 && H2 <= 10) LTR_MLP_BWD(50, 10);
  else LTR_MLP_BWD(64, 16);
#undef LTR_MLP_BWD
  LTR_CUDA(cudaGetLastError());
  const int rgrid = (len + 63) / 64;
  mlp_reduce_kernel<<<rgrid, 256, 0, st>>>(partials, grid, len, grads_out, 1, F, F, H1 * F, box, xrank, xworld);
  LTR_CUDA(cudaGetLastError());
  return after();
}

extern "C" {

int ltr_mlp_backward(const float* features, long long rows, int F, const float* w1, const float* b1, int H1,
                     const float* w2, const float* b2, int H2, const float* w3, const float* b3, const float* hz,
                     const float* dscores, float* grads_out, void* workspace, size_t workspace_bytes, void* stream) {
  return mlp_backward_impl(features, rows, F, w1, b1, H1, w2, b2, H2, w3, b3, hz, dscores, grads_out, workspace,
                           workspace_bytes, stream, nullptr);
}

int ltr_mlp_backward_allreduce(const float* features, long long rows, int F, const float* w1, const float* b1, int H1,
                               const float* w2, const float* b2, int H2, const float* w3, const float* b3,
                               const float* hz, const float* dscores, float* grads_out, void* workspace,
                               size_t workspace_bytes, ltr_p2p* p2p, void* stream) {
  if (!p2p) return LTR_EINVAL;
  return mlp_backward_impl(features, rows, F, w1, b1, H1, w2, b2, H2, w3, b3, hz, dscores, grads_out, workspace,
                           workspace_bytes, stream, p2p);
}

// debugging aid, not part of the public header: copies the trace of the last traced backward launch
int ltr_mlp_trace_read(long long* host_out /* 24 * 32 */) {
  long long* p = mlp_trace_buffer();
  if (!p) return LTR_EINVAL;
  LTR_CUDA(cudaMemcpy(host_out, p, 24 * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  return LTR_OK;
}

}  // extern "C"
