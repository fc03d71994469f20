// ltr_pair_tiles.cuh -- register-tiled O(L^2) pair evaluation, every unordered pair ONCE.
//
// Used by the sigmoid-weighted losses whose pair weight is symmetric and whose winner is
// decided by relevance (PairwiseLogisticLoss, LambdaARPLoss2, LambdaNDCGLoss2):
//
//     loss_b  = sum_{rel_i > rel_j} w_ij * log2(1 + exp(-sigma (s_i - s_j)))
//     lambda  = sigma / ln2 * w_ij * sigmoid(-sigma (s_winner - s_loser))
//     grad[winner] -= lambda,  grad[loser] += lambda
//
// Layout.  The query's documents sit in RANK order (descending score).  They are cut into C
// chunks of R consecutive ranks; lane l owns chunk l as ROWS (in registers).  In step m every
// lane pairs its rows with the COLUMNS of chunk (l + m) mod C, read from shared memory, so the
// C lanes cover every unordered chunk pair exactly once in floor(C/2) steps (plus the
// within-chunk triangle).  Row gradients accumulate in registers; column gradients of a step
// are added to a warp-private shared array with one vector read-modify-write (every lane
// touches a different chunk, so there are no conflicts and no atomics).
//
// Pair math (2 MUFU + 11 FP32 ops per pair, branch-free in the winner: see pair_once).
// exp(-sigma (s_i - s_j)) is FACTORED as a_i * b_j with a_i = exp(-sigma (s_i - mid)), b_j = exp(+sigma (s_j - mid)) computed once per
// document (double-precision exponent, so the product is good to a few float32 ulps); the pair
// then needs only rcp and lg2 on the MUFU pipe.  With q = a_i b_j and p = 1 + q:
//     i wins (G_i > G_j):  sigmoid(-x) = q / p,  log2(1 + e^-x) = lg2(p)
//     j wins (G_i < G_j):  sigmoid(-x) = 1 / p,  log2(1 + e^-x) = lg2(p) + (e_i - e_j)
// (e = sigma (s - mid) log2 e, so e_i - e_j = -lg2 q).  The factored form needs
// sigma * (max - min) * log2 e <= kFactoredRange so that q stays finite; queries outside that
// range take the stable form exp(-|x|) (3 MUFU) instead.
//
// Padding costs nothing per pair: a padded document is given b = 0 and G = 0 as a column and
// a = 0, G = +BIG as a row, which sends every pair that involves it down the "i wins" branch
// with q = 0, i.e. an exact zero contribution to loss and gradients.  (Stable form: padding is
// a document of gain 0 scored -1e30, whose pairs evaluate to exp(-1e30) = 0 exactly.)
//
// delta_|i-j| (LambdaNDCGLoss2, pairwise_lambda.py:206-211) only depends on the rank distance:
// a step with chunk distance d needs the 2R-1 values delta[|R d + e|], e in (-R, R), which are
// fetched as two 128-bit read-only loads from a per-R window table (PairTables, built once per
// device and L1-resident afterwards).
#pragma once

#include "ltr_common.cuh"

namespace ltr {

constexpr float kFactoredRange = 64.0f;    // max |sigma| * (max - min) * log2(e) for q = a_i * b_j
constexpr float kBigGain = 1.0e30f;

// Per-document factors of one query in shared memory, RANK order, structure of arrays so that
// the 32 lanes of a step read 32 consecutive chunks (conflict-free 128-bit loads).
struct PairSoA {
  float* a;   // exp(-sigma (s - mid))   [stable form: sigma * s]
  float* b;   // exp(+sigma (s - mid))   [stable form: unused]
  float* e;   // sigma (s - mid) log2 e  [stable form: unused]
  float* g;   // weight basis: normalised gain G (NDCG2) or float(rel) (ARP2, logistic)
};

// pair weight: 1 | |g_i - g_j| | delta |g_i - g_j| (winner by relevance); hinge; two-sided w_i, w_j
enum : int { TW_UNIT = 0, TW_DIFF = 1, TW_DELTA = 2, TW_HINGE = 3, TW_TWO = 4 };

// One pair: row (ra, re, rg) against column (cx, ce, cg), cx = b_j (factored) or sigma s_j
// (stable).  racc / cacc receive -lambda' / +lambda', lambda' > 0 when the row wins; the caller
// scales by sigma / ln 2 at the end.
template <int TW, bool FACTORED>
__device__ __forceinline__ void pair_once(float ra, float re, float rg, float cx, float ce, float cg,
                                          float dw, float& lacc, float& racc, float& cacc) {
  const float gd = rg - cg;
  if constexpr (TW == TW_HINGE) {
    // pairwise_additive.py:108-112 with ra = s_i, cx = s_j (raw scores): loss = 1.0 - (s_winner -
    // s_loser) in float32 with the reference's two roundings (negating a float32 difference is
    // exact), inactive when it is < 0; the kink loss == 0 stays active (gradient -1 / +1).
    // Padding is (score -1e30, relevance -1e30): it always loses with loss < 0.
    // Branch-free (10 FP32 ops): w = sign(gd) by clamping the integer-valued grade difference,
    // l = fma(-w, d, 1) = fl(1 - (+-d)) (w d is exact, so this is the reference's rounding),
    // cnt = l >= 0 ? w : 0 (0 for equal grades), loss += |cnt| l.
    const float d = ra - cx;
    const float w = fminf(fmaxf(gd, -1.0f), 1.0f);
    const float l = fmaf(-w, d, 1.0f);
    const float cnt = !(l < 0.0f) ? w : 0.0f;     // +1: the row wins and the pair is active
    lacc = fmaf(fabsf(cnt), l, lacc);
    racc -= cnt;
    cacc += cnt;
    return;
  }
  if constexpr (TW == TW_TWO) {
    // LambdaARPLoss1 / LambdaNDCGLoss1 (pairwise_lambda.py:114-117, :165-173): every ORDERED pair
    // carries the weight of its first document, so an unordered pair contributes
    //   w_i log2(1 + e^-x) + w_j log2(1 + e^+x),   x = sigma (s_i - s_j),
    // and d/ds_i = sigma / ln 2 * (w_j sigmoid(x) - w_i sigmoid(-x)) = -d/ds_j.
    // rg / cg hold the weights; dw is the row's validity (1, or 0 for a padded row, whose weight
    // is 0 as well); padded columns have b = 0 and weight 0.
    if constexpr (FACTORED) {
      const float q = ra * cx;               // e^-x
      const float p = q + 1.0f;
      const float r = rcp_approx(p);         // sigmoid(x)
      const float lg = lg2_approx(p);        // log2(1 + e^-x); log2(1 + e^x) = lg + (e_i - e_j)
      const float wj = cg * dw;
      lacc = fmaf(rg + wj, lg, lacc);
      lacc = fmaf(wj, re - ce, lacc);
      const float gc = r * fmaf(-rg, q, wj);
      racc += gc;
      cacc -= gc;
    } else {
      // stable form; padding is (score -1e30, weight 0): every term is 0 * finite
      const float x = ra - cx;
      const float u = -fabsf(x) * kLog2e;
      const float t = ex2_approx(u);
      const float p = 1.0f + t;
      const float r = rcp_approx(p);
      const float lg = lg2_approx(p);
      const float tr = t * r;
      const bool xpos = x >= 0.0f;
      const float l_ij = xpos ? lg : lg - u;
      const float l_ji = xpos ? lg - u : lg;
      lacc = fmaf(rg, l_ij, lacc);
      lacc = fmaf(cg, l_ji, lacc);
      const float gc = cg * (xpos ? r : tr) - rg * (xpos ? tr : r);
      racc += gc;
      cacc -= gc;
    }
    return;
  }
  float ws;   // signed weight, > 0 when the row (i) wins
  if constexpr (TW == TW_DELTA) ws = dw * gd;
  else if constexpr (TW == TW_DIFF) ws = gd;
  else ws = fminf(fmaxf(gd, -1.0f), 1.0f);   // integer grades: sign(gd)
  if constexpr (FACTORED) {
    // Branch-free in the winner (12 FP32 + 2 MUFU per pair).  With w = |ws|, r = 1 / p:
    //   row wins:     loss w lg2(p),               row gradient -w q / p = w r - w
    //   column wins:  loss w (lg2(p) + e_i - e_j), row gradient +w / p   = w r
    // so with K = max(ws, 0) (= w iff the row wins):
    //   loss += w lg2(p) + (K - ws) (e_i - e_j),   row += w r - K,   column -= w r - K.
    // Padding (q = 0, p = r = 1, lg = 0, ws >= 0) gives fma(w, 1, -w) = 0 exactly.
    const float q = ra * cx;
    const float p = q + 1.0f;
    const float r = rcp_approx(p);
    const float lg = lg2_approx(p);
    const float K = fmaxf(ws, 0.0f);
    lacc = fmaf(fabsf(ws), lg, lacc);
    lacc = fmaf(K - ws, re - ce, lacc);
    const float v = fmaf(fabsf(ws), r, -K);
    racc += v;
    cacc -= v;
    return;
  } else {
    const float d = ra - cx;                 // x if the row wins, -x otherwise
    const float u = -fabsf(d) * kLog2e;
    const float t = ex2_approx(u);
    const float p = 1.0f + t;
    const float r = rcp_approx(p);
    const float lg = lg2_approx(p);
    const bool iwins = gd > 0.0f;
    const bool xneg = iwins ? d < 0.0f : d > 0.0f;   // x = sigma (s_winner - s_loser) < 0
    const float sg = xneg ? r : t * r;
    const float l = xneg ? lg - u : lg;
    const float lp = iwins ? l : -l;           // ws * lp == |ws| * log2(1 + e^-x)
    lacc = fmaf(ws, lp, lacc);
    const float lam = ws * sg;
    racc -= lam;
    cacc += lam;
  }
}

// R consecutive floats starting at p[first] (first is a multiple of R; p is 16-byte aligned)
template <int R>
__device__ __forceinline__ void load_chunk(const float* __restrict__ p, int first, float (&v)[R]) {
  if constexpr (R == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + first);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (R == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p + first);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = p[first + r];
  }
}

// Window tables: for R rows per chunk and chunk distance d in (-kMaxChunks, kMaxChunks), entry
// (d + kMaxChunks) holds delta[|R d + e|] at slot e + R - 1 (8 floats, the last ones unused).
constexpr int kMaxChunks = 32;
__host__ __device__ constexpr int window_table_floats() { return (2 * kMaxChunks + 1) * 8; }

__device__ __forceinline__ void load_window(const float4* __restrict__ w4, float (&dwin)[8]) {
  const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
  dwin[0] = w0.x; dwin[1] = w0.y; dwin[2] = w0.z; dwin[3] = w0.w;
  dwin[4] = w1.x; dwin[5] = w1.y; dwin[6] = w1.z; dwin[7] = w1.w;
}

// R consecutive floats of chunk `chunk`; chunks are S floats apart (S = R: contiguous ranks,
// S = 4: every chunk padded to 4 slots, which makes the access one 128-bit load for every R)
template <int R, int S>
__device__ __forceinline__ void load_chunk_s(const float* __restrict__ p, int chunk, float (&v)[R]) {
  if constexpr (S == 4 && R < 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + chunk * 4);
    const float u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = u[r];
  } else {
    load_chunk<R>(p, chunk * S, v);
  }
}

// All-pairs pass over one "ring" of C <= 32 chunks of R ranks held by one warp (C*R >= n).
//   it    : rank-ordered factors of this block in shared memory, chunk c at [S c, S c + R)
//   gcol  : warp-private column-gradient accumulators, chunk c at gcol[4c .. 4c+R), zero on entry;
//           with S == 4 (the warp-per-query kernel) it has 64 chunks: lanes without a chunk dump
//           their (all-zero) column sums into chunk 32 + lane
//   wtab  : window table for this R (TW_DELTA only)
// Returns per-lane partial loss; racc[r] holds -sum lambda' of the lane's rows.
// FACTORED with S == 4 takes a fast path: a lane without rows carries a = 0 and the padding gain,
// so its pairs contribute exact zeros and nothing has to be predicated except the address of its
// column update; only the doubled last step of an even ring goes through the guarded code.
template <int TW, bool FACTORED, int R, int S = R>
__device__ __forceinline__ float ring_pass(const PairSoA& it, float* __restrict__ gcol,
                                           const float* __restrict__ wtab, int C, int n, int lane,
                                           float (&racc)[R]) {
  constexpr bool kFast = FACTORED && S == 4;
  const bool active = lane < C;
  const int me = active ? lane : 0;
  const float* colx = FACTORED ? it.b : it.a;
  float ra[R], re[R], rg[R], rv[R];
  load_chunk_s<R, S>(it.a, me, ra);
  load_chunk_s<R, S>(it.g, me, rg);
  if constexpr (FACTORED) load_chunk_s<R, S>(it.e, me, re);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rv[r] = 1.0f;
    if constexpr (FACTORED) {
      const bool valid = active && (me * R + r < n);
      ra[r] = valid ? ra[r] : 0.0f;
      rg[r] = valid ? rg[r] : (TW == TW_TWO ? 0.0f : kBigGain);
      rv[r] = valid ? 1.0f : 0.0f;
    } else {
      re[r] = 0.0f;                          // stable form: padding is already (-1e30, gain 0)
    }
    racc[r] = 0.0f;
  }
  float lacc = 0.0f;

  // within-chunk triangle (rank distance c - r > 0)
  {
    float dwin[8];
    if constexpr (TW == TW_DELTA) load_window(reinterpret_cast<const float4*>(wtab) + kMaxChunks * 2, dwin);
    float cx[R], ce[R], cg[R];
    load_chunk_s<R, S>(colx, me, cx);
    load_chunk_s<R, S>(it.g, me, cg);
    if constexpr (FACTORED) load_chunk_s<R, S>(it.e, me, ce);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = r + 1; c < R; ++c) {
        float dw = rv[r];
        if constexpr (TW == TW_DELTA) dw = dwin[c - r + R - 1];
        pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[c], FACTORED ? ce[c] : 0.0f, cg[c], dw, lacc,
                                racc[r], racc[c]);
      }
    }
  }

  const int steps = C >> 1;
  // even ring: the last step pairs chunk l with l + C/2 from both ends; only the lower half commits
  const bool even = (C & 1) == 0;
  const bool dup_last = even && lane >= steps;
  int pc = me;
  int m = 1;
  if constexpr (kFast) {
    const int nfast = even ? steps - 1 : steps;
    for (; m <= nfast; ++m) {
      pc = pc + 1 == C ? 0 : pc + 1;
      float dwin[8];
      if constexpr (TW == TW_DELTA)
        load_window(reinterpret_cast<const float4*>(wtab) + (pc - me + kMaxChunks) * 2, dwin);
      float4* g4 = reinterpret_cast<float4*>(gcol) + (active ? pc : 32 + lane);
      const float4 gold = *g4;
      float cx[R], ce[R], cg[R];
      load_chunk_s<R, S>(colx, pc, cx);
      load_chunk_s<R, S>(it.g, pc, cg);
      load_chunk_s<R, S>(it.e, pc, ce);
      float tc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = 0; c < R; ++c) {
          float dw = rv[r];
          if constexpr (TW == TW_DELTA) dw = dwin[c - r + R - 1];
          pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[c], ce[c], cg[c], dw, lacc, racc[r], tc[c]);
        }
      }
      *g4 = make_float4(gold.x + tc[0], gold.y + tc[1], gold.z + tc[2], gold.w + tc[3]);
      __syncwarp();
    }
  }
  for (; m <= steps; ++m) {
    pc = pc + 1 == C ? 0 : pc + 1;
    const bool commit = active && !(dup_last && m == steps);
    float dwin[8];
    if constexpr (TW == TW_DELTA)
      load_window(reinterpret_cast<const float4*>(wtab) + (pc - me + kMaxChunks) * 2, dwin);
    float cx[R], ce[R], cg[R];
    load_chunk_s<R, S>(colx, pc, cx);
    load_chunk_s<R, S>(it.g, pc, cg);
    if constexpr (FACTORED) load_chunk_s<R, S>(it.e, pc, ce);
    float tl = 0.0f, tr[R], tc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { tr[r] = 0.0f; tc[r] = 0.0f; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float dw = rv[r];
        if constexpr (TW == TW_DELTA) dw = dwin[c - r + R - 1];
        pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[c], FACTORED ? ce[c] : 0.0f, cg[c], dw, tl, tr[r],
                                tc[c]);
      }
    }
    if (commit) {
      lacc += tl;
#pragma unroll
      for (int r = 0; r < R; ++r) racc[r] += tr[r];
      // column gradients: chunk pc owns gcol[4 pc .. 4 pc + 3] (one vector read-modify-write)
      if constexpr (R == 4) {
        float4* g4 = reinterpret_cast<float4*>(gcol) + pc;
        float4 g = *g4;
        g.x += tc[0]; g.y += tc[1]; g.z += tc[2]; g.w += tc[3];
        *g4 = g;
      } else if constexpr (R == 2) {
        float2* g2 = reinterpret_cast<float2*>(gcol + pc * 4);
        float2 g = *g2;
        g.x += tc[0]; g.y += tc[1];
        *g2 = g;
      } else {
#pragma unroll
        for (int c = 0; c < R; ++c) gcol[pc * 4 + c] += tc[c];
      }
    }
    __syncwarp();
  }
  return active ? lacc : 0.0f;
}

// Off-diagonal 128 x 128 rank tile: rows = the 32 chunks of 4 ranks starting at rank `row_base`
// (all valid: only a query's last 128-rank block can be partial), columns = the 32 chunks
// starting at rank `col_base` > row_base.  In step m lane l meets column chunk (l + m) & 31, so
// the 32 lanes touch 32 different chunks in every step.  Padded columns contribute exact zeros.
// delta windows come straight from the raw table: rank distances are positive multiples of 4
// plus (-3..3), i.e. two aligned 128-bit loads.
//   gcol : warp-private [128] column accumulators of this tile, zero on entry
template <int TW, bool FACTORED>
__device__ __forceinline__ float tile_pass(const PairSoA& it, int row_base, int col_base,
                                           const float* __restrict__ delta, float* __restrict__ gcol,
                                           int lane, float (&racc)[4]) {
  constexpr int R = 4;
  const float* colx = FACTORED ? it.b : it.a;
  float ra[R], re[R], rg[R];
  load_chunk<R>(it.a, row_base + lane * R, ra);
  load_chunk<R>(it.g, row_base + lane * R, rg);
  if constexpr (FACTORED) load_chunk<R>(it.e, row_base + lane * R, re);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if constexpr (!FACTORED) re[r] = 0.0f;
    racc[r] = 0.0f;
  }
  float lacc = 0.0f;
  const int dist0 = (col_base - row_base) >> 2;   // chunk distance of column chunk 0 from row chunk 0
  for (int m = 0; m < 32; ++m) {
    const int pc = (lane + m) & 31;
    float dwin[8];
    if constexpr (TW == TW_DELTA) {
      // delta[4d - 4 .. 4d + 3]: slot e + 4, so the window (e = -3..3) starts one float in
      const float4* w4 = reinterpret_cast<const float4*>(delta) + (dist0 + pc - lane - 1);
      const float4 w0 = *w4, w1 = *(w4 + 1);
      dwin[0] = w0.y; dwin[1] = w0.z; dwin[2] = w0.w;
      dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w; dwin[7] = 0.0f;
    }
    float cx[R], ce[R], cg[R];
    load_chunk<R>(colx, col_base + pc * R, cx);
    load_chunk<R>(it.g, col_base + pc * R, cg);
    if constexpr (FACTORED) load_chunk<R>(it.e, col_base + pc * R, ce);
    float tc[R];
#pragma unroll
    for (int c = 0; c < R; ++c) tc[c] = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float dw = 1.0f;
        if constexpr (TW == TW_DELTA) dw = dwin[c - r + R - 1];
        pair_once<TW, FACTORED>(ra[r], re[r], rg[r], cx[c], FACTORED ? ce[c] : 0.0f, cg[c], dw, lacc,
                                racc[r], tc[c]);
      }
    }
    float4* g4 = reinterpret_cast<float4*>(gcol) + pc;
    float4 g = *g4;
    g.x += tc[0]; g.y += tc[1]; g.z += tc[2]; g.w += tc[3];
    *g4 = g;
    __syncwarp();
  }
  return lacc;
}

}  // namespace ltr
