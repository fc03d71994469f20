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
// Pair math of the factored form (pair_fact): per unordered pair 1.5 MUFU + 10.5 FP32 lane-operations
// + 2 max on the alu pipe, issued as packed f32x2 instructions on TWO adjacent columns at a time
// (FFMA2 / FADD2 / FMUL2: abs, negation, the broadcast of a row value and the swap of the two halves
// are operand modifiers), branch-free in the winner: 178 SASS instructions per 16 pairs per lane
// (measured ceilings of the pipes: tools/issue_peak.cu, profiles/issue_peaks.json).
// exp(-sigma (s_i - s_j)) is FACTORED as a_i * b_j with a_i = exp(-sigma (s_i - mid)), b_j = exp(+sigma (s_j - mid)) computed once per
// document (double-precision exponent, so the product is good to a few float32 ulps); the pair
// then needs only rcp and lg2 on the MUFU pipe, and the two pairs of a packed instruction share ONE
// reciprocal: R = 1 / (p_0 p_1), 1 / p_0 = R p_1, 1 / p_1 = R p_0.  With q = a_i b_j and p = 1 + q:
//     i wins (G_i > G_j):  sigmoid(-x) = q / p,  log2(1 + e^-x) = lg2(p)
//     j wins (G_i < G_j):  sigmoid(-x) = 1 / p,  log2(1 + e^-x) = lg2(p) + (e_i - e_j)
// (e = sigma (s - mid) log2 e, so e_i - e_j = -lg2 q).  The factored form needs
// sigma * (max - min) * log2 e <= kFactoredRange so that p_0 p_1 stays finite; queries outside that
// range take the stable form exp(-|x|) (3 MUFU, scalar) instead.
// Winner-by-relevance losses, ws = signed pair weight (> 0 when the row wins), w = |ws|:
//     loss += w lg2(p) + max(-ws, 0) (e_i - e_j),   row += w r - max(ws, 0),   column -= the same
// (9 packed FP32 operations; the two max run on the alu pipe, beside the FP32 datapath).
//
// Padding: a padded ROW carries a = 0 (so p = 1, lg2 = 0, r = 1 exactly: both pairs of its packed
// instruction have p = 1) and the gain +BIG (it "wins" everything: max(-ws, 0) = 0, w r - max(ws, 0) = 0);
// a padded COLUMN carries b = 0 and the smallest gain of the query's valid documents (it loses
// everything: max(-ws, 0) = 0 and w lg2(1) = 0).  Every pair that involves padding contributes an exact
// zero to the loss; a padded column that shares its reciprocal with a valid one sees r = 1 +- 1 ulp
// instead of 1 and leaves w (r - 1) ~ 1e-7 of ONE pair weight in that row's gradient (once per row at
// most: only the last chunk of a query mixes valid and padded columns).
// (Stable form: padding is a document of gain 0 scored -1e30, whose pairs evaluate to exp(-1e30) = 0.)
//
// delta_|i-j| (LambdaNDCGLoss2, pairwise_lambda.py:206-211) only depends on the rank distance:
// a step with chunk distance d needs the 2R-1 values delta[|R d + e|], e in (-R, R), which are
// fetched as two 128-bit read-only loads from a per-R window table (PairTables, built once per
// device and L1-resident afterwards).
#pragma once

#include "ltr_common.cuh"

namespace ltr {

constexpr float kFactoredRange = 60.0f;    // max |sigma| * (max - min) * log2(e): p_0 * p_1 <= 2^121 stays finite
constexpr float kBigGain = 1.0e30f;

// winner decided by relevance (pair_fact)
__host__ __device__ constexpr bool tw_winner(int tw) { return tw == 0 || tw == 1 || tw == 2; }   // UNIT, DIFF, DELTA

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

// ---- one pair (float) or the pairs of two adjacent columns (float2, packed f32x2 arithmetic) -----------
template <typename V> __device__ __forceinline__ V vsplat(float x);
template <> __device__ __forceinline__ float vsplat<float>(float x) { return x; }
template <> __device__ __forceinline__ float2 vsplat<float2>(float x) { return make_float2(x, x); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float vabs(float a) { return fabsf(a); }
__device__ __forceinline__ float2 vabs(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float vclamp_one(float a) { return fminf(fmaxf(a, -1.0f), 1.0f); }
__device__ __forceinline__ float2 vclamp_one(float2 a) { return make_float2(vclamp_one(a.x), vclamp_one(a.y)); }
__device__ __forceinline__ float vrelu(float a) { return fmaxf(a, 0.0f); }
__device__ __forceinline__ float2 vrelu(float2 a) { return make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)); }
__device__ __forceinline__ float vhsum(float a) { return a; }
__device__ __forceinline__ float vhsum(float2 a) { return a.x + a.y; }
// p -> (1 / p, lg2 p); the two pairs of the packed form share one reciprocal
__device__ __forceinline__ void rcp_lg2(float p, float& r, float& lg) {
  r = rcp_approx(p);
  lg = lg2_approx(p);
}
__device__ __forceinline__ void rcp_lg2(float2 p, float2& r, float2& lg) {
  const float R = rcp_approx(p.x * p.y);
  r = __fmul2_rn(make_float2(R, R), make_float2(p.y, p.x));   // one FMUL2: broadcast R, swapped halves of p
  lg = make_float2(lg2_approx(p.x), lg2_approx(p.y));
}

// Factored pair(s): row (ra, re, rg, rv) against column(s) (cx = b_j, ce, cg); dwh = delta window
// value(s) (TW_DELTA).  TW_TWO: rv = validity of the row (1 or 0).  racc / cacc receive -lambda' / +lambda'.
template <int TW, typename V>
__device__ __forceinline__ void pair_fact(V ra, V re, V rg, V rv, V cx, V ce, V cg, V dwh, V& lacc, V& racc,
                                          V& cacc) {
  if constexpr (TW == TW_TWO) {
    // w_i log2(1 + e^-x) + w_j log2(1 + e^+x), x = sigma (s_i - s_j) (see pair_once)
    const V q = vmul(ra, cx);
    const V p = vadd(q, vsplat<V>(1.0f));
    V r, lg;
    rcp_lg2(p, r, lg);
    const V wj = vmul(cg, rv);
    lacc = vfma(vadd(rg, wj), lg, lacc);
    lacc = vfma(wj, vadd(re, vneg(ce)), lacc);
    const V gc = vmul(r, vfma(vneg(rg), q, wj));
    racc = vadd(racc, gc);
    cacc = vadd(cacc, vneg(gc));
  } else {
    // ws = signed weight (> 0: the row wins), w = |ws|, K = max(ws, 0), D = max(-ws, 0) = K - ws:
    //   loss += w lg2(p) + D (e_i - e_j),   row += w r - K,   column -= the same.
    // K and D are max operations: they run on the alu pipe beside the packed FP32 datapath.
    const V gd = vadd(rg, vneg(cg));
    V ws;
    if constexpr (TW == TW_DELTA) ws = vmul(dwh, gd);
    else if constexpr (TW == TW_DIFF) ws = gd;
    else ws = vclamp_one(gd);                   // integer grades: sign(gd)
    const V p = vfma(ra, cx, vsplat<V>(1.0f));
    V r, lg;
    rcp_lg2(p, r, lg);
    lacc = vfma(vabs(ws), lg, lacc);
    lacc = vfma(vrelu(vneg(ws)), vadd(re, vneg(ce)), lacc);
    const V v = vfma(vabs(ws), r, vneg(vrelu(ws)));
    racc = vadd(racc, v);
    cacc = vadd(cacc, vneg(v));
  }
}

// Per-document factors of the factored form in pair_fact's conventions.  w = weight basis of the
// document (normalised gain G, float(rel), or the ordered-pair weight of TW_TWO); (k_hi, k_lo) =
// sigma log2(e) split into two floats, so that e = (s - mid) k is exact to ~2^-48 relative.
template <int TW>
__device__ __forceinline__ void doc_factors(float s, float mid, float k_hi, float k_lo, float w, float& fa,
                                            float& fb, float& fe, float& fg) {
  const float c = s - mid;
  const float eh = c * k_hi;
  const float el = fmaf(c, k_lo, fmaf(c, k_hi, -eh)) * kLn2;   // (c k - eh) ln 2
  fa = ex2_approx(-eh) * (1.0f - el);
  fb = ex2_approx(eh) * (1.0f + el);
  fe = eh;
  fg = w;
}

// Accumulators of one lane over a run of ring steps: packed halves (even / odd column of a pair)
// plus scalar ones for the odd last column of a 1- or 3-rank chunk.
template <int R>
struct RowAcc {
  float2 l2;
  float l1;
  float2 r2[R];
  float r1[R];
  __device__ __forceinline__ void clear() {
    l2 = make_float2(0.0f, 0.0f);
    l1 = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) { r2[r] = make_float2(0.0f, 0.0f); r1[r] = 0.0f; }
  }
  __device__ __forceinline__ float loss() const { return l2.x + l2.y + l1; }
  __device__ __forceinline__ float row(int r) const { return r2[r].x + r2[r].y + r1[r]; }
  __device__ __forceinline__ void add(const RowAcc& o) {
    l2 = vadd(l2, o.l2);
    l1 += o.l1;
#pragma unroll
    for (int r = 0; r < R; ++r) { r2[r] = vadd(r2[r], o.r2[r]); r1[r] += o.r1[r]; }
  }
};

// The R x R pairs of one row chunk against one column chunk (factored form), columns two at a time.
//   dwin : delta window of this chunk distance, dwin[k - r + R - 1] for row r and column k (TW_DELTA)
//   tc   : column gradients of this step, tc[k] += ...
template <int TW, int R>
__device__ __forceinline__ void chunk_pairs(const float (&ra)[R], const float (&re)[R], const float (&rg)[R],
                                            const float (&rv)[R], const float (&cx)[4], const float (&ce)[4],
                                            const float (&cg)[4], const float (&dwin)[8], RowAcc<R>& acc,
                                            float (&tc)[4]) {
  constexpr int CP = R / 2;
#pragma unroll
  for (int cp = 0; cp < CP; ++cp) {
    const int k = 2 * cp;
    const float2 cx2 = make_float2(cx[k], cx[k + 1]), ce2 = make_float2(ce[k], ce[k + 1]);
    const float2 cg2 = make_float2(cg[k], cg[k + 1]);
    float2 tc2 = make_float2(tc[k], tc[k + 1]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 dw2 = make_float2(0.0f, 0.0f);
      if constexpr (TW == TW_DELTA) dw2 = make_float2(dwin[k - r + R - 1], dwin[k - r + R]);
      pair_fact<TW, float2>(vsplat<float2>(ra[r]), vsplat<float2>(re[r]), vsplat<float2>(rg[r]),
                            vsplat<float2>(rv[r]), cx2, ce2, cg2, dw2, acc.l2, acc.r2[r], tc2);
    }
    tc[k] = tc2.x;
    tc[k + 1] = tc2.y;
  }
  if constexpr ((R & 1) != 0) {
    constexpr int k = R - 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float dw = 0.0f;
      if constexpr (TW == TW_DELTA) dw = dwin[k - r + R - 1];
      pair_fact<TW, float>(ra[r], re[r], rg[r], rv[r], cx[k], ce[k], cg[k], dw, acc.l1, acc.r1[r], tc[k]);
    }
  }
}

// The triangle inside a chunk (rank distance k - r > 0), scalar.
template <int TW, int R>
__device__ __forceinline__ void chunk_triangle(const float (&ra)[R], const float (&re)[R], const float (&rg)[R],
                                               const float (&rv)[R], const float (&cx)[4], const float (&ce)[4],
                                               const float (&cg)[4], const float (&dwin)[8], float& tl,
                                               float (&tr)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int k = r + 1; k < R; ++k) {
      float dw = 0.0f;
      if constexpr (TW == TW_DELTA) dw = dwin[k - r + R - 1];
      pair_fact<TW, float>(ra[r], re[r], rg[r], rv[r], cx[k], ce[k], cg[k], dw, tl, tr[r], tr[k]);
    }
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

// R column values of chunk `chunk` into a 4-slot array (the unused slots of a contiguous layout
// are never read by chunk_pairs)
template <int R, int S>
__device__ __forceinline__ void load_cols(const float* __restrict__ p, int chunk, float (&v)[4]) {
  if constexpr (S == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + chunk * 4);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    float u[R];
    load_chunk<R>(p, chunk * S, u);
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = r < R ? u[r < R ? r : 0] : 0.0f;
  }
}

// Factored form of ring_pass (below): packed pair math (chunk_pairs), same schedule.
template <int TW, int R, int S>
__device__ __forceinline__ float ring_pass_fact(const PairSoA& it, float* __restrict__ gcol,
                                                const float* __restrict__ wtab, int C, int n, int lane,
                                                float (&racc)[R]) {
  constexpr bool kFast = S == 4;
  const bool active = lane < C;
  const int me = active ? lane : 0;
  float ra[R], re[R], rg[R], rv[R];
  load_chunk_s<R, S>(it.a, me, ra);
  load_chunk_s<R, S>(it.g, me, rg);
  load_chunk_s<R, S>(it.e, me, re);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool valid = active && (me * R + r < n);
    ra[r] = valid ? ra[r] : 0.0f;
    rg[r] = valid ? rg[r] : (TW == TW_TWO ? 0.0f : kBigGain);
    rv[r] = valid ? 1.0f : 0.0f;
  }
  RowAcc<R> acc;
  acc.clear();
  float tl = 0.0f, tr[R];
#pragma unroll
  for (int r = 0; r < R; ++r) tr[r] = 0.0f;

  // within-chunk triangle (rank distance c - r > 0)
  if constexpr (R > 1) {
    float dwin[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if constexpr (TW == TW_DELTA) load_window(reinterpret_cast<const float4*>(wtab) + kMaxChunks * 2, dwin);
    float cx[4], ce[4], cg[4];
    load_cols<R, S>(it.b, me, cx);
    load_cols<R, S>(it.g, me, cg);
    load_cols<R, S>(it.e, me, ce);
    chunk_triangle<TW, R>(ra, re, rg, rv, cx, ce, cg, dwin, tl, tr);
  }

  const int steps = C >> 1;
  // even ring: the last step pairs chunk l with l + C/2 from both ends; only the lower half commits
  const bool even = (C & 1) == 0;
  const bool dup_last = even && lane >= steps;
  int pc = me;
  int m = 1;
  if constexpr (kFast) {
    // a lane without rows carries a = 0 and the padding gain: its pairs contribute exact zeros and only
    // the address of its column update is predicated (chunk 32 + lane is a dump)
    const int nfast = even ? steps - 1 : steps;
    for (; m <= nfast; ++m) {
      pc = pc + 1 == C ? 0 : pc + 1;
      float dwin[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      if constexpr (TW == TW_DELTA)
        load_window(reinterpret_cast<const float4*>(wtab) + (pc - me + kMaxChunks) * 2, dwin);
      float4* g4 = reinterpret_cast<float4*>(gcol) + (active ? pc : 32 + lane);
      const float4 gold = *g4;
      float cx[4], ce[4], cg[4];
      load_cols<R, S>(it.b, pc, cx);
      load_cols<R, S>(it.g, pc, cg);
      load_cols<R, S>(it.e, pc, ce);
      float tc[4] = {gold.x, gold.y, gold.z, gold.w};
      chunk_pairs<TW, R>(ra, re, rg, rv, cx, ce, cg, dwin, acc, tc);
      *g4 = make_float4(tc[0], tc[1], tc[2], tc[3]);
      __syncwarp();
    }
  }
  for (; m <= steps; ++m) {
    pc = pc + 1 == C ? 0 : pc + 1;
    const bool commit = active && !(dup_last && m == steps);
    float dwin[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if constexpr (TW == TW_DELTA)
      load_window(reinterpret_cast<const float4*>(wtab) + (pc - me + kMaxChunks) * 2, dwin);
    float cx[4], ce[4], cg[4];
    load_cols<R, S>(it.b, pc, cx);
    load_cols<R, S>(it.g, pc, cg);
    load_cols<R, S>(it.e, pc, ce);
    RowAcc<R> tmp;
    tmp.clear();
    float tc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    chunk_pairs<TW, R>(ra, re, rg, rv, cx, ce, cg, dwin, tmp, tc);
    if (commit) {
      acc.add(tmp);
      // column gradients: chunk pc owns gcol[4 pc .. 4 pc + 3]
#pragma unroll
      for (int c = 0; c < R; ++c) gcol[pc * 4 + c] += tc[c];
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < R; ++r) racc[r] = active ? acc.row(r) + tr[r] : 0.0f;
  return active ? acc.loss() + tl : 0.0f;
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
  if constexpr (FACTORED) return ring_pass_fact<TW, R, S>(it, gcol, wtab, C, n, lane, racc);
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
  if constexpr (FACTORED) {
    float ra[R], re[R], rg[R];
    const float rv[R] = {1.0f, 1.0f, 1.0f, 1.0f};
    load_chunk<R>(it.a, row_base + lane * R, ra);
    load_chunk<R>(it.g, row_base + lane * R, rg);
    load_chunk<R>(it.e, row_base + lane * R, re);
    RowAcc<R> acc;
    acc.clear();
    const int dist0 = (col_base - row_base) >> 2;   // chunk distance of column chunk 0 from row chunk 0
    for (int m = 0; m < 32; ++m) {
      const int pc = (lane + m) & 31;
      float dwin[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      if constexpr (TW == TW_DELTA) {
        // delta[4d - 4 .. 4d + 3]: slot e + 4, so the window (e = -3..3) starts one float in
        const float4* w4 = reinterpret_cast<const float4*>(delta) + (dist0 + pc - lane - 1);
        const float4 w0 = *w4, w1 = *(w4 + 1);
        dwin[0] = w0.y; dwin[1] = w0.z; dwin[2] = w0.w;
        dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w;
      }
      float cx[4], ce[4], cg[4];
      load_cols<R, 4>(it.b + col_base, pc, cx);
      load_cols<R, 4>(it.g + col_base, pc, cg);
      load_cols<R, 4>(it.e + col_base, pc, ce);
      float4* g4 = reinterpret_cast<float4*>(gcol) + pc;
      const float4 gold = *g4;
      float tc[4] = {gold.x, gold.y, gold.z, gold.w};
      chunk_pairs<TW, R>(ra, re, rg, rv, cx, ce, cg, dwin, acc, tc);
      *g4 = make_float4(tc[0], tc[1], tc[2], tc[3]);
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < R; ++r) racc[r] = acc.row(r);
    return acc.loss();
  } else {
    const float* colx = it.a;
    float ra[R], re[R], rg[R];
    load_chunk<R>(it.a, row_base + lane * R, ra);
    load_chunk<R>(it.g, row_base + lane * R, rg);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      re[r] = 0.0f;
      racc[r] = 0.0f;
    }
    float lacc = 0.0f;
    const int dist0 = (col_base - row_base) >> 2;
    for (int m = 0; m < 32; ++m) {
      const int pc = (lane + m) & 31;
      float dwin[8];
      if constexpr (TW == TW_DELTA) {
        const float4* w4 = reinterpret_cast<const float4*>(delta) + (dist0 + pc - lane - 1);
        const float4 w0 = *w4, w1 = *(w4 + 1);
        dwin[0] = w0.y; dwin[1] = w0.z; dwin[2] = w0.w;
        dwin[3] = w1.x; dwin[4] = w1.y; dwin[5] = w1.z; dwin[6] = w1.w; dwin[7] = 0.0f;
      }
      float cx[R], cg[R];
      load_chunk<R>(colx, col_base + pc * R, cx);
      load_chunk<R>(it.g, col_base + pc * R, cg);
      float tc[R];
#pragma unroll
      for (int c = 0; c < R; ++c) tc[c] = 0.0f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = 0; c < R; ++c) {
          float dw = 1.0f;
          if constexpr (TW == TW_DELTA) dw = dwin[c - r + R - 1];
          pair_once<TW, false>(ra[r], re[r], rg[r], cx[c], 0.0f, cg[c], dw, lacc, racc[r], tc[c]);
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
}

}  // namespace ltr
