/*
 * ltr_oracle.c -- CPU restatement of the pytorchltr loss / metric hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may
 * load it; nothing under pytorchltr_b200/ imports, links or calls it.
 *
 * Parity status: PINNED for every function except ltr_oracle_listnet.  The
 * restatement is checked (tests/test_oracle_*.py) against
 *   - the known-answer vectors of the reference's own tests
 *     (tests/loss/test_pairwise_additive.py, tests/loss/test_pairwise_lambda.py,
 *      tests/evaluation/test_dcg.py, tests/evaluation/test_arp.py, docs doctests), and
 *   - fixtures produced by running the unmodified reference (float64 scores, CPU
 *     autograd) on seeded random batches: the tests/golden .npz files, made by
 *     tests/golden/make_golden.py.
 * ListNet does not exist in the reference snapshot -> "parity unpinned" for it;
 * its oracle is the builder's own float64 statement of Cao et al. 2007 top-1 ListNet.
 *
 * Conventions (reference call signature `(scores, relevance, n)`):
 *   scores  float32 [B*L] row-major, relevance int64 [B*L], n int64 [B].
 *   Pair math is done in double on the float32 score values, i.e. what the
 *   reference computes when called with `scores.double()` ("ref64"); the
 *   score-independent weights (gains, discounts, delta) are float32 exactly as
 *   in the reference.  With f32 != 0 the pair math itself is done in float32
 *   with the reference's operation order ("ref32"; used for hinge, whose
 *   gradient is integer-valued and must match bit for bit).
 *   Outputs are double.  Ties in a ranking are broken lowest-index-first (the
 *   reference breaks them randomly, tensor_operations.py:43-45, so any order is
 *   a valid reference output); padded documents follow in index order.
 *
 * Every function is a plain loop over queries (OpenMP over b when built with
 * -fopenmp); there is no tiling, no vectorisation trickery and no shared state.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ADD_HINGE = 0, ADD_DCG_HINGE = 1, ADD_LOGISTIC = 2 };
enum { LAM_ARP1 = 0, LAM_ARP2 = 1, LAM_NDCG1 = 2, LAM_NDCG2 = 3 };

static int g_threads = 0; /* 0 = OpenMP default */

int ltr_oracle_version(void) { return 1; }

int ltr_oracle_set_threads(int t)
{
    g_threads = t;
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#endif
    return 0;
}

int ltr_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int64_t clamp_n(int64_t n, int L)
{
    if (n < 0) return 0;
    if (n > L) return L;
    return n;
}

/* ------------------------------------------------------------------------- */
/* rank_by_score: utils/tensor_operations.py:48-64 = mask_padded_values(-inf)  */
/* (:6-26) then a per-row descending argsort (:29-45).  Valid documents by    */
/* descending score, padded documents last.                                   */
/* ------------------------------------------------------------------------- */
typedef struct { double key; int64_t idx; } keyidx_t;

static int cmp_desc(const void *a, const void *b)
{
    const keyidx_t *x = (const keyidx_t *)a, *y = (const keyidx_t *)b;
    /* NaN sorts first in a descending torch.argsort (NaN is "largest"). */
    int xn = isnan(x->key), yn = isnan(y->key);
    if (xn != yn) return xn ? -1 : 1;
    if (!xn) {
        if (x->key > y->key) return -1;
        if (x->key < y->key) return 1;
    }
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* ranking[0..L): valid docs (idx < n) sorted by key desc, then idx n..L-1. */
static void rank_row_d(const double *key, int64_t n, int L, int64_t *ranking, keyidx_t *tmp)
{
    for (int64_t j = 0; j < n; ++j) { tmp[j].key = key[j]; tmp[j].idx = j; }
    qsort(tmp, (size_t)n, sizeof(keyidx_t), cmp_desc);
    for (int64_t j = 0; j < n; ++j) ranking[j] = tmp[j].idx;
    for (int64_t j = n; j < L; ++j) ranking[j] = j;
}

int ltr_oracle_rank_by_score(const float *scores, const int64_t *n, int B, int L,
                             int64_t *ranking)
{
#pragma omp parallel
    {
        keyidx_t *tmp = (keyidx_t *)malloc(sizeof(keyidx_t) * (size_t)(L > 0 ? L : 1));
        double *key = (double *)malloc(sizeof(double) * (size_t)(L > 0 ? L : 1));
#pragma omp for schedule(dynamic, 16)
        for (int b = 0; b < B; ++b) {
            for (int j = 0; j < L; ++j) key[j] = (double)scores[(size_t)b * L + j];
            rank_row_d(key, clamp_n(n[b], L), L, ranking + (size_t)b * L, tmp);
        }
        free(tmp); free(key);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Pairwise additive losses: loss/pairwise_additive.py:51-90 (template),       */
/* :107-113 (hinge), :132-133 (DCG-hinge modifier), :158-163 (logistic).       */
/* ------------------------------------------------------------------------- */
int ltr_oracle_pairwise_additive(int mode, int f32, const float *scores, const int64_t *rel,
                                 const int64_t *n, int B, int L, double sigma,
                                 double *loss, double *grad)
{
    if (mode < ADD_HINGE || mode > ADD_LOGISTIC) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (int b = 0; b < B; ++b) {
        const float *s = scores + (size_t)b * L;
        const int64_t *y = rel + (size_t)b * L;
        double *g = grad ? grad + (size_t)b * L : NULL;
        const int64_t nb = clamp_n(n[b], L);
        double acc = 0.0;
        if (g) for (int j = 0; j < L; ++j) g[j] = 0.0;
        /* pairs with max(i, j) >= n are zeroed (pairwise_additive.py:75-81) */
        for (int64_t i = 0; i < nb; ++i) {
            for (int64_t j = 0; j < nb; ++j) {
                if (y[i] - y[j] <= 0) continue;      /* :111 / :162 rel_pair_diffs <= 0 */
                if (mode == ADD_LOGISTIC) {
                    /* :161 log2(1 + exp(-sigma * (s_i - s_j))) */
                    double l, dl;
                    if (f32) {
                        float d = s[i] - s[j];
                        float e = expf(-(float)sigma * d);
                        l = (double)log2f(1.0f + e);
                        dl = -(double)((float)sigma * (e / (1.0f + e))) / M_LN2;
                    } else {
                        double d = (double)s[i] - (double)s[j];
                        double e = exp(-sigma * d);
                        l = log2(1.0 + e);
                        dl = -sigma * (e / (1.0 + e)) / M_LN2;
                    }
                    acc += l;
                    if (g) { g[i] += dl; g[j] -= dl; }
                } else {
                    /* :109-112 loss = 1.0 - (s_i - s_j); loss[loss < 0] = 0 */
                    double l;
                    if (f32) {
                        float d = s[i] - s[j];        /* first float32 rounding  */
                        float lf = 1.0f - d;          /* second float32 rounding */
                        if (lf < 0.0f) continue;
                        l = (double)lf;
                    } else {
                        l = 1.0 - ((double)s[i] - (double)s[j]);
                        if (l < 0.0) continue;
                    }
                    acc += l;                          /* kink l == 0 stays active: d/ds = -1 */
                    if (g) { g[i] -= 1.0; g[j] += 1.0; }
                }
            }
        }
        if (mode == ADD_DCG_HINGE) {
            /* :132-133 -1 / ln(2 + h)  (natural log) */
            double lg = log(2.0 + acc);
            double scale = 1.0 / ((2.0 + acc) * lg * lg);
            if (g) for (int j = 0; j < L; ++j) g[j] *= scale;
            acc = -1.0 / lg;
        }
        loss[b] = acc;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* LambdaLoss family: loss/pairwise_lambda.py:50-92 (template), :114-117       */
/* (ARP1), :135-140 (ARP2), :165-173 (NDCG1), :198-218 (NDCG2), :221-228       */
/* (_ndcg_gains), :231-241 (_max_dcg).                                         */
/* ------------------------------------------------------------------------- */

/* _max_dcg (:231-241): relevance sorted descending (padding last), float32
 * gains (2^rel - 1) / float32 discounts log2(2 + r), summed.  The reference sums
 * in float32 in ATen's order; here the float32 terms are added in double and the
 * result rounded to float32 once (within 1 ulp of any float32 summation order). */
static float max_dcg_row(const int64_t *y_ranked, int64_t nb, int L, keyidx_t *tmp,
                         double *key, int64_t *rk)
{
    for (int j = 0; j < L; ++j) key[j] = (double)y_ranked[j];
    rank_row_d(key, nb, L, rk, tmp);
    double acc = 0.0;
    for (int64_t r = 0; r < nb; ++r) {
        float gain = (float)(ldexp(1.0, (int)y_ranked[rk[r]]) - 1.0);
        float disc = log2f(2.0f + (float)r);
        acc += (double)(gain / disc);
    }
    return (float)acc;
}

int ltr_oracle_lambda(int mode, int f32, const float *scores, const int64_t *rel,
                      const int64_t *n, int B, int L, double sigma,
                      double *loss, double *grad, int64_t *ranking_out)
{
    if (mode < LAM_ARP1 || mode > LAM_NDCG2) return -1;
    (void)f32; /* the lambda family is always evaluated as ref64 */
    /* delta table of NDCG2 (:206-211): float32, delta[k] = |1/D(k) - 1/D(k+1)|,
     * D(k) = log2(2 + k). */
    float *delta = (float *)malloc(sizeof(float) * (size_t)(L + 1));
    float *disc = (float *)malloc(sizeof(float) * (size_t)(L + 2));
    for (int k = 0; k <= L; ++k) disc[k] = log2f(2.0f + (float)k);
    for (int k = 0; k < L; ++k) delta[k] = fabsf(1.0f / disc[k] - 1.0f / disc[k + 1]);

#pragma omp parallel
    {
        size_t LL = (size_t)(L > 0 ? L : 1);
        keyidx_t *tmp = (keyidx_t *)malloc(sizeof(keyidx_t) * LL);
        double *key = (double *)malloc(sizeof(double) * LL);
        int64_t *rk = (int64_t *)malloc(sizeof(int64_t) * LL);
        int64_t *rk2 = (int64_t *)malloc(sizeof(int64_t) * LL);
        double *ss = (double *)malloc(sizeof(double) * LL);   /* scores in rank order */
        int64_t *ys = (int64_t *)malloc(sizeof(int64_t) * LL); /* relevance in rank order */
        float *gain = (float *)malloc(sizeof(float) * LL);
        double *gs = (double *)malloc(sizeof(double) * LL);   /* grad in rank order */
#pragma omp for schedule(dynamic, 8)
        for (int b = 0; b < B; ++b) {
            const float *s = scores + (size_t)b * L;
            const int64_t *y = rel + (size_t)b * L;
            const int64_t nb = clamp_n(n[b], L);
            /* :66-70 ranking = rank_by_score; gather scores and relevance */
            for (int j = 0; j < L; ++j) key[j] = (double)s[j];
            rank_row_d(key, nb, L, rk, tmp);
            for (int j = 0; j < L; ++j) { ss[j] = (double)s[rk[j]]; ys[j] = y[rk[j]]; gs[j] = 0.0; }
            if (ranking_out) memcpy(ranking_out + (size_t)b * L, rk, sizeof(int64_t) * (size_t)L);

            if (mode == LAM_NDCG1 || mode == LAM_NDCG2) {
                /* _ndcg_gains :221-228: (2^rel - 1) / max_dcg, max_dcg == 0 -> 1 */
                float md = max_dcg_row(ys, nb, L, tmp, key, rk2);
                if (md == 0.0f) md = 1.0f;
                for (int j = 0; j < L; ++j)
                    gain[j] = (float)(ldexp(1.0, (int)ys[j]) - 1.0) / md;
            }

            double acc = 0.0;
            /* pairs with max(i, j) >= n are zeroed (:80-86); i, j are rank positions */
            for (int64_t i = 0; i < nb; ++i) {
                for (int64_t j = 0; j < nb; ++j) {
                    double w;  /* exponent of the sigmoid == weight of log2(1 + e^-sd) */
                    switch (mode) {
                    case LAM_ARP1:                       /* :117 sigmoid ** rel_i */
                        w = (double)(float)ys[i];
                        break;
                    case LAM_ARP2:                       /* :137-140 */
                        if (ys[i] - ys[j] <= 0) continue;
                        w = (double)(ys[i] - ys[j]);
                        break;
                    case LAM_NDCG1:                      /* :168-173 gains / discounts[i] */
                        w = (double)(gain[i] / disc[i]);
                        break;
                    default: {                           /* :201-218 */
                        if (ys[i] - ys[j] <= 0) continue;
                        int64_t k = i > j ? i - j : j - i;
                        w = (double)(delta[k] * fabsf(gain[i] - gain[j]));
                        break;
                    }
                    }
                    /* sigmoid = 1 / (1 + exp(-sigma * (s_i - s_j)));
                     * loss_ij = -log2(sigmoid ** w)  (== w * log2(1 + e^{-sigma d})) */
                    double d = ss[i] - ss[j];
                    double e = exp(-sigma * d);
                    double sig = 1.0 / (1.0 + e);
                    if (mode == LAM_ARP2)
                        acc += w * log2(1.0 + e);        /* :138 written directly this way */
                    else
                        acc += -log2(pow(sig, w));
                    /* d loss_ij / d s_i = -sigma * w * (1 - sig) / ln 2, and minus that for s_j */
                    double lam = -sigma * w * (e * sig) / M_LN2;
                    gs[i] += lam; gs[j] -= lam;
                }
            }
            loss[b] = acc;
            if (grad) {
                double *g = grad + (size_t)b * L;
                for (int j = 0; j < L; ++j) g[j] = 0.0;
                /* backward of the gather (:69): scatter to the original doc index */
                for (int j = 0; j < L; ++j) g[rk[j]] += gs[j];
            }
        }
        free(tmp); free(key); free(rk); free(rk2); free(ss); free(ys); free(gain); free(gs);
    }
    free(delta); free(disc);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* ListNet (top-1, Cao et al. 2007).  NOT IN THE REFERENCE: parity unpinned.   */
/* P = softmax(relevance) and Q = log_softmax(scores) over the valid docs      */
/* (masking idiom of utils/tensor_operations.py:81-87); loss = -sum P_i Q_i;   */
/* grad = softmax(scores) - P; n = 0 -> 0.                                    */
/* ------------------------------------------------------------------------- */
int ltr_oracle_listnet(const float *scores, const int64_t *rel, const int64_t *n,
                       int B, int L, double *loss, double *grad)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < B; ++b) {
        const float *s = scores + (size_t)b * L;
        const int64_t *y = rel + (size_t)b * L;
        double *g = grad ? grad + (size_t)b * L : NULL;
        const int64_t nb = clamp_n(n[b], L);
        if (g) for (int j = 0; j < L; ++j) g[j] = 0.0;
        if (nb == 0) { loss[b] = 0.0; continue; }
        double ms = -INFINITY, my = -INFINITY;
        for (int64_t j = 0; j < nb; ++j) {
            if ((double)s[j] > ms) ms = (double)s[j];
            if ((double)y[j] > my) my = (double)y[j];
        }
        double zs = 0.0, zy = 0.0;
        for (int64_t j = 0; j < nb; ++j) { zs += exp((double)s[j] - ms); zy += exp((double)y[j] - my); }
        double lzs = log(zs), acc = 0.0;
        for (int64_t j = 0; j < nb; ++j) {
            double p = exp((double)y[j] - my) / zy;
            double q = ((double)s[j] - ms) - lzs;
            acc -= p * q;
            if (g) g[j] = exp(q) - p;
        }
        loss[b] = acc;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* dcg / ndcg: evaluation/dcg.py:41-99 and :8-38.                              */
/* k <= 0 means k=None (all cut-offs, out is [B*L]); else out is [B] holding   */
/* dcg[:, :k][:, -1] = DCG@min(k, L).  Padded relevance is NOT masked (:85).   */
/* ------------------------------------------------------------------------- */
static void dcg_row(const double *key, const int64_t *y, int64_t nb, int L, int exp_gain,
                    int64_t *rk, keyidx_t *tmp, double *cum)
{
    rank_row_d(key, nb, L, rk, tmp);
    double acc = 0.0;
    for (int r = 0; r < L; ++r) {
        float rs = (float)y[rk[r]];                              /* :85 .float() */
        if (exp_gain) rs = powf(2.0f, rs) - 1.0f;                /* :92 */
        float per = rs / log2f((float)r + 2.0f);                 /* :93 */
        acc += (double)per;                                      /* :94 cumsum */
        cum[r] = acc;
    }
}

int ltr_oracle_dcg(const float *scores, const int64_t *rel, const int64_t *n, int B, int L,
                   int k, int exp_gain, int normalized, double *out)
{
    if (L <= 0) return -1;
#pragma omp parallel
    {
        keyidx_t *tmp = (keyidx_t *)malloc(sizeof(keyidx_t) * (size_t)L);
        double *key = (double *)malloc(sizeof(double) * (size_t)L);
        int64_t *rk = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
        double *cum = (double *)malloc(sizeof(double) * (size_t)L);
        double *icum = (double *)malloc(sizeof(double) * (size_t)L);
#pragma omp for schedule(dynamic, 32)
        for (int b = 0; b < B; ++b) {
            const int64_t *y = rel + (size_t)b * L;
            const int64_t nb = clamp_n(n[b], L);
            for (int j = 0; j < L; ++j) key[j] = (double)scores[(size_t)b * L + j];
            dcg_row(key, y, nb, L, exp_gain, rk, tmp, cum);
            if (normalized) {
                /* ndcg :36-38: idcg = dcg(relevance.float(), ...); idcg == 0 -> 1 */
                for (int j = 0; j < L; ++j) key[j] = (double)(float)y[j];
                dcg_row(key, y, nb, L, exp_gain, rk, tmp, icum);
            }
            if (k > 0) {
                int kk = k < L ? k : L;
                double v = cum[kk - 1];
                if (normalized) { double d = icum[kk - 1]; if (d == 0.0) d = 1.0; v /= d; }
                out[b] = v;
            } else {
                for (int r = 0; r < L; ++r) {
                    double v = cum[r];
                    if (normalized) { double d = icum[r]; if (d == 0.0) d = 1.0; v /= d; }
                    out[(size_t)b * L + r] = v;
                }
            }
        }
        free(tmp); free(key); free(rk); free(cum); free(icum);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* arp: evaluation/arp.py:7-42.  sum_r (r+1) * rel_sort[r] / sum_r rel_sort[r]  */
/* over valid ranks (padded masked to 0, :38); denominator 0 -> 1 (:41).       */
/* ------------------------------------------------------------------------- */
int ltr_oracle_arp(const float *scores, const int64_t *rel, const int64_t *n, int B, int L,
                   double *out)
{
#pragma omp parallel
    {
        size_t LL = (size_t)(L > 0 ? L : 1);
        keyidx_t *tmp = (keyidx_t *)malloc(sizeof(keyidx_t) * LL);
        double *key = (double *)malloc(sizeof(double) * LL);
        int64_t *rk = (int64_t *)malloc(sizeof(int64_t) * LL);
#pragma omp for schedule(dynamic, 32)
        for (int b = 0; b < B; ++b) {
            const int64_t *y = rel + (size_t)b * L;
            const int64_t nb = clamp_n(n[b], L);
            for (int j = 0; j < L; ++j) key[j] = (double)scores[(size_t)b * L + j];
            rank_row_d(key, nb, L, rk, tmp);
            double srp = 0.0, nrp = 0.0;
            for (int64_t r = 0; r < nb; ++r) {
                double rs = (double)(float)y[rk[r]];
                srp += (double)(r + 1) * rs;
                nrp += rs;
            }
            if (nrp == 0.0) nrp = 1.0;
            out[b] = srp / nrp;
        }
        free(tmp); free(key); free(rk);
    }
    return 0;
}
