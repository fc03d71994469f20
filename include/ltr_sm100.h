/*
 * ltr_sm100.h -- C ABI of libltr_sm100.so: the pytorchltr loss / metric hot path
 * `(scores, relevance, n) -> per-query loss | metric` as hand-written CUDA for
 * NVIDIA B200 (sm_100a).
 *
 * The reference (rjagerman/pytorchltr) is pure Python on this path; it has no FFI.
 * Each entry point below therefore replaces one Python-level interface of the
 * reference (cited file:line, relative to the reference root) and is what a ctypes
 * binding inside those modules would call -- see INTEGRATION.md for the stubs.
 *
 * Common contract
 *   - Padded batches, row-major and contiguous: scores float32 [B*L] (ld = L),
 *     relevance int64 (the reference's dtype), int32, int16 or uint8 [B*L] (`rel_bytes` = 8, 4, 2
 *     or 1), n int64 or int32 [B] (`n_bytes` = 8 or 4).  Documents j >= n[b] are padding.  n is
 *     clamped to [0, L].
 *   - Device entry points take DEVICE pointers, enqueue on `stream` (a cudaStream_t,
 *     NULL = legacy default stream), never synchronise, never allocate and keep no
 *     pointer after returning: they are CUDA-graph capturable and re-entrant.
 *   - `*_host` entry points take HOST pointers (pinned memory makes the copies
 *     asynchronous) plus a caller-owned device workspace; they enqueue
 *     H2D copies -> kernel -> D2H copies and return without synchronising: everything has
 *     completed once `stream` has.  Large batches are cut into chunks of queries whose copies run
 *     on two library-owned streams (created once per device on first use) so that H2D, kernel and
 *     D2H overlap; these entry points are not meant for CUDA-graph capture.
 *   - The first launch on a device fills ~70 KB of score-independent tables on the caller's stream
 *     and creates one event; no call ever synchronises.
 *   - The caller owns every buffer.  Inputs are read-only.
 *   - Return value: 0 on success, a negative LTR_E* code otherwise; nothing is thrown
 *     and the process is never terminated.  ltr_strerror() names the code,
 *     ltr_last_cuda_error() returns the cudaError_t behind the last LTR_ECUDA of the
 *     calling thread.
 *   - 1 <= L <= LTR_MAX_LIST_SIZE; B >= 0 (B == 0 is a no-op).
 *   - Rankings break score ties lowest-index-first and place padded documents last in
 *     index order.  (The reference breaks ties with a random permutation drawn from the
 *     global torch RNG, utils/tensor_operations.py:43-45; any tie order is a valid
 *     reference output.  This library consumes no RNG state.)
 */
#ifndef LTR_SM100_H_
#define LTR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTR_VERSION 200            /* major*100 + minor */
#define LTR_MAX_LIST_SIZE 4096

/* error codes */
#define LTR_OK 0
#define LTR_EINVAL (-1)            /* bad argument (NULL pointer, bad mode, bad dtype width, k < 0) */
#define LTR_EUNSUPPORTED (-2)      /* L > LTR_MAX_LIST_SIZE or not an sm_100 device */
#define LTR_ECUDA (-3)             /* a CUDA runtime call failed: see ltr_last_cuda_error() */

/* `mode` of ltr_pairwise_additive: loss/pairwise_additive.py */
#define LTR_ADD_HINGE 0            /* PairwiseHingeLoss      :93-113  */
#define LTR_ADD_DCG_HINGE 1        /* PairwiseDCGHingeLoss   :116-133 */
#define LTR_ADD_LOGISTIC 2         /* PairwiseLogisticLoss   :136-163 */

/* `mode` of ltr_lambda: loss/pairwise_lambda.py */
#define LTR_LAM_ARP1 0             /* LambdaARPLoss1   :95-117  */
#define LTR_LAM_ARP2 1             /* LambdaARPLoss2   :120-140 */
#define LTR_LAM_NDCG1 2            /* LambdaNDCGLoss1  :143-173 */
#define LTR_LAM_NDCG2 3            /* LambdaNDCGLoss2  :176-218 */

/* `metric` of ltr_rank_metrics */
#define LTR_METRIC_DCG 0           /* evaluation/dcg.py:41-99 */
#define LTR_METRIC_NDCG 1          /* evaluation/dcg.py:8-38  */
#define LTR_METRIC_ARP 2           /* evaluation/arp.py:7-42  */

/* loss family selector of the host-buffer entry point */
#define LTR_FAMILY_ADDITIVE 0
#define LTR_FAMILY_LAMBDA 1
#define LTR_FAMILY_LISTNET 2

int ltr_version(void);
const char *ltr_strerror(int rc);
int ltr_last_cuda_error(void);

/*
 * Replaces _PairwiseAdditiveLoss.forward (loss/pairwise_additive.py:51-90) with the
 * per-pair bodies :107-113 / :158-163 and the DCG modifier :132-133, plus the
 * backward pass autograd derives from them.
 *   loss_out    [B]   per-query loss.
 *   dscores_out [B*L] d loss_b / d scores[b, :] (0 on padding), NULL to skip.
 *   loss_sum    [1]   NULL, or a device scalar to which sum_b loss_b is added
 *                     atomically (caller zeroes it; feeds the scalar all-reduce).
 * sigma is only used by LTR_ADD_LOGISTIC.
 */
int ltr_pairwise_additive(int mode, const float *scores, const void *rel, int rel_bytes,
                          const void *n, int n_bytes, int B, int L, float sigma,
                          float *loss_out, float *dscores_out, float *loss_sum, void *stream);

/*
 * Replaces LambdaLoss.forward (loss/pairwise_lambda.py:50-92) incl. rank_by_score
 * (utils/tensor_operations.py:48-64), the gathers, the per-pair bodies (:114-117,
 * :135-140, :165-173, :198-218), _ndcg_gains (:221-228), _max_dcg (:231-241), and
 * the backward pass (scatter through the gather).
 *   ranking_out [B*L] int64, NULL to skip: the ranking used (== ltr_rank_by_score).
 */
int ltr_lambda(int mode, const float *scores, const void *rel, int rel_bytes, const void *n,
               int n_bytes, int B, int L, float sigma, float *loss_out, float *dscores_out,
               int64_t *ranking_out, float *loss_sum, void *stream);

/*
 * Same two entry points with a caller-owned SCHEDULING WORKSPACE (device memory, 16-byte
 * aligned, at least ltr_schedule_workspace_bytes(B) bytes, contents irrelevant on entry and
 * undefined on return; it must not be shared by launches that may run concurrently).
 * The cost of a query grows with n^2, so a batch mixes queries that differ 4x and more in work.
 * With the workspace the launch first sorts the queries by decreasing n (one small kernel on
 * `stream`) and the loss kernel starts the long ones first, handing the rest out dynamically:
 * same results, shorter tail.  Without it (or workspace == NULL) the queries are taken in
 * batch order.  Still no allocation, no synchronisation, CUDA-graph capturable.
 */
size_t ltr_schedule_workspace_bytes(int B);
int ltr_pairwise_additive_ws(int mode, const float *scores, const void *rel, int rel_bytes,
                             const void *n, int n_bytes, int B, int L, float sigma,
                             float *loss_out, float *dscores_out, float *loss_sum,
                             void *workspace, size_t workspace_bytes, void *stream);
int ltr_lambda_ws(int mode, const float *scores, const void *rel, int rel_bytes, const void *n,
                  int n_bytes, int B, int L, float sigma, float *loss_out, float *dscores_out,
                  int64_t *ranking_out, float *loss_sum, void *workspace, size_t workspace_bytes,
                  void *stream);

/*
 * ListNet (top-1 softmax cross entropy).  Named by the task, absent from the reference
 * snapshot: softmax(relevance) vs log_softmax(scores) over the valid documents, masking
 * idiom of utils/tensor_operations.py:81-87.  Parity unpinned.
 */
int ltr_listnet(const float *scores, const void *rel, int rel_bytes, const void *n, int n_bytes,
                int B, int L, float *loss_out, float *dscores_out, float *loss_sum, void *stream);

/*
 * Replaces dcg / ndcg (evaluation/dcg.py:41-99, :8-38) and arp (evaluation/arp.py:7-42).
 *   k > 0 : out[b * out_ld] = metric@min(k, L)   (the reference's dcg[:, :k][:, -1])
 *   k == 0: (dcg / ndcg only) all cut-offs, out[b * out_ld + r] for r < L  (k=None)
 *   exp_gain: 1 -> gain 2^rel - 1, 0 -> gain rel.  Ignored by LTR_METRIC_ARP (k too).
 * As in the reference, dcg does not mask the relevance of padded documents (:85).
 */
int ltr_rank_metrics(int metric, const float *scores, const void *rel, int rel_bytes,
                     const void *n, int n_bytes, int B, int L, int k, int exp_gain, float *out,
                     int out_ld, void *stream);

/* Replaces rank_by_score (utils/tensor_operations.py:48-64): ranking_out [B*L] int64. */
int ltr_rank_by_score(const float *scores, const void *n, int n_bytes, int B, int L,
                      int64_t *ranking_out, void *stream);

/*
 * Backward of every loss above: out[b, j] = g[b * g_stride] * dscores[b, j]  (the chain rule
 * autograd applies to the saved per-query gradient).  g_stride is 1 for a dense upstream
 * gradient and 0 for a broadcast one (what `loss.sum().backward()` / `.mean()` produce), which
 * saves materialising it.  `out` may alias `dscores`.
 */
int ltr_scale_rows(const float *g, int g_stride, const float *dscores, float *out, int B, int L,
                   void *stream);

/*
 * Fused linear scorer + ListNet (the caller side of the loss path,
 * examples/01-basic-usage.py:44,72: `loss_fn(torch.nn.Linear(F, 1)(features), relevance, n)`
 * followed by `.backward()`): reads the feature tensor ONCE and returns, per launch,
 *   scores_out  [B*L]      w . x + b                   (NULL to skip; padded rows included)
 *   loss_out    [B]        ListNet loss per query       (same arithmetic as ltr_listnet)
 *   dscores_out [B*L]      d loss_b / d scores[b, :]    (NULL to skip)
 *   qgrad_out   [B*(F+1)]  per-query parameter gradient: qgrad[b, f] = sum_l dscores[b, l] *
 *                          features[b, l, f] for f < F, qgrad[b, F] = sum_l dscores[b, l] (bias).
 * features float32 [B*L*F] row-major, weight float32 [F], bias float32 [1] or NULL.
 * When F % 4 == 0, F <= 1024, (rel_bytes * L) % 16 == 0, features / rel are 16-byte aligned and
 * 4 L F + rel_bytes L bytes fit in shared memory, the block of a query is staged by TMA (double
 * buffered) and read once; every other shape runs a tiled kernel that streams the block twice inside
 * the same launch (the second pass is served by L2).
 *
 * ltr_linear_listnet_backward: the backward pass for an upstream gradient g [B] (g_stride 1) or a
 * broadcast one (g_stride 0): dweight_out[f] = sum_b g[b] qgrad[b, f], dbias_out[0] (NULL to skip)
 * = sum_b g[b] qgrad[b, F].  No second pass over the features.  `workspace`: device memory,
 * >= ltr_linear_listnet_workspace_bytes(F) bytes (partial sums, reduced in a fixed order).
 */
size_t ltr_linear_listnet_workspace_bytes(int F);
int ltr_linear_listnet(const float *features, const float *weight, const float *bias,
                       const void *rel, int rel_bytes, const void *n, int n_bytes, int B, int L,
                       int F, float *scores_out, float *loss_out, float *dscores_out,
                       float *qgrad_out, float *loss_sum, void *stream);
int ltr_linear_listnet_backward(const float *qgrad, const float *g, int g_stride, int B, int F,
                                float *dweight_out, float *dbias_out, void *workspace,
                                size_t workspace_bytes, void *stream);

/*
 * MLP scorer: the reference's documented ranker, docs/source/getting-started.rst:42-51
 *   torch.nn.Sequential(Linear(F, H1), ReLU, Linear(H1, H2), ReLU, Linear(H2, 1))   (136 -> 50 -> 10 -> 1)
 * applied to the flat feature block in front of any loss of this header:
 *   scores_out[r] = w3 . relu(W2 relu(W1 features[r, :] + b1) + b2) + b3,  r < rows (= B*L).
 * features float32 [rows*F] row-major; w1 [H1*F], w2 [H2*H1], w3 [H2] row-major as torch.nn.Linear.weight;
 * b1 [H1], b2 [H2], b3 [1] or NULL.  Layer 1 runs on the tensor cores (tcgen05.mma kind::tf32 fed by TMA
 * tensor copies, accumulators in tensor memory): TF32 operands, float32 accumulation -- the arithmetic of
 * torch.backends.cuda.matmul.allow_tf32 = True; layers 2-3 are float32.  The features are read once, only
 * the score leaves the SM.  Requires F % 4 == 0, 16-byte aligned features / w1, H1 <= 64, H2 <= 16
 * (LTR_EUNSUPPORTED otherwise: the caller keeps its own modules for such a model).  Any width F: W1 stays in
 * shared memory beside whole 128-document tiles while that fits (F <= 288), wider rows (Yahoo's 699 -> 700)
 * stream W1 with the tile in 64-column slabs that accumulate in tensor memory.  Rows of a multiple of 32 bytes
 * (F % 8 == 0) stream about 20 % faster than F % 8 == 4.
 * hz_out (NULL to skip; what a training step passes): [rows * ltr_mlp_hz_pitch(H1, H2)] floats, the activations
 * ltr_mlp_backward can start from instead of recomputing layer 1 -- per document relu(W1 x + b1) in columns
 * [0, H1), the layer-2 pre-activation in [Z0, Z0 + H2), 1.0 in column ONE (its products are db1 / db2), zeros
 * elsewhere; (Z0, ONE, pitch) = (32, 40, 44) for H1 <= 32, H2 <= 8 and (52, 62, 64) for other H1 <= 50, H2 <= 10
 * (256 B per document, written with 128-bit stores next to the 544 B of features read).  ltr_mlp_hz_pitch is 0
 * for larger hidden layers: they have no such path.
 */
int ltr_mlp_hz_pitch(int H1, int H2);
int ltr_mlp_scores(const float *features, long long rows, int F, const float *w1, const float *b1,
                   int H1, const float *w2, const float *b2, int H2, const float *w3,
                   const float *b3, float *scores_out, float *hz_out, void *stream);

/*
 * Backward pass of ltr_mlp_scores for an upstream gradient dscores [rows] (d loss / d scores, e.g. the
 * dscores_out of any loss above after ltr_scale_rows; 0 on padded documents): a second pass over the
 * features that recomputes the hidden layers, forms dZ1 on chip and accumulates dW1 = dZ1^T X on the tensor
 * cores (tcgen05.mma kind::tf32, M = 64, the TMA-staged feature tile as the MN-major operand, the accumulator
 * resident in tensor memory for the whole launch).  grads_out [ltr_mlp_grad_len(F, H1, H2)] receives
 *   [dW1 (H1*F) | db1 (H1) | dW2 (H2*H1) | db2 (H2) | dW3 (H2) | db3 (1)]
 * in torch.nn.Linear's layouts.  dW1 multiplies TF32 operands (dZ1 and the features), everything else is
 * float32; sums are formed in a fixed order (bit-reproducible).  `workspace`: device memory, >=
 * ltr_mlp_workspace_bytes(F, H1, H2) bytes.  Same shape limits as ltr_mlp_scores; without hz one feature tile,
 * W1 and the dZ1 operand must fit in shared memory together (F up to 136; LTR_EUNSUPPORTED beyond) -- with hz
 * any width: a CTA accumulates up to 224 columns of dW1, wider rows are split into column slabs over the CTAs.
 * d loss / d features is not formed (the features are data, not a trainable module's output).
 * hz: the activation rows kept by ltr_mlp_scores (hz_out), or NULL.  With them the launch reads features and
 * activations once each (no W1, no layer 1 again), uses the forward pass's own ReLU masks (the gradient of
 * exactly the function the forward kernel computed) and forms ALL sums over documents as one tensor-core
 * product per 8 documents: [dZ1^T ; dZ2^T] (64 x 8) . [X | H1 Z2 1] gives dW1, dW2, db1, db2 in one accumulator.
 */
size_t ltr_mlp_grad_len(int F, int H1, int H2);
size_t ltr_mlp_workspace_bytes(int F, int H1, int H2);
int ltr_mlp_backward(const float *features, long long rows, int F, const float *w1, const float *b1,
                     int H1, const float *w2, const float *b2, int H2, const float *w3,
                     const float *b3, const float *hz, const float *dscores, float *grads_out,
                     void *workspace, size_t workspace_bytes, void *stream);

/*
 * Position-biased click model (click_simulation/pbm.py:12-63): for the document
 * d = rankings[b, r] at rank r,
 *   propensity_out[b, d] = 1 / (2 + r)^eta  if r < min(n[b], cutoff)  else 0    (cutoff 0 = none)
 *   click_prob_out[b, d] = relevance_probs[ys[b, d]] * propensity_out[b, d]
 * both float32 [B*L] in DOCUMENT order (the reference inverts the ranking with a second argsort).
 * rankings int64 [B*L] must hold a permutation of 0..L-1 per row (ltr_rank_by_score's output);
 * grades outside 0..n_probs-1 are clamped.  The Bernoulli draw of the clicks is left to the caller.
 */
int ltr_pbm_probabilities(const int64_t *rankings, const void *ys, int ys_bytes, const void *n,
                          int n_bytes, const float *relevance_probs, int n_probs, int cutoff,
                          float eta, int B, int L, float *click_prob_out, float *propensity_out,
                          void *stream);

/*
 * Ragged -> padded batch on the device.  Replaces the Python loop of
 * SVMRankDataset.collate_fn()._collate_fn with the default ListSampler
 * (datasets/svmrank/svmrank.py:135-205, datasets/list_sampler.py:5-19) for a dataset held in device
 * memory as features float32 [N*F], relevance int64 [N], offsets int64 [Q+1] (query q owns the
 * documents offsets[q] .. offsets[q+1]); qidx int64 [B] are the queries of the batch.
 *   feat_out [B*L*F] the first min(count, L) documents of every query, zero padded
 *   rel_out  [B*L]   int64, zero padded (svmrank.py:149-150)
 *   n_out    [B]     int64 min(count, L)                       (svmrank.py:194)
 *   count_out [B]    int64 untruncated list sizes, NULL to skip
 * L is the batch's list_size: max over the batch of min(count, max_list_size), computed by the
 * caller (it sizes the outputs).
 */
int ltr_collate(const float *features, const int64_t *relevance, const int64_t *offsets,
                const int64_t *qidx, int B, int L, int F, float *feat_out, int64_t *rel_out,
                int64_t *n_out, int64_t *count_out, void *stream);

/*
 * Sampled collation (list samplers, datasets/list_sampler.py:19-61, consulted by collate_fn only for
 * queries with more than `L` documents, svmrank.py:159-190): such a query takes the documents
 * sel[b * sel_ld + l], l < L (indices inside the query: a permutation prefix produced on the device,
 * e.g. ltr_rank_by_score over random keys for the UniformSampler); shorter queries are copied in order.
 * sel == NULL: every query in order (== ltr_collate).
 */
int ltr_collate_sampled(const float *features, const int64_t *relevance, const int64_t *offsets,
                        const int64_t *qidx, const int64_t *sel, int sel_ld, int B, int L, int F,
                        float *feat_out, int64_t *rel_out, int64_t *n_out, int64_t *count_out,
                        void *stream);

/*
 * Sparse collation (svmrank.py:163-177, 198-203): features in CSR form over the documents (indptr
 * [N+1], indices, values); returns the COO triplets (batch row, list position, feature) as coo_out
 * [3 * nnz_out] (three planes) + val_out [nnz_out] of a sparse (B, L, F) tensor, and rel_out / n_out as
 * above.  out_ptr [B * L + 1]: exclusive scan of the non-zero counts of the selected documents (the
 * caller computes it from indptr; out_ptr[B * L] == nnz_out).
 */
int ltr_collate_sparse(const int64_t *indptr, const int64_t *indices, const float *values,
                       const int64_t *relevance, const int64_t *offsets, const int64_t *qidx,
                       const int64_t *sel, int sel_ld, const int64_t *out_ptr, int B, int L,
                       int64_t nnz_out, int64_t *coo_out, float *val_out, int64_t *rel_out,
                       int64_t *n_out, void *stream);

/*
 * Scalar all-reduce over NVLink peer memory (the path's only exchange, the [sum, count]
 * behind a global mean).  Every rank of ONE node creates a mailbox (device memory, exported by CUDA IPC),
 * the 64-byte handles are exchanged by the caller (e.g. torch.distributed.all_gather_object) and
 * connected; ltr_p2p_allreduce_sum then replaces values[0..k) (device, k <= 4) by the sum over all
 * ranks with one single-CTA kernel on `stream`: remote 64-bit stores carrying {value, sequence number},
 * a spin on the own mailbox, a sum in rank order (bit-identical on every rank).  CUDA-graph capturable
 * (the sequence number lives on the device).  Every rank must call it the same number of times; a missing
 * peer trips a ~2 s timeout and sets the error flag (ltr_p2p_error) instead of hanging.  The reference has
 * no distributed code; this replaces the ncclAllReduce a torch.distributed caller would issue.
 */
typedef struct ltr_p2p ltr_p2p;
int ltr_p2p_create(int rank, int world, ltr_p2p **out, unsigned char *handle_out /* 64 bytes */);
int ltr_p2p_connect(ltr_p2p *p, const unsigned char *handles /* world x 64 bytes, rank order */);
int ltr_p2p_allreduce_sum(ltr_p2p *p, float *values, int k, void *stream);
int ltr_p2p_error(ltr_p2p *p);
void ltr_p2p_destroy(ltr_p2p *p);

/*
 * Vector form of the exchange, for the one other exchange step next to the path: the parameter gradient of a
 * data-parallel ranker (SURVEY.md 8(f) N1 with 8(e)'s query sharding; the reference has no distributed code --
 * this is the all-reduce torch's DistributedDataParallel would issue through NCCL).  Same mailbox, same
 * protocol, any grid: values[0..k) (device) is replaced by the sum over the ranks, 65536 floats per launch.
 * ltr_mlp_backward_allreduce is ltr_mlp_backward with that exchange FUSED into its final reduction: the thread
 * that sums element k over the CTAs' partial vectors pushes the sum into every peer's mailbox over NVLink and
 * adds what the peers pushed, in rank order -- grads_out leaves the launch already summed over the ranks,
 * bit-identical on every rank, one launch fewer than reduce + all-reduce and no library collective (gradients
 * longer than 65536 floats take the unfused vector exchange after the reduction).  Every rank must make the
 * call with the same model shape (a rank with rows == 0 contributes zeros).  With the loss already divided by
 * the global query count (sharded_mean_loss), the SUM over ranks is the gradient of the global mean.
 */
int ltr_p2p_allreduce_vec(ltr_p2p *p, float *values, long long k, void *stream);
int ltr_mlp_backward_allreduce(const float *features, long long rows, int F, const float *w1,
                               const float *b1, int H1, const float *w2, const float *b2, int H2,
                               const float *w3, const float *b3, const float *hz,
                               const float *dscores, float *grads_out, void *workspace,
                               size_t workspace_bytes, ltr_p2p *p2p, void *stream);

/*
 * Host-buffer form of the three loss families (the call the reference's CPU path is
 * compared with end to end): copies scores / relevance / n from host memory into `workspace`,
 * runs the fused loss + gradient kernel and copies loss_out [B] and dscores_out [B*L] (if not
 * NULL) back to host memory.  `workspace` is a device buffer of at least
 * ltr_host_workspace_bytes(B, L) bytes (sized for int64 inputs; enough for every width).
 *   ltr_loss_host      int64 relevance and n (the reference's dtypes).
 *   ltr_loss_host_ex   relevance / n of any accepted width; keep_dscores != 0 computes the gradient
 *                      into the workspace even when h_dscores_out is NULL (the usual training step:
 *                      the backward pass scales it by the upstream gradient first, see
 *                      ltr_scale_rows_host).
 */
size_t ltr_host_workspace_bytes(int B, int L);
/* Byte offset, inside that workspace, of the DEVICE copy of dscores [B*L] left behind by
 * ltr_loss_host / ltr_loss_host_ex (valid once the stream has reached it). */
size_t ltr_host_workspace_dscores_offset(int B, int L);
int ltr_loss_host(int family, int mode, const float *h_scores, const int64_t *h_rel,
                  const int64_t *h_n, int B, int L, float sigma, float *h_loss_out,
                  float *h_dscores_out, void *workspace, size_t workspace_bytes, void *stream);
int ltr_loss_host_ex(int family, int mode, const float *h_scores, const void *h_rel, int rel_bytes,
                     const void *h_n, int n_bytes, int B, int L, float sigma, float *h_loss_out,
                     float *h_dscores_out, int keep_dscores, void *workspace,
                     size_t workspace_bytes, void *stream);
/*
 * Backward pass of a host caller whose upstream gradient is one scalar for every query (what
 * `loss.sum().backward()` / `loss.mean().backward()` produce): h_out[b, j] = g * d_dscores[b, j],
 * scaled on the device (into d_scratch [B*L], which may be NULL when g == 1) and copied to host
 * memory chunk by chunk, the copy of one chunk overlapping the scaling of the next.
 */
int ltr_scale_rows_host(float g, const float *d_dscores, float *h_out, int B, int L,
                        float *d_scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LTR_SM100_H_ */
