/*
 * ltr_svmrank.h -- C ABI of libltr_svmrank.so: multithreaded SVMrank text ingestion (host code, no CUDA).
 *
 * Replaces the reference's parser entry point
 *   int parse_svmrank_file(char* path, double** xs, shape* xs_shape, int** ys, long** qids)
 * (pytorchltr/datasets/svmrank/parser/svmrank_parser.h:174-515, bound by svmrank_parser.pyx:9-17): same
 * grammar, same values (bit-identical doubles), same error codes.  The result is held by the library
 * until the caller has copied it into its own arrays, so no ownership of malloc'd buffers crosses the
 * boundary.
 */
#ifndef LTR_SVMRANK_H_
#define LTR_SVMRANK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTR_SVMRANK_OK 0
#define LTR_SVMRANK_FILE_ERROR 1     /* PARSE_FILE_ERROR   (errno is set)        svmrank_parser.h:22 */
#define LTR_SVMRANK_FORMAT_ERROR 2   /* PARSE_FORMAT_ERROR                       svmrank_parser.h:23 */
#define LTR_SVMRANK_MEMORY_ERROR 3   /* PARSE_MEMORY_ERROR                       svmrank_parser.h:24 */

typedef struct ltr_svmrank_result ltr_svmrank_result;

/* Parses `path` with `n_threads` threads (< 1: one per hardware thread). */
int ltr_svmrank_parse(const char *path, int n_threads, ltr_svmrank_result **out);
uint64_t ltr_svmrank_rows(const ltr_svmrank_result *r);   /* documents                         */
uint64_t ltr_svmrank_cols(const ltr_svmrank_result *r);   /* max column + 1 - min column, :478 */
uint64_t ltr_svmrank_nnz(const ltr_svmrank_result *r);    /* col:value tokens                  */
/* Dense row-major matrix rows x cols (zero filled here), labels int32 [rows], qids int64 [rows]; a NULL
 * output is skipped. */
int ltr_svmrank_fill_f64(const ltr_svmrank_result *r, double *xs, int32_t *ys, int64_t *qids, int n_threads);
int ltr_svmrank_fill_f32(const ltr_svmrank_result *r, float *xs, int32_t *ys, int64_t *qids, int n_threads);
/* CSR over the documents: indptr [rows + 1], indices / values [nnz] (columns minus the minimum column). */
int ltr_svmrank_fill_csr(const ltr_svmrank_result *r, int64_t *indptr, int64_t *indices, float *values);
void ltr_svmrank_release(ltr_svmrank_result *r);

#ifdef __cplusplus
}
#endif
#endif /* LTR_SVMRANK_H_ */
