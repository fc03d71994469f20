// svmrank_parser.cpp -- multithreaded SVMrank text ingestion: libltr_svmrank.so.
//
// Replaces the reference's single-threaded C parser pytorchltr/datasets/svmrank/parser/svmrank_parser.h
// (:174-515, a table-driven DFA over 8 KB fread chunks) and its Cython wrapper svmrank_parser.pyx:19-59.
// Same grammar, same values, same result layout (dense row-major matrix of width max_col + 1 - min_col,
// int32 labels, int64 query ids), different machine: the file is read once into memory, cut into one
// slice per thread at line boundaries, every thread parses its slice into private COO buffers with a
// hand-written scanner (no per-byte table lookups), and the dense matrix is then filled in parallel
// (float64 like the reference, or float32 -- what the GPU path consumes -- without a second copy).
//
// Grammar (reference DFA, :80-131): line := ' '* label ' '+ "qid:" digits (' '+ col ':' value)* ' '*
// ['#' comment] '\n'; a line that starts with '#' is a comment; '\r' before '\n' is accepted after a
// token; value := '-'* digits ['.' digits [('e' | 'E') ['+' | '-'] digits]].  Anything else (an empty
// line, a letter in a number, an exponent without a fraction) is a format error, as in the reference.
// Values are sign * integer_digits * pow(10, exponent - decimals) in double, the reference's arithmetic
// (:395-398), so the parsed doubles are bit-identical.
//
// Deliberate differences: a line that ends right after its qid (no features) still records the qid (the
// reference stores a qid only when a space follows it, :88 / :123, which leaves its qid array one short
// of its row count); a last value that is not followed by a newline keeps its sign (the reference drops
// it at end of file, :432).
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace {

enum : int { PARSE_OK = 0, PARSE_FILE_ERROR = 1, PARSE_FORMAT_ERROR = 2, PARSE_MEMORY_ERROR = 3 };

struct Slice {
  const char* begin = nullptr;
  const char* end = nullptr;
  std::vector<int32_t> ys;
  std::vector<int64_t> qids;
  std::vector<uint32_t> row;     // row inside the slice
  std::vector<uint32_t> col;
  std::vector<double> val;
  uint32_t min_col = 0xffffffffu, max_col = 0;
  bool any_col = false;
  int status = PARSE_OK;
};

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// pow(10, e) for the exponents that occur in practice, computed once by the same libm call the reference
// makes per value (:397), so the products are bit-identical.
struct Pow10Table {
  static constexpr int kRange = 64;
  double v[2 * kRange + 1];
  Pow10Table() { for (int e = -kRange; e <= kRange; ++e) v[e + kRange] = pow(10, static_cast<double>(e)); }
  double operator()(long e) const {
    return (e >= -kRange && e <= kRange) ? v[e + kRange] : pow(10, static_cast<double>(e));
  }
};
const Pow10Table kPow10;

// Parses the lines of one slice (the slice ends right after a '\n' or at end of file).
void parse_slice(Slice* s) {
  const char* p = s->begin;
  const char* const end = s->end;
  try {
    while (p < end) {
      while (p < end && *p == ' ') ++p;                       // START_Y: leading blanks
      if (p >= end) break;
      if (*p == '#') {                                        // comment line
        while (p < end && *p != '\n') ++p;
        if (p < end) ++p;
        continue;
      }
      if (!is_digit(*p)) { s->status = PARSE_FORMAT_ERROR; return; }
      int64_t y = 0;
      while (p < end && is_digit(*p)) y = y * 10 + (*p++ - '0');
      if (p >= end || *p != ' ') { s->status = PARSE_FORMAT_ERROR; return; }
      while (p < end && *p == ' ') ++p;
      if (end - p < 5 || memcmp(p, "qid:", 4) != 0 || !is_digit(p[4])) { s->status = PARSE_FORMAT_ERROR; return; }
      p += 4;
      int64_t qid = 0;
      while (p < end && is_digit(*p)) qid = qid * 10 + (*p++ - '0');
      const uint32_t row = static_cast<uint32_t>(s->ys.size());
      s->ys.push_back(static_cast<int32_t>(y));
      s->qids.push_back(qid);
      // features
      for (;;) {
        bool blank = false;
        while (p < end && *p == ' ') { ++p; blank = true; }
        if (p >= end) break;
        const char c = *p;
        if (c == '\n') { ++p; break; }
        if (c == '#' || c == '\r') {                          // SKIP to the end of the line
          while (p < end && *p != '\n') ++p;
          if (p < end) ++p;
          break;
        }
        if (!blank || !is_digit(c)) { s->status = PARSE_FORMAT_ERROR; return; }
        uint64_t col = 0;
        while (p < end && is_digit(*p)) col = col * 10 + static_cast<uint64_t>(*p++ - '0');
        if (p >= end || *p != ':' || col > 0xfffffffeull) { s->status = PARSE_FORMAT_ERROR; return; }
        ++p;
        long sign = 1;
        while (p < end && *p == '-') { sign = -1; ++p; }
        if (p >= end || !is_digit(*p)) { s->status = PARSE_FORMAT_ERROR; return; }
        long val = 0, decplaces = 0, expval = 0, expsign = 1;
        while (p < end && is_digit(*p)) val = val * 10 + (*p++ - '0');
        if (p < end && *p == '.') {
          ++p;
          if (p >= end || !is_digit(*p)) { s->status = PARSE_FORMAT_ERROR; return; }
          while (p < end && is_digit(*p)) { val = val * 10 + (*p++ - '0'); ++decplaces; }
          if (p < end && (*p == 'e' || *p == 'E')) {
            ++p;
            // 'e' ['+' | '-'] digits*: after a sign the reference accepts an empty exponent (:119-126)
            bool exp_sign = false;
            if (p < end && (*p == '-' || *p == '+')) { if (*p == '-') expsign = -1; ++p; exp_sign = true; }
            if (!exp_sign && (p >= end || !is_digit(*p))) { s->status = PARSE_FORMAT_ERROR; return; }
            while (p < end && is_digit(*p)) expval = expval * 10 + (*p++ - '0');
          }
        }
        // the token must be followed by a blank, a comment, a line end or the end of the file
        if (p < end && *p != ' ' && *p != '#' && *p != '\r' && *p != '\n') { s->status = PARSE_FORMAT_ERROR; return; }
        double fv = static_cast<double>(sign * val);
        fv = fv * kPow10(expval * expsign - decplaces);                             // reference :395-398
        s->row.push_back(row);
        s->col.push_back(static_cast<uint32_t>(col));
        s->val.push_back(fv);
        const uint32_t c32 = static_cast<uint32_t>(col);
        if (c32 < s->min_col) s->min_col = c32;
        if (c32 > s->max_col) s->max_col = c32;
        s->any_col = true;
      }
    }
  } catch (const std::bad_alloc&) {
    s->status = PARSE_MEMORY_ERROR;
  }
}

}  // namespace

extern "C" {

struct ltr_svmrank_result {
  std::vector<Slice> slices;
  std::vector<uint64_t> row_base;   // first global row of every slice
  std::vector<char> text;
  uint64_t rows = 0, cols = 0;
  uint32_t min_col = 0;
};

// 0 = ok, 1 = file error (errno set), 2 = format error, 3 = out of memory  (svmrank_parser.h:21-24)
int ltr_svmrank_parse(const char* path, int n_threads, ltr_svmrank_result** out) {
  if (!path || !out) return PARSE_FORMAT_ERROR;
  *out = nullptr;
  FILE* fp = fopen(path, "rb");
  if (!fp) return PARSE_FILE_ERROR;
  ltr_svmrank_result* r = nullptr;
  try {
    r = new ltr_svmrank_result();
    if (fseek(fp, 0, SEEK_END) != 0) { fclose(fp); delete r; return PARSE_FILE_ERROR; }
    const long size = ftell(fp);
    if (size < 0 || fseek(fp, 0, SEEK_SET) != 0) { fclose(fp); delete r; return PARSE_FILE_ERROR; }
    r->text.resize(static_cast<size_t>(size));
    if (size > 0 && fread(r->text.data(), 1, static_cast<size_t>(size), fp) != static_cast<size_t>(size)) {
      fclose(fp);
      delete r;
      return PARSE_FILE_ERROR;
    }
    fclose(fp);
    fp = nullptr;
    if (n_threads < 1) n_threads = static_cast<int>(std::thread::hardware_concurrency());
    if (n_threads < 1) n_threads = 1;
    const size_t min_slice = 1u << 16;
    size_t want = r->text.size() / min_slice + 1;
    if (want > static_cast<size_t>(n_threads)) want = static_cast<size_t>(n_threads);
    const char* base = r->text.data();
    const char* const end = base + r->text.size();
    r->slices.resize(want);
    const char* cur = base;
    for (size_t i = 0; i < want; ++i) {
      const char* stop = i + 1 == want ? end : base + (r->text.size() * (i + 1)) / want;
      if (stop < cur) stop = cur;
      while (stop < end && stop > base && stop[-1] != '\n') ++stop;      // cut right after a newline
      r->slices[i].begin = cur;
      r->slices[i].end = stop;
      cur = stop;
    }
    std::vector<std::thread> pool;
    for (size_t i = 1; i < want; ++i) pool.emplace_back(parse_slice, &r->slices[i]);
    parse_slice(&r->slices[0]);
    for (auto& t : pool) t.join();
    uint32_t min_col = 0xffffffffu, max_col = 0;
    bool any = false;
    r->row_base.resize(want + 1);
    uint64_t rows = 0;
    for (size_t i = 0; i < want; ++i) {
      const Slice& s = r->slices[i];
      if (s.status != PARSE_OK) {
        const int st = s.status;
        delete r;
        return st;
      }
      r->row_base[i] = rows;
      rows += s.ys.size();
      if (s.any_col) {
        any = true;
        min_col = std::min(min_col, s.min_col);
        max_col = std::max(max_col, s.max_col);
      }
    }
    r->row_base[want] = rows;
    r->rows = rows;
    r->min_col = any ? min_col : 0;
    r->cols = any ? static_cast<uint64_t>(max_col) + 1 - min_col : 0;     // nr_cols - min_col, :478
    std::vector<char>().swap(r->text);                                      // the text is no longer needed
  } catch (const std::bad_alloc&) {
    if (fp) fclose(fp);
    delete r;
    return PARSE_MEMORY_ERROR;
  }
  *out = r;
  return PARSE_OK;
}

uint64_t ltr_svmrank_rows(const ltr_svmrank_result* r) { return r ? r->rows : 0; }
uint64_t ltr_svmrank_cols(const ltr_svmrank_result* r) { return r ? r->cols : 0; }
uint64_t ltr_svmrank_nnz(const ltr_svmrank_result* r) {
  uint64_t n = 0;
  if (r) for (const Slice& s : r->slices) n += s.val.size();
  return n;
}

}  // extern "C"

namespace {
template <typename T>
int fill_impl(const ltr_svmrank_result* r, T* xs, int32_t* ys, int64_t* qids, int n_threads) {
  if (!r) return PARSE_FORMAT_ERROR;
  const size_t n = r->slices.size();
  auto work = [&](size_t i) {
    const Slice& s = r->slices[i];
    const uint64_t base = r->row_base[i];
    if (ys) memcpy(ys + base, s.ys.data(), s.ys.size() * sizeof(int32_t));
    if (qids) memcpy(qids + base, s.qids.data(), s.qids.size() * sizeof(int64_t));
    if (xs) {
      const uint64_t cols = r->cols;
      memset(xs + base * cols, 0, s.ys.size() * cols * sizeof(T));
      for (size_t k = 0; k < s.val.size(); ++k)       // later duplicates overwrite earlier ones, :481-483
        xs[(base + s.row[k]) * cols + (s.col[k] - r->min_col)] = static_cast<T>(s.val[k]);
    }
  };
  if (n_threads == 1 || n <= 1) {
    for (size_t i = 0; i < n; ++i) work(i);
  } else {
    std::vector<std::thread> pool;
    for (size_t i = 1; i < n; ++i) pool.emplace_back(work, i);
    work(0);
    for (auto& t : pool) t.join();
  }
  return PARSE_OK;
}
}  // namespace

extern "C" {

// xs: rows * cols values (any content on entry), ys: rows, qids: rows; NULL skips an output.
int ltr_svmrank_fill_f64(const ltr_svmrank_result* r, double* xs, int32_t* ys, int64_t* qids, int n_threads) {
  return fill_impl<double>(r, xs, ys, qids, n_threads);
}
int ltr_svmrank_fill_f32(const ltr_svmrank_result* r, float* xs, int32_t* ys, int64_t* qids, int n_threads) {
  return fill_impl<float>(r, xs, ys, qids, n_threads);
}

// CSR export (the sparse=True datasets): indptr rows + 1, indices / values nnz (columns minus min_col, in
// file order inside a row).
int ltr_svmrank_fill_csr(const ltr_svmrank_result* r, int64_t* indptr, int64_t* indices, float* values) {
  if (!r || !indptr) return PARSE_FORMAT_ERROR;
  uint64_t nnz = 0;
  for (size_t i = 0; i < r->slices.size(); ++i) {
    const Slice& s = r->slices[i];
    const uint64_t base = r->row_base[i];
    size_t k = 0;
    for (size_t row = 0; row < s.ys.size(); ++row) {
      indptr[base + row] = static_cast<int64_t>(nnz);
      while (k < s.val.size() && s.row[k] == row) {
        if (indices) indices[nnz] = static_cast<int64_t>(s.col[k] - r->min_col);
        if (values) values[nnz] = static_cast<float>(s.val[k]);
        ++nnz;
        ++k;
      }
    }
  }
  indptr[r->rows] = static_cast<int64_t>(nnz);
  return PARSE_OK;
}

void ltr_svmrank_release(ltr_svmrank_result* r) { delete r; }

}  // extern "C"
