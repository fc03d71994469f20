// svmrank_parser.cpp -- multithreaded SVMrank text ingestion: libltr_svmrank.so.
//
// Replaces the reference's single-threaded C parser pytorchltr/datasets/svmrank/parser/svmrank_parser.h
// (:174-515, a table-driven DFA over 8 KB fread chunks) and its Cython wrapper svmrank_parser.pyx:19-59.
// Same grammar, same values, same result layout (dense row-major matrix of width max_col + 1 - min_col,
// int32 labels, int64 query ids), different machine: the file is mapped (the page cache's pages, nothing
// copied), cut into one slice per thread at line boundaries, every thread
// parses its slice into private COO buffers sized once from a count of the ':' in the slice, and the dense
// matrix is then filled in parallel (float64 like the reference, or float32 -- what the GPU path consumes
// -- without a second copy).  The scanner has two tiers: the common feature token
//   ' ' digits ':' ['-'] digits ['.' digits] followed by ' ' or '\n'   (at most 8 digits per run)
// is measured with two 16-byte SSE2 compares (where do the digits stop?) and converted with the 8-digit
// SWAR multiply trick -- no data-dependent loop, so none of the ~3 branch mispredictions per token that
// cost the byte-at-a-time loop ~200 cycles per value; every other token (exponents, several signs, long
// runs, comments, '\r', errors) takes the byte-at-a-time path below, which alone decides what is an error.
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

#include <emmintrin.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

namespace {

enum : int { PARSE_OK = 0, PARSE_FILE_ERROR = 1, PARSE_FORMAT_ERROR = 2, PARSE_MEMORY_ERROR = 3 };

// growable array without value-initialisation (the COO buffers are written once, front to back)
template <typename T>
struct Buf {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  Buf() = default;
  Buf(const Buf&) = delete;
  Buf& operator=(const Buf&) = delete;
  Buf(Buf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  ~Buf() { free(p); }
  void reserve(size_t c) {
    if (c <= cap) return;
    T* q = static_cast<T*>(realloc(p, c * sizeof(T)));
    if (!q) throw std::bad_alloc();
    p = q;
    cap = c;
  }
  void push_back(T v) {
    if (n == cap) reserve(cap ? 2 * cap : 1024);
    p[n++] = v;
  }
  size_t size() const { return n; }
  const T* data() const { return p; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct Slice {
  const char* begin = nullptr;
  const char* end = nullptr;
  Buf<int32_t> ys;
  Buf<int64_t> qids;
  Buf<uint64_t> row_start;   // index of the row's first value in col / val
  Buf<uint32_t> col;
  Buf<double> val;
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

// ---- fast tier -------------------------------------------------------------------------------------------
constexpr size_t kTextPad = 64;   // readable bytes behind the text (zeros): 16-byte loads never leave the mapping

// bit i set: byte i of the 16 bytes at p is not a decimal digit; bit 16 always set
inline uint32_t nondigit_mask(const char* p) {
  const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
  const __m128i lo = _mm_cmplt_epi8(v, _mm_set1_epi8('0'));
  const __m128i hi = _mm_cmpgt_epi8(v, _mm_set1_epi8('9'));
  return static_cast<uint32_t>(_mm_movemask_epi8(_mm_or_si128(lo, hi))) | 0x10000u;
}

inline uint32_t nondigit_bits(__m128i v) {
  const __m128i lo = _mm_cmplt_epi8(v, _mm_set1_epi8('0'));
  const __m128i hi = _mm_cmpgt_epi8(v, _mm_set1_epi8('9'));
  return static_cast<uint32_t>(_mm_movemask_epi8(_mm_or_si128(lo, hi)));
}
// bit i set: byte i is a blank or a line feed
inline uint32_t sep_mask(__m128i v) {
  return static_cast<uint32_t>(_mm_movemask_epi8(
      _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8(' ')), _mm_cmpeq_epi8(v, _mm_set1_epi8('\n')))));
}

// the decimal number in the len (1..8) digits at p
inline uint64_t parse_digits(const char* p, int len) {
  uint64_t v;
  memcpy(&v, p, 8);
  // first digit to the top byte position it would have in an 8-digit number, '0' in front
  if (len < 8) v = (v << (64 - 8 * len)) | (0x3030303030303030ull >> (8 * len));
  v = (v & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;
  v = (v & 0x00FF00FF00FF00FFull) * 6553601 >> 16;
  return (v & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32;
}

const uint64_t kPow10Int[9] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull};

inline size_t count_byte(const char* p, const char* end, char c) {
  size_t n = 0;
  const __m128i needle = _mm_set1_epi8(c);
  for (; p + 16 <= end; p += 16)
    n += static_cast<size_t>(__builtin_popcount(static_cast<unsigned>(
        _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p)), needle)))));
  for (; p < end; ++p) n += *p == c;
  return n;
}

// Parses the lines of one slice (the slice ends right after a '\n' or at end of file; kTextPad readable
// bytes follow the text).
void parse_slice(Slice* s) {
  const char* p = s->begin;
  const char* const end = s->end;
  // LTR_SVMRANK_SLOW=1: byte-at-a-time tier only (tests compare the two tiers on hostile inputs)
  const char* const slow_env = getenv("LTR_SVMRANK_SLOW");
  const bool fast = !(slow_env && slow_env[0] == '1');
  try {
    {
      // every value and every "qid:" has one ':' -- an upper bound that is exact for files without comments
      const size_t colons = count_byte(p, end, ':');
      const size_t lines = count_byte(p, end, '\n') + 1;
      s->ys.reserve(lines);
      s->qids.reserve(lines);
      s->row_start.reserve(lines);
      s->col.reserve(colons + 1);
      s->val.reserve(colons + 1);
    }
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
      s->row_start.push_back(s->val.size());
      s->ys.push_back(static_cast<int32_t>(y));
      s->qids.push_back(qid);
      // features
      for (;;) {
        // fast tier: one blank, then a token  digits ':' ['-'] digits ['.' digits]  that ends at a ' ' or '\n' within
        // 32 bytes.  Only "where is the next separator" feeds the next iteration (one load, two compares, one
        // count): the digit work of a token hangs off the pointer chain, so consecutive tokens overlap.
        if (fast && *p == ' ' && p + 1 < end) {
          const char* q = p + 1;
          const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(q));
          const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(q + 16));
          const uint32_t sep = sep_mask(v0) | (sep_mask(v1) << 16);
          const uint32_t nondig = nondigit_bits(v0) | (nondigit_bits(v1) << 16);
          const int len = sep ? __builtin_ctz(sep) : 32;                       // token length
          const int nc = __builtin_ctz(nondig | 0x80000000u);                  // column digits
          if (len < 32 && q + len < end && nc >= 1 && nc <= 8 && q[nc] == ':') {
            const int neg = q[nc + 1] == '-';
            const int v = nc + 1 + neg;                                        // first digit of the value
            const uint32_t m = (nondig >> v) | (0x80000000u >> v) | 0x80000000u;
            const int n1 = __builtin_ctz(m);
            const int dot = q[v + n1] == '.';
            const int n2 = dot ? __builtin_ctz((m >> (n1 + 1)) | 0x40000000u) : 0;
            if (n1 >= 1 && n1 <= 8 && n2 <= 8 && (!dot || n2 >= 1) && v + n1 + dot + n2 == len) {
              const uint64_t col = parse_digits(q, nc);
              uint64_t mag = parse_digits(q + v, n1);
              if (dot) mag = mag * kPow10Int[n2] + parse_digits(q + v + n1 + 1, n2);
              const long sval = neg ? -static_cast<long>(mag) : static_cast<long>(mag);
              const double fv = static_cast<double>(sval) * kPow10(-n2);        // reference :395-398
              const uint32_t c32 = static_cast<uint32_t>(col);
              s->col.push_back(c32);
              s->val.push_back(fv);
              if (c32 < s->min_col) s->min_col = c32;
              if (c32 > s->max_col) s->max_col = c32;
              s->any_col = true;
              p = q + len;
              continue;
            }
          }
        }
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
  char* text = nullptr;             // the file (mapped), at least kTextPad readable zero bytes behind it
  size_t text_bytes = 0, map_bytes = 0;
  void unmap() {
    if (text) munmap(text, map_bytes);
    text = nullptr;
  }
  ~ltr_svmrank_result() { unmap(); }
  uint64_t rows = 0, cols = 0;
  uint32_t min_col = 0;
};

// 0 = ok, 1 = file error (errno set), 2 = format error, 3 = out of memory  (svmrank_parser.h:21-24)
int ltr_svmrank_parse(const char* path, int n_threads, ltr_svmrank_result** out) {
  if (!path || !out) return PARSE_FORMAT_ERROR;
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return PARSE_FILE_ERROR;
  ltr_svmrank_result* r = nullptr;
  try {
    r = new ltr_svmrank_result();
    const off_t size = lseek(fd, 0, SEEK_END);
    if (size < 0) { const int e = errno; close(fd); delete r; errno = e; return PARSE_FILE_ERROR; }
    r->text_bytes = static_cast<size_t>(size);
    if (n_threads < 1) n_threads = static_cast<int>(std::thread::hardware_concurrency());
    if (n_threads < 1) n_threads = 1;
    {
      // The page cache's own pages, mapped read-only over an anonymous region one page longer than the file: the
      // zero page behind the text keeps the scanner's 16-byte loads inside the mapping, and nothing is copied.
      const size_t page = static_cast<size_t>(sysconf(_SC_PAGESIZE));
      const size_t file_span = (r->text_bytes + page - 1) / page * page;
      r->map_bytes = file_span + page;
      void* region = mmap(nullptr, r->map_bytes, PROT_READ, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (region == MAP_FAILED) { close(fd); delete r; return PARSE_MEMORY_ERROR; }
      r->text = static_cast<char*>(region);
      if (r->text_bytes > 0) {
        void* m = mmap(region, r->text_bytes, PROT_READ, MAP_PRIVATE | MAP_FIXED, fd, 0);
        if (m != MAP_FAILED) {
          madvise(region, r->text_bytes, MADV_SEQUENTIAL);
        } else {
          // a file that cannot be mapped (some network / pseudo file systems): read it into the region instead
          if (mprotect(region, r->map_bytes, PROT_READ | PROT_WRITE) != 0) {
            const int e = errno; close(fd); delete r; errno = e; return PARSE_FILE_ERROR;
          }
          size_t off = 0;
          while (off < r->text_bytes) {
            const ssize_t got = pread(fd, r->text + off, r->text_bytes - off, static_cast<off_t>(off));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) { const int e = got < 0 ? errno : EIO; close(fd); delete r; errno = e; return PARSE_FILE_ERROR; }
            off += static_cast<size_t>(got);
          }
        }
      }
      close(fd);
    }
    const size_t min_slice = 1u << 16;
    size_t want = r->text_bytes / min_slice + 1;
    if (want > static_cast<size_t>(n_threads)) want = static_cast<size_t>(n_threads);
    const char* base = r->text;
    const char* const end = base + r->text_bytes;
    r->slices.resize(want);
    const char* cur = base;
    for (size_t i = 0; i < want; ++i) {
      const char* stop = i + 1 == want ? end : base + (r->text_bytes * (i + 1)) / want;
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
    r->unmap();                                                             // the text is no longer needed
  } catch (const std::bad_alloc&) {
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
      for (size_t row = 0; row < s.ys.size(); ++row) {
        T* out = xs + (base + row) * cols;
        const size_t k1 = row + 1 < s.ys.size() ? s.row_start[row + 1] : s.val.size();
        for (size_t k = s.row_start[row]; k < k1; ++k)   // later duplicates overwrite earlier ones, :481-483
          out[s.col[k] - r->min_col] = static_cast<T>(s.val[k]);
      }
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
    for (size_t row = 0; row < s.ys.size(); ++row) {
      indptr[base + row] = static_cast<int64_t>(nnz);
      const size_t k1 = row + 1 < s.ys.size() ? s.row_start[row + 1] : s.val.size();
      for (size_t k = s.row_start[row]; k < k1; ++k) {
        if (indices) indices[nnz] = static_cast<int64_t>(s.col[k] - r->min_col);
        if (values) values[nnz] = static_cast<float>(s.val[k]);
        ++nnz;
      }
    }
  }
  indptr[r->rows] = static_cast<int64_t>(nnz);
  return PARSE_OK;
}

void ltr_svmrank_release(ltr_svmrank_result* r) { delete r; }

}  // extern "C"
