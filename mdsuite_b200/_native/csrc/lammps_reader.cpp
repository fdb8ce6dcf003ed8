// Native tabular-text trajectory reader behind include/mds_io.h (LAMMPS text dumps and extended
// XYZ files): mmap, multi-threaded configuration index (newline scan), multi-threaded tokenise +
// exact decimal -> float32((double) token).
//
// Follows the reference's reader semantics (mdsuite/file_io/lammps_trajectory_files.py:103-146,
// mdsuite/file_io/tabular_text_files.py:160-220): 9 header lines per configuration, n_atoms data
// lines split on whitespace, atoms ordered by the numeric value of the `id` column
// (utils/meta_functions.py:519-527), every kept token converted string -> float64 -> float32.
// Extended XYZ (mdsuite/file_io/extxyz_files.py:93-121): 2 header lines per configuration (atom
// count, key=value line), atoms kept in file order; the key=value line is handed to the caller
// (mds_dump_header_line), who owns the Lattice / Properties / time conventions.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "mds_io.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

}  // namespace

namespace mds_io {
// the library's thread-local error message, for the other translation units (h5_chunks.cpp)
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace mds_io

namespace {

constexpr int N_HEADER_LAMMPS = 9;
constexpr int N_HEADER_XYZ = 2;

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                         1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Correctly rounded decimal -> double for the token [p, e).  Fast path (Clinger): a mantissa below
// 2^53 and a power of ten up to 10^22 are both exact doubles, so one IEEE multiply or divide is
// the correctly rounded result.  Everything else (long mantissas, huge exponents, nan/inf, hex)
// goes through strtod on a NUL-terminated copy.
double parse_double(const char* p, const char* e) {
  const char* s = p;
  bool neg = false;
  if (s < e && (*s == '+' || *s == '-')) { neg = (*s == '-'); ++s; }
  uint64_t m = 0;
  int digits = 0, frac = 0;
  bool any = false, ok = true;
  const char* q = s;
  for (; q < e && *q >= '0' && *q <= '9'; ++q) {
    any = true;
    if (m || *q != '0') { if (digits < 19) { m = m * 10 + (uint64_t)(*q - '0'); ++digits; } else ok = false; }
  }
  if (q < e && *q == '.') {
    ++q;
    for (; q < e && *q >= '0' && *q <= '9'; ++q) {
      any = true;
      if (m || *q != '0') {
        if (digits < 19) { m = m * 10 + (uint64_t)(*q - '0'); ++digits; ++frac; } else ok = false;
      } else {
        ++frac;  // leading zero of the fraction
      }
    }
  }
  int ex = 0;
  if (any && q < e && (*q == 'e' || *q == 'E')) {
    const char* r = q + 1;
    bool eneg = false;
    if (r < e && (*r == '+' || *r == '-')) { eneg = (*r == '-'); ++r; }
    if (r < e && *r >= '0' && *r <= '9') {
      int v = 0;
      for (; r < e && *r >= '0' && *r <= '9'; ++r) v = v < 100000 ? v * 10 + (*r - '0') : v;
      ex = eneg ? -v : v;
      q = r;
    }
  }
  if (any && ok && q == e && m < (1ull << 53)) {
    const int p10 = ex - frac;
    if (m == 0) return neg ? -0.0 : 0.0;
    if (p10 >= -22 && p10 <= 22) {
      const double d = (double)m;
      const double v = p10 < 0 ? d / kPow10[-p10] : d * kPow10[p10];
      return neg ? -v : v;
    }
  }
  char buf[96];
  const size_t n = std::min<size_t>((size_t)(e - p), sizeof(buf) - 1);
  memcpy(buf, p, n);
  buf[n] = 0;
  char* endp = nullptr;
  const double v = strtod(buf, &endp);
  if (endp == buf) return std::nan("");
  return v;
}

// Fused form for the hot loop: parse the number starting at p (p < end, *p not a space), return
// the pointer one past its token.  Same two paths as parse_double.
inline const char* parse_token(const char* p, const char* end, double* out) {
  const char* s = p;
  bool neg = false;
  if (*s == '+' || *s == '-') { neg = (*s == '-'); ++s; }
  uint64_t m = 0;
  int digits = 0, frac = 0;
  const char* q = s;
  const char* d0 = q;
  while (q < end && (unsigned)(*q - '0') <= 9u) {
    if (m | (uint64_t)(*q - '0')) { m = m * 10 + (uint64_t)(*q - '0'); ++digits; }
    ++q;
  }
  bool any = q > d0;
  if (q < end && *q == '.') {
    ++q;
    const char* f0 = q;
    while (q < end && (unsigned)(*q - '0') <= 9u) {
      if (m | (uint64_t)(*q - '0')) { m = m * 10 + (uint64_t)(*q - '0'); ++digits; }
      ++q;
    }
    frac = (int)(q - f0);
    any = any || q > f0;
  }
  int ex = 0;
  if (any && q < end && (*q == 'e' || *q == 'E')) {
    const char* r = q + 1;
    bool eneg = false;
    if (r < end && (*r == '+' || *r == '-')) { eneg = (*r == '-'); ++r; }
    if (r < end && (unsigned)(*r - '0') <= 9u) {
      int v = 0;
      for (; r < end && (unsigned)(*r - '0') <= 9u; ++r) v = v < 100000 ? v * 10 + (*r - '0') : v;
      ex = eneg ? -v : v;
      q = r;
    }
  }
  const bool at_end = q >= end || is_space(*q) || *q == '\n';
  // digits <= 19 keeps m from overflowing; m < 2^53 makes (double)m exact
  if (any && at_end && digits <= 19 && m < (1ull << 53)) {
    const int p10 = ex - frac;
    if (m == 0) { *out = neg ? -0.0 : 0.0; return q; }
    if (p10 >= -22 && p10 <= 22) {
      const double d = (double)m;
      const double v = p10 < 0 ? d / kPow10[-p10] : d * kPow10[p10];
      *out = neg ? -v : v;
      return q;
    }
  }
  const char* e = p;
  while (e < end && !is_space(*e) && *e != '\n') ++e;
  *out = parse_double(p, e);
  return e;
}

inline const char* next_line(const char* p, const char* end) {
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  return nl ? nl + 1 : end;
}

// tokens of the line [p, e): fills tok_b / tok_e up to max_tok, returns the count found
inline int split_line(const char* p, const char* e, const char** tok_b, const char** tok_e,
                      int max_tok) {
  int n = 0;
  while (p < e) {
    while (p < e && is_space(*p)) ++p;
    if (p >= e || *p == '\n') break;
    const char* b = p;
    while (p < e && !is_space(*p) && *p != '\n') ++p;
    if (n < max_tok) { tok_b[n] = b; tok_e[n] = p; }
    ++n;
  }
  return n;
}

}  // namespace

struct mds_dump {
  int fd = -1;
  const char* data = nullptr;
  size_t size = 0;
  int n_threads = 1;
  int n_header = N_HEADER_LAMMPS;  // header lines in front of the atom lines of every configuration
  mds_dump_info info;
  std::vector<std::string> columns;
  std::vector<size_t> frame_off;  // byte offset of every configuration, + file size at the end
};

namespace {

void parallel_for(int n_threads, int64_t n, const std::function<void(int64_t, int64_t, int)>& fn) {
  if (n <= 0) return;
  const int t = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
  if (t == 1) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t)t);
  for (int k = 0; k < t; ++k) {
    const int64_t lo = n * k / t, hi = n * (k + 1) / t;
    th.emplace_back([=, &fn]() { fn(lo, hi, k); });
  }
  for (auto& x : th) x.join();
}

std::string line_at(const char* p, const char* end) {
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  const char* e = nl ? nl : end;
  while (e > p && (e[-1] == '\r' || e[-1] == ' ')) --e;
  return std::string(p, e);
}

int read_header(mds_dump* h) {
  const char* p = h->data;
  const char* end = h->data + h->size;
  std::string lines[N_HEADER_LAMMPS];
  for (int i = 0; i < N_HEADER_LAMMPS; ++i) {
    if (p >= end) return fail(MDS_IO_ERR_FORMAT, "file ends inside the first header");
    lines[i] = line_at(p, end);
    p = next_line(p, end);
  }
  if (lines[0].rfind("ITEM: TIMESTEP", 0) != 0)
    return fail(MDS_IO_ERR_FORMAT, "first line is not 'ITEM: TIMESTEP'");
  h->info.timestep0 = atoll(lines[1].c_str());
  h->info.n_atoms = atoll(lines[3].c_str());
  if (h->info.n_atoms < 0) return fail(MDS_IO_ERR_FORMAT, "negative atom count in the header");
  for (int k = 0; k < 3; ++k) {
    const std::string& l = lines[5 + k];
    const char* tb[4];
    const char* te[4];
    const int n = split_line(l.data(), l.data() + l.size(), tb, te, 4);
    if (n < 2) return fail(MDS_IO_ERR_FORMAT, "box bounds line %d has fewer than 2 numbers", 6 + k);
    h->info.box_lo[k] = parse_double(tb[0], te[0]);
    h->info.box_hi[k] = parse_double(tb[1], te[1]);
  }
  {
    const std::string& l = lines[8];
    std::vector<const char*> tb(256), te(256);
    const int n = split_line(l.data(), l.data() + l.size(), tb.data(), te.data(), 256);
    if (n < 3 || std::string(tb[0], te[0]) != "ITEM:" || std::string(tb[1], te[1]) != "ATOMS")
      return fail(MDS_IO_ERR_FORMAT, "header line 9 is not 'ITEM: ATOMS <columns>'");
    for (int i = 2; i < std::min(n, 256); ++i) h->columns.emplace_back(tb[i], te[i]);
  }
  h->info.n_columns = (int32_t)h->columns.size();
  h->info.id_column = -1;
  h->info.species_column = -1;
  int type_col = -1;
  for (int i = 0; i < (int)h->columns.size(); ++i) {
    if (h->columns[i] == "id") h->info.id_column = i;
    if (h->columns[i] == "element" && h->info.species_column < 0) h->info.species_column = i;
    if (h->columns[i] == "type" && type_col < 0) type_col = i;
  }
  if (h->info.species_column < 0) h->info.species_column = type_col;
  return MDS_IO_OK;
}

// extended XYZ: line 1 is the atom count, line 2 the key=value comment; the number of columns is
// the token count of the first atom line (they carry no names in the file)
int read_header_xyz(mds_dump* h) {
  const char* p = h->data;
  const char* end = h->data + h->size;
  const std::string first = line_at(p, end);
  {
    const char* tb[2];
    const char* te[2];
    const int n = split_line(first.data(), first.data() + first.size(), tb, te, 2);
    char* endp = nullptr;
    const long long v = n == 1 ? strtoll(first.c_str(), &endp, 10) : -1;
    if (n != 1 || v < 0 || (endp && *endp && !is_space(*endp)))
      return fail(MDS_IO_ERR_FORMAT, "first line of an extended XYZ file must be the atom count, got '%s'",
                  first.substr(0, 60).c_str());
    h->info.n_atoms = v;
  }
  p = next_line(p, end);
  if (p >= end) return fail(MDS_IO_ERR_FORMAT, "file ends inside the first header");
  p = next_line(p, end);
  h->info.n_columns = 0;
  if (h->info.n_atoms > 0) {
    if (p >= end) return fail(MDS_IO_ERR_FORMAT, "file ends before the first atom line");
    const std::string l = line_at(p, end);
    h->info.n_columns = split_line(l.data(), l.data() + l.size(), nullptr, nullptr, 0);
    if (h->info.n_columns < 1) return fail(MDS_IO_ERR_FORMAT, "the first atom line is empty");
  }
  for (int i = 0; i < h->info.n_columns; ++i) h->columns.emplace_back("c" + std::to_string(i));
  h->info.id_column = -1;
  h->info.species_column = -1;
  return MDS_IO_OK;
}

// configuration offsets: count newlines per slab in parallel, then place every
// (n_atoms + n_header)-th line start
int index_frames(mds_dump* h) {
  const size_t size = h->size;
  const char* data = h->data;
  const int64_t lpf = h->info.n_atoms + h->n_header;
  const int slabs = std::max(1, std::min<int>(h->n_threads * 4, (int)(size / (1 << 20)) + 1));
  std::vector<size_t> lo((size_t)slabs + 1);
  for (int s = 0; s <= slabs; ++s) lo[(size_t)s] = size * (size_t)s / (size_t)slabs;
  std::vector<int64_t> count((size_t)slabs, 0);
  parallel_for(h->n_threads, slabs, [&](int64_t a, int64_t b, int) {
    for (int64_t s = a; s < b; ++s) {
      const char* p = data + lo[(size_t)s];
      const char* e = data + lo[(size_t)s + 1];
      int64_t c = 0;
      while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        if (!nl) break;
        ++c;
        p = nl + 1;
      }
      count[(size_t)s] = c;
    }
  });
  int64_t total = std::accumulate(count.begin(), count.end(), (int64_t)0);
  if (size > 0 && data[size - 1] != '\n') ++total;  // last line without a newline
  if (lpf <= 0 || total % lpf != 0)
    return fail(MDS_IO_ERR_FORMAT, "%lld lines is not a whole number of configurations of %lld lines",
                (long long)total, (long long)lpf);
  const int64_t n_frames = total / lpf;
  h->info.n_configurations = n_frames;
  h->frame_off.assign((size_t)n_frames + 1, size);
  std::vector<int64_t> first_line((size_t)slabs + 1, 0);
  for (int s = 0; s < slabs; ++s) first_line[(size_t)s + 1] = first_line[(size_t)s] + count[(size_t)s];
  if (n_frames > 0) h->frame_off[0] = 0;
  parallel_for(h->n_threads, slabs, [&](int64_t a, int64_t b, int) {
    for (int64_t s = a; s < b; ++s) {
      const char* p = data + lo[(size_t)s];
      const char* e = data + lo[(size_t)s + 1];
      int64_t line = first_line[(size_t)s];  // newlines seen before this slab
      while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        if (!nl) break;
        ++line;  // the byte after nl starts line number `line`
        if (line % lpf == 0 && line / lpf < n_frames)
          h->frame_off[(size_t)(line / lpf)] = (size_t)(nl + 1 - data);
        p = nl + 1;
      }
    }
  });
  return MDS_IO_OK;
}

struct RowRef {
  double id;
  int32_t row;
};

// parse one configuration: selected columns of every atom line into tmp[row][c] (+ ids)
int parse_frame(const mds_dump* h, int64_t frame, const int32_t* columns, int n_cols,
                bool sort_by_id, const int32_t* dest, float* out, std::vector<float>& tmp,
                std::vector<RowRef>& order, std::string& err) {
  const int64_t n = h->info.n_atoms;
  const char* p = h->data + h->frame_off[(size_t)frame];
  const char* end = h->data + h->frame_off[(size_t)frame + 1];
  for (int i = 0; i < h->n_header; ++i) p = next_line(p, end);
  const int n_file_cols = h->info.n_columns;
  const int idc = h->info.id_column;
  const bool staged = sort_by_id || dest != nullptr;  // rows do not land in file order
  float* dst = staged ? tmp.data() : out;
  bool dense_ids = sort_by_id;
  // want[c] = output slot of file column c (+1), 0 if the column is skipped; several output slots
  // may name the same file column
  std::vector<int32_t> want((size_t)n_file_cols, 0);
  bool simple = true;
  for (int c = 0; c < n_cols; ++c) {
    if (want[(size_t)columns[c]]) simple = false;  // duplicate column: generic path below
    want[(size_t)columns[c]] = c + 1;
  }
  std::vector<const char*> tb((size_t)n_file_cols), te((size_t)n_file_cols);
  for (int64_t r = 0; r < n; ++r) {
    if (p >= end) { err = "configuration ends early"; return MDS_IO_ERR_FORMAT; }
    int nt = 0;
    double id = 0.0;
    if (simple) {
      // one pass over the line: parse the wanted columns in place, skip the others
      float* row = dst + r * n_cols;
      for (;;) {
        while (p < end && is_space(*p)) ++p;
        if (p >= end) break;
        if (*p == '\n') { ++p; break; }
        if (nt < n_file_cols && (want[(size_t)nt] || nt == idc)) {
          double v;
          p = parse_token(p, end, &v);
          if (want[(size_t)nt]) row[want[(size_t)nt] - 1] = (float)v;
          if (nt == idc) id = v;
        } else {
          while (p < end && !is_space(*p) && *p != '\n') ++p;
        }
        ++nt;
      }
    } else {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
      const char* le = nl ? nl : end;
      nt = split_line(p, le, tb.data(), te.data(), n_file_cols);
      if (nt == n_file_cols) {
        for (int c = 0; c < n_cols; ++c)
          dst[r * n_cols + c] = (float)parse_double(tb[(size_t)columns[c]], te[(size_t)columns[c]]);
        if (sort_by_id) id = parse_double(tb[(size_t)idc], te[(size_t)idc]);
      }
      p = nl ? nl + 1 : end;
    }
    if (nt != n_file_cols) {
      char b[160];
      snprintf(b, sizeof(b), "configuration %lld, atom line %lld has %d columns, expected %d",
               (long long)frame, (long long)r, nt, n_file_cols);
      err = b;
      return MDS_IO_ERR_FORMAT;
    }
    if (sort_by_id) {
      order[(size_t)r] = {id, (int32_t)r};
      if (dense_ids && !(id >= 1.0 && id <= (double)n && id == std::floor(id))) dense_ids = false;
    }
  }
  if (!staged) return MDS_IO_OK;
  const size_t row_bytes = sizeof(float) * (size_t)n_cols;
  if (!sort_by_id) {  // file order, remapped rows
    for (int64_t r = 0; r < n; ++r)
      memcpy(out + (size_t)dest[r] * (size_t)n_cols, tmp.data() + (size_t)r * (size_t)n_cols, row_bytes);
    return MDS_IO_OK;
  }
  if (dense_ids) {
    // ids are a subset of 1..n: if they are a permutation, rank = id - 1 (checked by a bitmap)
    std::vector<char> seen((size_t)n, 0);
    bool perm = true;
    for (int64_t r = 0; r < n && perm; ++r) {
      const size_t k = (size_t)order[(size_t)r].id - 1;
      if (seen[k]) perm = false;
      seen[k] = 1;
    }
    if (perm) {
      for (int64_t r = 0; r < n; ++r) {
        const size_t k = (size_t)order[(size_t)r].id - 1;
        const size_t o = dest ? (size_t)dest[k] : k;
        memcpy(out + o * (size_t)n_cols, tmp.data() + (size_t)r * (size_t)n_cols, row_bytes);
      }
      return MDS_IO_OK;
    }
  }
  std::stable_sort(order.begin(), order.end(),
                   [](const RowRef& a, const RowRef& b) { return a.id < b.id; });
  for (int64_t k = 0; k < n; ++k) {
    const size_t o = dest ? (size_t)dest[k] : (size_t)k;
    memcpy(out + o * (size_t)n_cols, tmp.data() + (size_t)order[(size_t)k].row * (size_t)n_cols,
           row_bytes);
  }
  return MDS_IO_OK;
}

}  // namespace

extern "C" {

const char* mds_io_last_error(void) { return g_err.c_str(); }
int mds_io_abi_version(void) { return MDS_IO_ABI_VERSION; }

static int open_common(mds_dump_t** out, const char* path, int n_threads, bool xyz) {
  if (!out || !path) return fail(MDS_IO_ERR_INVALID, "mds_dump_open: NULL argument");
  *out = nullptr;
  mds_dump* h = new (std::nothrow) mds_dump();
  if (!h) return fail(MDS_IO_ERR_NOMEM, "mds_dump_open: out of memory");
  memset(&h->info, 0, sizeof(h->info));
  h->info.struct_size = sizeof(mds_dump_info);
  h->n_header = xyz ? N_HEADER_XYZ : N_HEADER_LAMMPS;
  h->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  h->fd = open(path, O_RDONLY);
  if (h->fd < 0) {
    const int rc = fail(MDS_IO_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    delete h;
    return rc;
  }
  struct stat st;
  if (fstat(h->fd, &st) != 0 || st.st_size <= 0) {
    const int rc = fail(MDS_IO_ERR_IO, "%s is empty or cannot be stat'ed", path);
    close(h->fd);
    delete h;
    return rc;
  }
  h->size = (size_t)st.st_size;
  h->info.file_bytes = (int64_t)h->size;
  void* m = mmap(nullptr, h->size, PROT_READ, MAP_PRIVATE, h->fd, 0);
  if (m == MAP_FAILED) {
    const int rc = fail(MDS_IO_ERR_IO, "mmap of %s failed: %s", path, strerror(errno));
    close(h->fd);
    delete h;
    return rc;
  }
  madvise(m, h->size, MADV_SEQUENTIAL);
  h->data = (const char*)m;
  int rc = xyz ? read_header_xyz(h) : read_header(h);
  if (rc == MDS_IO_OK) rc = index_frames(h);
  if (rc == MDS_IO_OK && !xyz) {
    h->info.timestep1 = h->info.timestep0;
    if (h->info.n_configurations > 1) {
      const char* p = h->data + h->frame_off[1];
      const char* end = h->data + h->size;
      p = next_line(p, end);
      h->info.timestep1 = atoll(line_at(p, end).c_str());
    }
  }
  if (rc != MDS_IO_OK) {
    mds_dump_close(h);
    return rc;
  }
  *out = h;
  return MDS_IO_OK;
}

int mds_dump_open(mds_dump_t** out, const char* path, int n_threads) {
  return open_common(out, path, n_threads, false);
}

int mds_xyz_open(mds_dump_t** out, const char* path, int n_threads) {
  return open_common(out, path, n_threads, true);
}

int64_t mds_dump_header_line(mds_dump_t* h, int64_t frame, int line, char* buf, int64_t buf_len) {
  if (!h || !buf || buf_len < 1) return fail(MDS_IO_ERR_INVALID, "mds_dump_header_line: bad argument");
  if (frame < 0 || frame >= h->info.n_configurations)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_header_line: configuration %lld of %lld",
                (long long)frame, (long long)h->info.n_configurations);
  if (line < 0 || line >= h->n_header)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_header_line: line %d of %d", line, h->n_header);
  const char* p = h->data + h->frame_off[(size_t)frame];
  const char* end = h->data + h->frame_off[(size_t)frame + 1];
  for (int i = 0; i < line; ++i) p = next_line(p, end);
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  const char* e = nl ? nl : end;
  if (e > p && e[-1] == '\r') --e;
  const int64_t len = (int64_t)(e - p);
  const int64_t n = std::min<int64_t>(len, buf_len - 1);
  memcpy(buf, p, (size_t)n);
  buf[n] = 0;
  return len;
}

void mds_dump_close(mds_dump_t* h) {
  if (!h) return;
  if (h->data) munmap((void*)h->data, h->size);
  if (h->fd >= 0) close(h->fd);
  delete h;
}

int mds_dump_info_get(mds_dump_t* h, mds_dump_info* out) {
  if (!h || !out) return fail(MDS_IO_ERR_INVALID, "mds_dump_info_get: NULL argument");
  if (out->struct_size != sizeof(mds_dump_info))
    return fail(MDS_IO_ERR_INVALID, "mds_dump_info_get: struct_size %u != %zu", out->struct_size,
                sizeof(mds_dump_info));
  *out = h->info;
  return MDS_IO_OK;
}

int mds_dump_column_name(mds_dump_t* h, int idx, char* buf, int buf_len) {
  if (!h || !buf || buf_len < 1) return fail(MDS_IO_ERR_INVALID, "mds_dump_column_name: bad argument");
  if (idx < 0 || idx >= (int)h->columns.size())
    return fail(MDS_IO_ERR_INVALID, "mds_dump_column_name: column %d of %zu", idx, h->columns.size());
  snprintf(buf, (size_t)buf_len, "%s", h->columns[(size_t)idx].c_str());
  return MDS_IO_OK;
}

int mds_dump_first_frame_labels(mds_dump_t* h, int column, int sort_by_id, char* buf,
                                int label_stride) {
  if (!h || !buf || label_stride < 2)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_first_frame_labels: bad argument");
  if (column < 0 || column >= h->info.n_columns)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_first_frame_labels: column %d", column);
  if (sort_by_id && h->info.id_column < 0)
    return fail(MDS_IO_ERR_FORMAT, "the dump has no 'id' column to sort by");
  if (h->info.n_configurations < 1) return MDS_IO_OK;
  const int64_t n = h->info.n_atoms;
  const char* p = h->data + h->frame_off[0];
  const char* end = h->data + h->frame_off[1];
  for (int i = 0; i < h->n_header; ++i) p = next_line(p, end);
  const int nc = h->info.n_columns;
  std::vector<const char*> tb((size_t)nc), te((size_t)nc);
  std::vector<RowRef> order((size_t)n);
  std::vector<std::string> labels((size_t)n);
  for (int64_t r = 0; r < n; ++r) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const int nt = split_line(p, le, tb.data(), te.data(), nc);
    if (nt != nc)
      return fail(MDS_IO_ERR_FORMAT, "first configuration, atom line %lld has %d columns, expected %d",
                  (long long)r, nt, nc);
    labels[(size_t)r].assign(tb[(size_t)column], te[(size_t)column]);
    order[(size_t)r] = {sort_by_id ? parse_double(tb[(size_t)h->info.id_column],
                                                 te[(size_t)h->info.id_column])
                                   : (double)r,
                        (int32_t)r};
    p = nl ? nl + 1 : end;
  }
  std::stable_sort(order.begin(), order.end(),
                   [](const RowRef& a, const RowRef& b) { return a.id < b.id; });
  for (int64_t k = 0; k < n; ++k)
    snprintf(buf + (size_t)k * (size_t)label_stride, (size_t)label_stride, "%s",
             labels[(size_t)order[(size_t)k].row].c_str());
  return MDS_IO_OK;
}

int mds_dump_read_columns(mds_dump_t* h, int64_t frame0, int64_t n_frames, const int32_t* columns,
                          int32_t n_cols, int sort_by_id, float* out) {
  return mds_dump_read_columns_mapped(h, frame0, n_frames, columns, n_cols, sort_by_id, nullptr, out);
}

int mds_dump_read_columns_mapped(mds_dump_t* h, int64_t frame0, int64_t n_frames,
                                 const int32_t* columns, int32_t n_cols, int sort_by_id,
                                 const int32_t* dest, float* out) {
  if (!h || !columns || !out || n_cols < 1)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_read_columns: bad argument");
  if (dest) {
    std::vector<char> seen((size_t)h->info.n_atoms, 0);
    for (int64_t k = 0; k < h->info.n_atoms; ++k) {
      if (dest[k] < 0 || dest[k] >= h->info.n_atoms || seen[(size_t)dest[k]])
        return fail(MDS_IO_ERR_INVALID, "mds_dump_read_columns_mapped: dest is not a permutation");
      seen[(size_t)dest[k]] = 1;
    }
  }
  if (frame0 < 0 || n_frames < 0 || frame0 + n_frames > h->info.n_configurations)
    return fail(MDS_IO_ERR_INVALID, "mds_dump_read_columns: configurations [%lld, %lld) of %lld",
                (long long)frame0, (long long)(frame0 + n_frames),
                (long long)h->info.n_configurations);
  for (int c = 0; c < n_cols; ++c)
    if (columns[c] < 0 || columns[c] >= h->info.n_columns)
      return fail(MDS_IO_ERR_INVALID, "mds_dump_read_columns: column %d of %d", columns[c],
                  h->info.n_columns);
  if (sort_by_id && h->info.id_column < 0)
    return fail(MDS_IO_ERR_FORMAT, "the dump has no 'id' column to sort by");
  const int64_t n = h->info.n_atoms;
  std::atomic<int> status(MDS_IO_OK);
  std::string first_err;
  std::atomic<bool> have_err(false);
  std::atomic<int64_t> next(0);
  const int t = (int)std::max<int64_t>(1, std::min<int64_t>(h->n_threads, n_frames));
  auto worker = [&]() {
    std::vector<float> tmp((sort_by_id || dest) ? (size_t)n * (size_t)n_cols : 0);
    std::vector<RowRef> order(sort_by_id ? (size_t)n : 0);
    std::string err;
    for (;;) {
      const int64_t f = next.fetch_add(1);
      if (f >= n_frames || status.load() != MDS_IO_OK) break;
      const int rc = parse_frame(h, frame0 + f, columns, n_cols, sort_by_id != 0, dest,
                                 out + (size_t)f * (size_t)n * (size_t)n_cols, tmp, order, err);
      if (rc != MDS_IO_OK) {
        bool expected = false;
        if (have_err.compare_exchange_strong(expected, true)) first_err = err;
        status.store(rc);
        break;
      }
    }
  };
  if (t == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < t; ++k) th.emplace_back(worker);
    for (auto& x : th) x.join();
  }
  if (status.load() != MDS_IO_OK) return fail(status.load(), "%s", first_err.c_str());
  return MDS_IO_OK;
}

int mds_parse_float32(const char* tokens, int64_t n, int32_t stride, float* out) {
  if (!tokens || !out || stride < 1) return fail(MDS_IO_ERR_INVALID, "mds_parse_float32: bad argument");
  for (int64_t i = 0; i < n; ++i) {
    const char* p = tokens + (size_t)i * (size_t)stride;
    const char* e = p + strnlen(p, (size_t)stride);
    while (p < e && is_space(*p)) ++p;
    while (e > p && is_space(e[-1])) --e;
    if (p >= e) { out[i] = std::nanf(""); continue; }
    // the hot loop's fused scanner (which falls back to parse_double); both must agree
    double a = 0.0;
    parse_token(p, e, &a);
    const double b = parse_double(p, e);
    if (!(a == b || (a != a && b != b)) || std::signbit(a) != std::signbit(b))
      return fail(MDS_IO_ERR_FORMAT, "internal: the two number parsers disagree on '%.*s'",
                  (int)(e - p), p);
    out[i] = (float)a;
  }
  return MDS_IO_OK;
}

}  // extern "C"
