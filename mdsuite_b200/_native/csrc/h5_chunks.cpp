// Chunk decoding for the HDF5 frame feed (include/mds_io.h, "mds_h5_*" and "mds_inflate"):
// what stands between a reference project's database.hdf5 and the GPU.
//
// The reference stores every species' Positions as an atom-major (n_atoms, n_configurations, 3)
// float32 dataset, chunked and gzip-compressed (mdsuite/database/simulation_database.py:449-499),
// and reads it back through h5py fancy indexing (:594-639), i.e. libhdf5 inflating one chunk after
// the other on one thread.  Here the chunks a batch touches are decoded by all host cores at once:
// every worker thread preads a chunk, inflates it with the decoder below (written for this job:
// 64-bit bit buffer, two-level Huffman tables, 8-byte match copies), undoes shuffle / fletcher32 if
// the file has them, and scatters the chunk's part of the requested box straight into the caller's
// buffer with the caller's strides — which is how the atom-major chunks become the frame-major
// batches the RDF kernels take, without an intermediate array.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "mds_io.h"

namespace mds_io {
void set_error(const std::string& msg);  // lammps_reader.cpp (thread-local message)
}

namespace {

// ---- inflate (RFC 1951) + zlib wrapper (RFC 1950) ---------------------------------------------

constexpr int LIT_BITS = 11;   // primary table of the literal/length code
constexpr int DIST_BITS = 8;   // primary table of the distance code
// table entry: bits 0-7 total code length, bits 8-15 kind / extra-bit count, bits 16-31 value
constexpr uint32_t KIND_LITERAL = 0x80u << 8;
constexpr uint32_t KIND_EOB = 0x40u << 8;
constexpr uint32_t KIND_SUB = 0x20u << 8;      // pointer to a second-level table (bits 0-7: its width)
constexpr uint32_t KIND_INVALID = 0x10u << 8;  // code not in use

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                               67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                               4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513,
                                769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Table {
  std::vector<uint32_t> e;  // primary (1 << bits) entries, then the second-level tables
  int bits = 0;
};

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// payload of symbol `sym` of the literal/length (is_dist = false) or the distance code
inline bool symbol_entry(int sym, bool is_dist, uint32_t* out) {
  if (is_dist) {
    if (sym >= 30) return false;
    *out = ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
  } else if (sym < 256) {
    *out = ((uint32_t)sym << 16) | KIND_LITERAL;
  } else if (sym == 256) {
    *out = KIND_EOB;
  } else {
    if (sym >= 286) return false;
    *out = ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
  }
  return true;
}

// Canonical Huffman decoding table from code lengths (0 = unused).  Returns false for an
// over-subscribed set; incomplete sets are allowed (unused entries stay invalid), as zlib does for
// the single-code distance tree.
bool build_table(const uint8_t* lens, int n, bool is_dist, int primary_bits, Table& t) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  count[0] = 0;
  int left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
  }
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code + (uint32_t)count[l - 1]) << 1;
    next_code[l] = code;
  }
  t.bits = primary_bits;
  const uint32_t primary = 1u << primary_bits;
  t.e.assign(primary, KIND_INVALID | 1u);
  // second-level tables: widest code per primary prefix
  std::vector<uint8_t> sub_bits;
  std::vector<uint32_t> codes((size_t)n, 0);
  bool any_long = false;
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    codes[(size_t)s] = reverse_bits(next_code[l]++, l);
    if (l > primary_bits) any_long = true;
  }
  if (any_long) {
    sub_bits.assign(primary, 0);
    for (int s = 0; s < n; ++s) {
      const int l = lens[s];
      if (l > primary_bits) {
        uint8_t& b = sub_bits[codes[(size_t)s] & (primary - 1)];
        b = std::max<uint8_t>(b, (uint8_t)(l - primary_bits));
      }
    }
    for (uint32_t p = 0; p < primary; ++p)
      if (sub_bits[p]) {
        const uint32_t off = (uint32_t)t.e.size();
        if (off > 0xffffu) return false;
        t.e.resize(t.e.size() + (1u << sub_bits[p]), KIND_INVALID | (uint32_t)(primary_bits + 1));
        t.e[p] = (off << 16) | KIND_SUB | sub_bits[p];
      }
  }
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    uint32_t payload;
    if (!symbol_entry(s, is_dist, &payload)) {
      // symbols 286/287 and distance codes 30/31 may carry lengths in a fixed block but must
      // never be decoded
      payload = KIND_INVALID;
    }
    const uint32_t c = codes[(size_t)s];
    if (l <= primary_bits) {
      for (uint32_t i = c; i < primary; i += 1u << l) t.e[i] = payload | (uint32_t)l;
    } else {
      const uint32_t p = c & (primary - 1);
      const uint32_t off = t.e[p] >> 16;
      const int sb = (int)(t.e[p] & 0xffu);
      for (uint32_t i = c >> primary_bits; i < (1u << sb); i += 1u << (l - primary_bits))
        t.e[off + i] = payload | (uint32_t)l;
    }
  }
  return true;
}

struct Inflater {
  const uint8_t* in;
  const uint8_t* in_end;
  uint8_t* out;
  uint8_t* out_begin;
  uint8_t* out_end;
  uint64_t buf = 0;
  int cnt = 0;  // valid bits in buf
  Table lit, dist;
  Table fixed_lit, fixed_dist;
  bool have_fixed = false;
  const char* error = nullptr;

  inline void refill() {
    if (in_end - in >= 8) {
      uint64_t w;
      memcpy(&w, in, 8);
      buf |= w << cnt;
      in += (63 - cnt) >> 3;
      cnt |= 56;
    } else {
      while (cnt <= 56 && in < in_end) {
        buf |= (uint64_t)*in++ << cnt;
        cnt += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void drop(int n) { buf >>= n; cnt -= n; }
  inline bool need(int n) {
    if (cnt < n) refill();
    return cnt >= n;
  }
  inline uint32_t take(int n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }

  static inline uint32_t lookup(const Table& t, uint64_t bits) {
    uint32_t e = t.e[bits & ((1u << t.bits) - 1)];
    if (e & KIND_SUB) e = t.e[(e >> 16) + ((bits >> t.bits) & ((1u << (e & 0xffu)) - 1))];
    return e;
  }

  bool stored_block() {
    drop(cnt & 7);  // to the byte boundary
    if (!need(32)) return fail("truncated stored block");
    const uint32_t len = take(16), nlen = take(16);
    if ((len ^ 0xffffu) != nlen) return fail("stored block length check failed");
    // bytes still in the bit buffer first
    uint32_t left = len;
    while (left && cnt >= 8) {
      if (out >= out_end) return fail("output buffer too small");
      *out++ = (uint8_t)take(8);
      --left;
    }
    if (left) {
      buf = 0; cnt = 0;
      if ((size_t)(in_end - in) < left) return fail("truncated stored block");
      if ((size_t)(out_end - out) < left) return fail("output buffer too small");
      memcpy(out, in, left);
      in += left;
      out += left;
    }
    return true;
  }

  bool read_dynamic_tables() {
    if (!need(14)) return fail("truncated block header");
    const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
    if (hlit > 286 || hdist > 30) return fail("too many length or distance codes");
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i) {
      if (!need(3)) return fail("truncated code-length code");
      cl[kClOrder[i]] = (uint8_t)take(3);
    }
    Table clt;
    {
      // the code-length alphabet is decoded with a plain 7-bit table
      int count[8] = {0};
      for (int i = 0; i < 19; ++i) count[cl[i]]++;
      count[0] = 0;
      int left = 1;
      for (int l = 1; l <= 7; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return fail("over-subscribed code-length code");
      }
      uint32_t next_code[8], code = 0;
      for (int l = 1; l <= 7; ++l) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
      }
      clt.bits = 7;
      clt.e.assign(128, KIND_INVALID | 1u);
      for (int s = 0; s < 19; ++s)
        if (cl[s]) {
          const uint32_t c = reverse_bits(next_code[cl[s]]++, cl[s]);
          for (uint32_t i = c; i < 128; i += 1u << cl[s]) clt.e[i] = ((uint32_t)s << 16) | cl[s];
        }
    }
    uint8_t lens[286 + 30];
    int n = 0;
    while (n < hlit + hdist) {
      refill();
      if (cnt == 0) return fail("truncated code lengths");
      const uint32_t e = clt.e[peek(7)];
      if (e & KIND_INVALID) return fail("invalid code-length symbol");
      if ((int)(e & 0xff) > cnt) return fail("truncated code lengths");
      drop((int)(e & 0xff));
      const int sym = (int)(e >> 16);
      if (sym < 16) {
        lens[n++] = (uint8_t)sym;
      } else {
        int rep, val = 0;
        if (sym == 16) {
          if (n == 0) return fail("repeat without a previous length");
          if (!need(2)) return fail("truncated code lengths");
          val = lens[n - 1];
          rep = 3 + (int)take(2);
        } else if (sym == 17) {
          if (!need(3)) return fail("truncated code lengths");
          rep = 3 + (int)take(3);
        } else {
          if (!need(7)) return fail("truncated code lengths");
          rep = 11 + (int)take(7);
        }
        if (n + rep > hlit + hdist) return fail("code lengths overrun");
        while (rep--) lens[n++] = (uint8_t)val;
      }
    }
    if (lens[256] == 0) return fail("no end-of-block code");
    if (!build_table(lens, hlit, false, LIT_BITS, lit)) return fail("invalid literal/length code");
    if (!build_table(lens + hlit, hdist, true, DIST_BITS, dist)) return fail("invalid distance code");
    return true;
  }

  void build_fixed() {
    uint8_t l[288];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    build_table(l, 288, false, LIT_BITS, fixed_lit);
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) d[i] = 5;
    build_table(d, 32, true, DIST_BITS, fixed_dist);
    have_fixed = true;
  }

  bool fail(const char* msg) {
    if (!error) error = msg;
    return false;
  }

  // The decoder state lives in locals for the duration of a block: stores through the output
  // pointer (uint8_t*) may alias anything, members would be re-read from memory after each one.
  bool huffman_block(const Table& lt_table, const Table& dt_table) {
    const uint32_t* const lt = lt_table.e.data();
    const uint32_t* const dt = dt_table.e.data();
    const uint32_t lmask = (1u << lt_table.bits) - 1, dmask = (1u << dt_table.bits) - 1;
    const int lbits = lt_table.bits, dbits = dt_table.bits;
    const uint8_t* ip = in;
    const uint8_t* const ie = in_end;
    uint8_t* op = out;
    uint8_t* const oe = out_end;
    uint8_t* const ob = out_begin;
    uint64_t bb = buf;
    int bc = cnt;
    const char* err = nullptr;
    bool done = false;
#define MDS_LOOKUP(tab, mask, bits, e)                                                   \
  do {                                                                                   \
    (e) = (tab)[bb & (mask)];                                                            \
    if ((e) & KIND_SUB) (e) = (tab)[((e) >> 16) + ((bb >> (bits)) & ((1u << ((e) & 0xffu)) - 1))]; \
  } while (0)
#define MDS_REFILL_FAST()                 \
  do {                                    \
    uint64_t w_;                          \
    memcpy(&w_, ip, 8);                   \
    bb |= w_ << bc;                       \
    ip += (63 - bc) >> 3;                 \
    bc |= 56;                             \
  } while (0)
    while (!done && !err) {
      // fast loop: enough input for whole symbols without running dry, enough room for the
      // longest match plus the 8-byte copy overshoot
      while (ie - ip >= 16 && oe - op >= 274) {
        MDS_REFILL_FAST();  // >= 56 bits
        uint32_t e = lt[bb & lmask];
        // Literals straight out of the primary table are at most LIT_BITS = 11 bits long: five of
        // them fit one refill.  (Positions barely compress — most of the stream is literals.)
#define MDS_LITERAL_STEP()                 \
  do {                                     \
    bb >>= (e & 63);                       \
    bc -= (int)(e & 63);                   \
    *op++ = (uint8_t)(e >> 16);            \
    e = lt[bb & lmask];                    \
  } while (0)
        if (__builtin_expect((e & KIND_LITERAL) != 0, 1)) {
          MDS_LITERAL_STEP();
          if (e & KIND_LITERAL) {
            MDS_LITERAL_STEP();
            if (e & KIND_LITERAL) {
              MDS_LITERAL_STEP();
              if (e & KIND_LITERAL) {
                MDS_LITERAL_STEP();
                if (e & KIND_LITERAL) {
                  bb >>= (e & 63); bc -= (int)(e & 63);
                  *op++ = (uint8_t)(e >> 16);
                  continue;
                }
              }
            }
          }
          MDS_REFILL_FAST();  // e was looked up in bits that are still the lowest ones
        }
#undef MDS_LITERAL_STEP
        if (e & KIND_SUB) {  // a code longer than the primary table
          e = lt[(e >> 16) + ((bb >> lbits) & ((1u << (e & 0xffu)) - 1))];
          if (e & KIND_LITERAL) {
            bb >>= (e & 63); bc -= (int)(e & 63);
            *op++ = (uint8_t)(e >> 16);
            continue;
          }
        }
        if (e & (KIND_EOB | KIND_INVALID)) {
          if (e & KIND_INVALID) { err = "invalid literal/length code"; break; }
          bb >>= (e & 63); bc -= (int)(e & 63);
          done = true;
          break;
        }
        bb >>= (e & 63); bc -= (int)(e & 63);
        const int lx = (int)((e >> 8) & 0xff);
        const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1ull << lx) - 1));  // <= 15 + 5 bits of >= 56
        bb >>= lx; bc -= lx;
        uint32_t d;
        MDS_LOOKUP(dt, dmask, dbits, d);
        if (d & KIND_INVALID) { err = "invalid distance code"; break; }
        bb >>= (d & 63); bc -= (int)(d & 63);
        const int dx = (int)((d >> 8) & 0xff);
        const uint32_t distance = (d >> 16) + (uint32_t)(bb & ((1ull << dx) - 1));  // 48 bits at most
        bb >>= dx; bc -= dx;
        if (distance > (uint32_t)(op - ob)) { err = "distance beyond the start of the output"; break; }
        const uint8_t* from = op - distance;
        uint8_t* to = op;
        op += len;
        if (distance >= 8) {
          do {
            memcpy(to, from, 8);
            to += 8; from += 8;
          } while (to < op);
        } else if (distance == 1) {
          memset(to, *from, len);
        } else {
          do { *to++ = *from++; } while (to < op);
        }
      }
      if (done || err) break;
      // careful path near the ends of the buffers: one symbol, every step checked
      in = ip; out = op; buf = bb; cnt = bc;
      const bool ok = careful_symbol(lt_table, dt_table, &done);
      ip = in; op = out; bb = buf; bc = cnt;
      if (!ok) { err = error ? error : "inflate failed"; break; }
    }
#undef MDS_LOOKUP
#undef MDS_REFILL_FAST
    in = ip; out = op; buf = bb; cnt = bc;
    if (err) return fail(err);
    return true;
  }

  // one symbol (a literal, a match or the end-of-block code) with every access checked
  bool careful_symbol(const Table& lt, const Table& dt, bool* end_of_block) {
    refill();
    uint32_t e = lookup(lt, buf);
    if (e & KIND_INVALID) return fail(cnt ? "invalid literal/length code" : "truncated stream");
    if ((int)(e & 0xff) > cnt) return fail("truncated stream");
    drop((int)(e & 0xff));
    if (e & KIND_LITERAL) {
      if (out >= out_end) return fail("output buffer too small");
      *out++ = (uint8_t)(e >> 16);
      return true;
    }
    if (e & KIND_EOB) {
      *end_of_block = true;
      return true;
    }
    const int lx = (int)((e >> 8) & 0xff);
    if (!need(lx)) return fail("truncated stream");
    uint32_t len = (e >> 16) + take(lx);
    refill();
    const uint32_t d = lookup(dt, buf);
    if (d & KIND_INVALID) return fail(cnt ? "invalid distance code" : "truncated stream");
    if ((int)(d & 0xff) > cnt) return fail("truncated stream");
    drop((int)(d & 0xff));
    const int dx = (int)((d >> 8) & 0xff);
    if (!need(dx)) return fail("truncated stream");
    const uint32_t distance = (d >> 16) + take(dx);
    if (distance > (uint32_t)(out - out_begin)) return fail("distance beyond the start of the output");
    if ((size_t)(out_end - out) < len) return fail("output buffer too small");
    const uint8_t* from = out - distance;
    while (len--) *out++ = *from++;
    return true;
  }

  bool run() {
    for (;;) {
      if (!need(3)) return fail("truncated stream");
      const uint32_t last = take(1), type = take(2);
      bool ok;
      if (type == 0) ok = stored_block();
      else if (type == 1) {
        if (!have_fixed) build_fixed();
        ok = huffman_block(fixed_lit, fixed_dist);
      } else if (type == 2) ok = read_dynamic_tables() && huffman_block(lit, dist);
      else ok = fail("reserved block type");
      if (!ok) return false;
      if (last) return true;
    }
  }
};

uint32_t adler32(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
  while (n) {
    size_t k = std::min<size_t>(n, 5552);
    n -= k;
    while (k >= 8) {
      a += p[0]; b += a; a += p[1]; b += a; a += p[2]; b += a; a += p[3]; b += a;
      a += p[4]; b += a; a += p[5]; b += a; a += p[6]; b += a; a += p[7]; b += a;
      p += 8; k -= 8;
    }
    while (k--) { a += *p++; b += a; }
    a %= 65521u; b %= 65521u;
  }
  return (b << 16) | a;
}

// zlib stream -> dst.  Returns nullptr on success, else a message.
const char* inflate_zlib(Inflater& inf, const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                         size_t* produced, bool verify) {
  if (n < 6) return "zlib stream too short";
  const uint32_t cmf = src[0], flg = src[1];
  if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0) return "bad zlib header";
  if (flg & 0x20) return "zlib preset dictionaries are not supported";
  inf.in = src + 2;
  inf.in_end = src + n;  // (the 4 checksum bytes are never reached by a well-formed stream)
  inf.out = inf.out_begin = dst;
  inf.out_end = dst + cap;
  inf.buf = 0;
  inf.cnt = 0;
  inf.error = nullptr;
  if (!inf.run()) return inf.error ? inf.error : "inflate failed";
  *produced = (size_t)(inf.out - dst);
  if (verify) {
    // the checksum follows the last block at the next byte boundary; bytes already in the bit
    // buffer are given back first
    const uint8_t* tail = inf.in - (inf.cnt >> 3);
    if (inf.in_end - tail < 4) return "zlib checksum missing";
    const uint32_t want = ((uint32_t)tail[0] << 24) | ((uint32_t)tail[1] << 16) |
                          ((uint32_t)tail[2] << 8) | tail[3];
    if (adler32(dst, *produced) != want) return "zlib checksum mismatch";
  }
  return nullptr;
}

// ---- chunks -> caller's box --------------------------------------------------------------------

struct Worker {
  Inflater inf;
  std::vector<uint8_t> raw, a, b;
};

const char* decode_chunk(Worker& w, int fd, const mds_h5_chunk& c, const mds_h5_layout& L,
                         const uint8_t** data, size_t* size) {
  w.raw.resize((size_t)c.nbytes);
  size_t got = 0;
  while (got < (size_t)c.nbytes) {
    const ssize_t r = pread(fd, w.raw.data() + got, (size_t)c.nbytes - got, (off_t)(c.address + (int64_t)got));
    if (r <= 0) return "pread of a chunk failed";
    got += (size_t)r;
  }
  size_t chunk_bytes = (size_t)L.elem_size;
  for (int d = 0; d < L.rank; ++d) chunk_bytes *= (size_t)L.chunk_dims[d];
  const uint8_t* cur = w.raw.data();
  size_t cur_n = w.raw.size();
  std::vector<uint8_t>* bufs[2] = {&w.a, &w.b};
  int which = 0;
  // the pipeline is undone back to front; bit i of the mask = filter i was skipped for this chunk
  for (int i = L.n_filters - 1; i >= 0; --i) {
    if (c.filter_mask & (1u << i)) continue;
    std::vector<uint8_t>& dst = *bufs[which];
    switch (L.filter_ids[i]) {
      case 1: {  // deflate
        dst.resize(chunk_bytes + 8);
        size_t produced = 0;
        if (const char* err = inflate_zlib(w.inf, cur, cur_n, dst.data(), chunk_bytes + 8, &produced,
                                           L.verify_checksums != 0))
          return err;
        cur = dst.data();
        cur_n = produced;
        which ^= 1;
        break;
      }
      case 2: {  // shuffle: byte k of every element was stored together
        const size_t es = (size_t)(L.filter_cd0[i] > 0 ? L.filter_cd0[i] : L.elem_size);
        const size_t n = cur_n / es;
        dst.resize(cur_n);
        for (size_t k = 0; k < es; ++k) {
          const uint8_t* s = cur + k * n;
          uint8_t* d = dst.data() + k;
          for (size_t j = 0; j < n; ++j) d[j * es] = s[j];
        }
        memcpy(dst.data() + n * es, cur + n * es, cur_n - n * es);
        cur = dst.data();
        which ^= 1;
        break;
      }
      case 3:  // fletcher32: trailing 4-byte checksum
        if (cur_n < 4) return "fletcher32 chunk too short";
        cur_n -= 4;
        break;
      default:
        return "unsupported HDF5 filter";
    }
  }
  if (cur_n < chunk_bytes) return "chunk is shorter than its declared shape";
  *data = cur;
  *size = cur_n;
  return nullptr;
}

// Copy the part of the chunk that lies in the box [lo, hi) to out; out index of dataset element i
// is sum_d (i_d - lo_d) * out_strides[d] (in elements).
void place_chunk(const uint8_t* data, const mds_h5_chunk& c, const mds_h5_layout& L, uint8_t* out) {
  int64_t a[4] = {0, 0, 0, 0}, b[4] = {1, 1, 1, 1}, cs[4] = {0, 0, 0, 0};  // chunk-local box, strides
  const int r = L.rank;
  int64_t stride = 1;
  for (int d = r - 1; d >= 0; --d) {
    cs[d] = stride;
    stride *= L.chunk_dims[d];
  }
  int64_t base = 0;
  for (int d = 0; d < r; ++d) {
    const int64_t lo = std::max(L.box_lo[d], c.offsets[d]);
    const int64_t hi = std::min(L.box_hi[d], c.offsets[d] + L.chunk_dims[d]);
    if (hi <= lo) return;
    a[d] = lo - c.offsets[d];
    b[d] = hi - c.offsets[d];
    base += (lo - L.box_lo[d]) * L.out_strides[d];
  }
  const size_t es = (size_t)L.elem_size;
  // normalise to 4 dimensions (leading ones of extent 1)
  int64_t A[4] = {0, 0, 0, 0}, B[4] = {1, 1, 1, 1}, CS[4] = {0, 0, 0, 0}, OS[4] = {0, 0, 0, 0};
  for (int d = 0; d < r; ++d) {
    const int k = 4 - r + d;
    A[k] = a[d]; B[k] = b[d]; CS[k] = cs[d]; OS[k] = L.out_strides[d];
  }
  const int64_t n3 = B[3] - A[3];
  const bool inner_contig = (OS[3] == 1);
  // Loop order: the dimension with the larger output stride goes outside, so that writes walk the
  // output as sequentially as the box allows (atom-major chunks -> frame-major batches swap 1 and 2).
  const bool swap12 = OS[2] > OS[1];
  const int d1 = swap12 ? 2 : 1, d2 = swap12 ? 1 : 2;
  for (int64_t i0 = A[0]; i0 < B[0]; ++i0)
    for (int64_t i1 = A[d1]; i1 < B[d1]; ++i1) {
      const uint8_t* s_row = data + (size_t)(i0 * CS[0] + i1 * CS[d1] + A[d2] * CS[d2] + A[3] * CS[3]) * es;
      uint8_t* o_row = out + (size_t)(base + (i0 - A[0]) * OS[0] + (i1 - A[d1]) * OS[d1]) * es;
      for (int64_t i2 = A[d2]; i2 < B[d2]; ++i2) {
        if (inner_contig) {
          if (n3 * (int64_t)es == 12) {  // (x, y, z) float32: the case of every Positions dataset
            memcpy(o_row, s_row, 12);
          } else {
            memcpy(o_row, s_row, (size_t)n3 * es);
          }
        } else {
          for (int64_t i3 = 0; i3 < n3; ++i3)
            memcpy(o_row + (size_t)(i3 * OS[3]) * es, s_row + (size_t)(i3 * CS[3]) * es, es);
        }
        s_row += (size_t)CS[d2] * es;
        o_row += (size_t)OS[d2] * es;
      }
    }
}

}  // namespace

extern "C" {

int mds_inflate(const void* src, int64_t src_len, void* dst, int64_t dst_capacity, int64_t* dst_len,
                int verify_checksum) {
  if (!src || !dst || src_len < 0 || dst_capacity < 0 || !dst_len) {
    mds_io::set_error("mds_inflate: bad argument");
    return MDS_IO_ERR_INVALID;
  }
  Inflater inf;
  size_t produced = 0;
  if (const char* err = inflate_zlib(inf, (const uint8_t*)src, (size_t)src_len, (uint8_t*)dst,
                                     (size_t)dst_capacity, &produced, verify_checksum != 0)) {
    mds_io::set_error(std::string("mds_inflate: ") + err);
    return MDS_IO_ERR_FORMAT;
  }
  *dst_len = (int64_t)produced;
  return MDS_IO_OK;
}

int mds_h5_read_chunks(int fd, const mds_h5_layout* layout, const mds_h5_chunk* chunks,
                       int64_t n_chunks, void* out, int n_threads) {
  if (!layout || (!chunks && n_chunks > 0) || !out || layout->struct_size != sizeof(mds_h5_layout) ||
      layout->rank < 1 || layout->rank > 4 || layout->elem_size < 1 || layout->n_filters < 0 ||
      layout->n_filters > 8) {
    mds_io::set_error("mds_h5_read_chunks: bad argument");
    return MDS_IO_ERR_INVALID;
  }
  if (n_chunks == 0) return MDS_IO_OK;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n_chunks));
  std::atomic<int64_t> next(0);
  std::atomic<const char*> first_error(nullptr);
  auto work = [&]() {
    Worker w;
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_chunks || first_error.load()) return;
      const uint8_t* data = nullptr;
      size_t size = 0;
      if (const char* err = decode_chunk(w, fd, chunks[i], *layout, &data, &size)) {
        const char* expected = nullptr;
        first_error.compare_exchange_strong(expected, err);
        return;
      }
      place_chunk(data, chunks[i], *layout, (uint8_t*)out);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  if (const char* err = first_error.load()) {
    mds_io::set_error(std::string("mds_h5_read_chunks: ") + err);
    return MDS_IO_ERR_FORMAT;
  }
  return MDS_IO_OK;
}

}  // extern "C"
