// Angular distribution function for sm_100a behind include/mds_adf.h: neighbour search with the
// reference's float16 cutoff, then every (i; j, k) triple of a centre atom binned by its angle.
//
// Arithmetic contract (oracle/adf_oracle.py; reference utils/neighbour_list.py:63-78,104-177,
// utils/linalg.py:31-81), float32 with every operation rounded separately:
//   r = pos[i] - pos[j];  r -= rint(r / L) * L;  s = (x*x + y*y) + z*z;  n = sqrt(s)
//   neighbour iff float16(n) != 0 and float16(n) < float16(r_cut)   <=>  s_lo <= s < s_hi
//   u = r / n (per component);  c = clip((ux1*ux2 + uy1*uy2) + uz1*uz2, -1, 1)
//   theta = float32(acos(c)), bin by np.histogram over [0, 3.15]    <=>  c against cos_edges[]
//   weight = 1 / (n1 * n2)^norm_power   (formed as n1^-p * n2^-p, see normalise())
// The two "<=>" are threshold tables built on the host with the same IEEE operations, so the kernel
// needs neither a float16 conversion nor an exact arccosine: it guesses the bin with acosf() and
// settles it by comparing c with the table.
//
// Both orders (i, j, k) and (i, k, j) are triples of the reference.  A species combination
// (A, B, C), A <= B <= C, takes the triples with species (i, j, k) = (A, B, C): for B == C both
// orders (same angle and weight: added once with weight 2w), for B < C one of them.  So every
// unordered pair of neighbours with species >= A contributes exactly once.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mds_adf.h"
#include "rdf_common.cuh"

namespace {

thread_local std::string g_adf_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_adf_error = buf;
  return code;
}

#define ADF_CUDA_TRY(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(MDS_ADF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

// CTA shapes: 8 warps (5 CTAs per SM) by default; 4 warps (10 CTAs per SM) when combinations x bins
// is small — twice as many CTA-private histograms per SM halve the retries of the float64
// compare-and-swap atomics (one species, 500 bins: 8.4 instead of 9.7 ms on the 100k-atom LJ
// fluid; four combinations: 1.93 instead of 1.78 ms, hence the threshold).
constexpr int CENTRES_PER_WARP = 16;  // consecutive centres of an item one warp works through
constexpr int SMALL_HIST_LEN = 1024;  // combinations x bins up to which 4-warp CTAs are used
constexpr int MAX_NB_LIMIT = MDS_ADF_MAX_NEIGHBOURS;
constexpr int N_SPANS = 18;  // 9 cell rows x up to 2 pieces where the x window wraps
constexpr int MAX_SPECIES = 8;
// frame_flags bits written by the cell-count kernel
constexpr uint32_t ADF_FRAME_FAR = 1u;    // a coordinate beyond 8 box lengths: cell search declines
constexpr uint32_t ADF_FRAME_OUT_A = 2u;  // a coordinate outside [-L/8, 9L/8]
constexpr uint32_t ADF_FRAME_OUT_B = 4u;  // a coordinate outside [-5L/8, 5L/8]

// ---- host: threshold tables -------------------------------------------------------------------

// np.histogram's bin of a float32 x for `nbins` equal bins over [0, 3.15]
// (numpy/lib/_histograms_impl.py, the uniform-bin branch).  For float32 data NumPy keeps float32
// throughout: the edges are float32(linspace in float64), the scaled position is computed in
// float32 with float32(3.15) and float32(nbins), then the index is settled against the edges.
int numpy_bin_of(float x, int nbins) {
  const float hi32 = (float)MDS_ADF_RANGE_HI;
  const double step = MDS_ADF_RANGE_HI / (double)nbins;  // np.linspace: edge k = k * step
  auto edge = [&](int k) { return k == nbins ? hi32 : (float)((double)k * step); };
  volatile float q = x / hi32;  // two separately rounded float32 operations
  volatile float f = q * (float)nbins;
  int idx = (int)f;
  if (idx == nbins) idx -= 1;
  if (x < edge(idx)) idx -= 1;
  else if (x >= edge(idx + 1) && idx != nbins - 1) idx += 1;
  return idx;
}

int bin_of_cos(float c, int nbins) {
  const float theta = (float)acos((double)c);
  return numpy_bin_of(theta, nbins);
}

// order-preserving map float32 <-> uint32 for bisection over all floats in [-1, 1]
uint32_t key_of(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
float float_of(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void build_cos_edges(int nbins, std::vector<float>& edges) {
  edges.assign((size_t)nbins + 1, -2.0f);
  edges[0] = INFINITY;
  const uint32_t k_lo = key_of(-1.0f), k_hi = key_of(1.0f);
  for (int k = 1; k < nbins; ++k) {
    // bin_of_cos is non-increasing in c: find the largest c in [-1, 1] with bin >= k
    if (bin_of_cos(-1.0f, nbins) < k) { edges[(size_t)k] = -2.0f; continue; }
    uint32_t lo = k_lo, hi = k_hi;  // invariant: bin(lo) >= k
    if (bin_of_cos(1.0f, nbins) >= k) { edges[(size_t)k] = 1.0f; continue; }
    while (hi - lo > 1) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (bin_of_cos(float_of(mid), nbins) >= k) lo = mid; else hi = mid;
    }
    edges[(size_t)k] = float_of(lo);
  }
}

float smallest_s_with_sqrt_ge(float d) {
  if (d <= 0.0f) return 0.0f;
  float s = (float)((double)d * (double)d);
  while (s > 0.0f && sqrtf(s) >= d) s = nextafterf(s, -INFINITY);
  while (sqrtf(s) < d) s = nextafterf(s, INFINITY);
  return s;
}

inline float half_round(float f) { return __half2float(__float2half_rn(f)); }

// neighbour iff n_lo <= n < n_hi for n = float32 norm
int neighbour_window_norm(float cutoff, float* n_lo, float* n_hi) {
  if (!(cutoff > 0.0f) || !std::isfinite(cutoff) || cutoff > 60000.0f)
    return fail(MDS_ADF_ERR_INVALID, "bad cutoff %g (need 0 < cutoff <= 60000)", (double)cutoff);
  const float cut16 = half_round(cutoff);
  if (!(cut16 > 0.0f))
    return fail(MDS_ADF_ERR_INVALID, "cutoff %g is zero as a float16", (double)cutoff);
  float hi = cut16;  // float16(hi) == cut16, not below it; walk down to the first float that is
  while (half_round(nextafterf(hi, -INFINITY)) >= cut16) hi = nextafterf(hi, -INFINITY);
  // smallest float whose float16 is not zero: 2^-25 is the tie between 0 and the smallest
  // subnormal 2^-24 and goes to the even one, 0; the next float up rounds to 2^-24
  const float tie = ldexpf(1.0f, -25);
  const float lo = nextafterf(tie, INFINITY);
  if (half_round(tie) != 0.0f || half_round(lo) == 0.0f)
    return fail(MDS_ADF_ERR_INVALID, "internal: unexpected float16 rounding at 2^-25");
  *n_lo = lo;
  *n_hi = hi;
  return MDS_ADF_OK;
}

}  // namespace

// ---- device -----------------------------------------------------------------------------------

struct AdfGrid {
  int nx, ny, nz, nc;  // cells per dimension (each >= 3), nc = nx * ny * nz
  float box[3];
};

struct AdfParams {
  const float* pos;         // [F][N][3]
  const float4* sorted;     // [F][N] (x, y, z, species | cell << 3), atoms ordered by cell (cell path)
  const uint32_t* cell_off; // [F][nc + 1] first sorted slot of every cell (cell path)
  const uint32_t* frame_flags;  // [F] != 0: a coordinate beyond 8 box lengths (cell path declines)
  unsigned long long* pair_tests;
  AdfGrid grid;
  int only_flagged;         // all-atoms form: process the flagged configurations only
  int max_nb;               // capacity of a warp's neighbour list
  int hist_in_smem;         // 0: combinations x bins do not fit in shared memory, sums go straight
                            //    to the configuration's global histogram
  const uint8_t* species;   // [N]
  const float* cos_edges;   // [nbins + 1]
  double* out;              // [F][combos][nbins]
  unsigned long long* work_counter;
  unsigned long long* triples;
  int* status;              // [0]: overflow flag, [1]: largest neighbour list
  int n_atoms;
  int n_frames;
  int items_per_frame;
  int nbins;
  int n_combos;
  int norm_power;
  float box[3];
  float s_lo, s_hi;
  float bin_scale;          // nbins / 3.15
  int8_t combo_of[MAX_SPECIES * MAX_SPECIES * MAX_SPECIES];  // [A][B][C], B <= C
};

__device__ __forceinline__ float min_image(float d, float box) {
  const float q = __fdiv_rn(d, box);
  return __fsub_rn(d, __fmul_rn(rintf(q), box));
}

// Signed minimum image for |d| <= 1.25 box (all coordinates of the configuration inside one of
// the windows of mds::window_bits): there rint(d / box) is -1, 0 or 1, the reference's result has
// magnitude min(|d|, box - |d|) bit for bit and it is d itself iff |d| <= box - |d|
// (box - |d| is exact for |d| >= box / 2), otherwise d -+ box = -+(box - |d|).
__device__ __forceinline__ float min_image_fast(float d, float box) {
  const float a = fabsf(d);
  const float b = __fsub_rn(box, a);
  return (a <= b) ? d : (d > 0.0f ? -b : b);
}

// Test atom j (position pj, species sp_j) against centre i and append it to the warp's neighbour
// list: unit vector, norm, species.  Warp-collective (ballot); `valid` masks lanes without an atom.
template <bool FAST>
__device__ __forceinline__ void consider(bool valid, float xi, float yi, float zi, float xj, float yj,
                                         float zj, int sp_i, int sp_j, const AdfParams& p,
                                         float4* my_vec, uint8_t* my_sp, int& m, int lane) {
  bool is_nb = false;
  float dx = 0.f, dy = 0.f, dz = 0.f, s = 0.f;
  if (valid) {
    dx = __fsub_rn(xi, xj);
    dy = __fsub_rn(yi, yj);
    dz = __fsub_rn(zi, zj);
    if (FAST) {
      dx = min_image_fast(dx, p.box[0]);
      dy = min_image_fast(dy, p.box[1]);
      dz = min_image_fast(dz, p.box[2]);
    } else {
      dx = min_image(dx, p.box[0]);
      dy = min_image(dy, p.box[1]);
      dz = min_image(dz, p.box[2]);
    }
    s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    is_nb = (s >= p.s_lo) && (s < p.s_hi) && (sp_j >= sp_i);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, is_nb);
  if (is_nb) {
    const int slot = m + __popc(ballot & ((1u << lane) - 1u));
    if (slot < p.max_nb) {
      my_vec[slot] = make_float4(dx, dy, dz, s);  // normalised after the search, all lanes busy
      my_sp[slot] = (uint8_t)sp_j;
    }
  }
  m += __popc(ballot);
}

// (r, |r|^2) -> (r / |r|, |r|^-norm_power) for the m entries of a neighbour list.  The weight of a
// triple, 1 / (|r_ij| |r_ik|)^p, is formed as |r_ij|^-p * |r_ik|^-p: a few float32 roundings away
// from the contract's value (<= 4e-7 relative, inside the tolerance of the weighted sums; the
// counts, norm_power = 0, do not depend on it).
__device__ __forceinline__ void normalise(float4* my_vec, int m, int lane, int norm_power) {
  for (int k = lane; k < m; k += 32) {
    const float4 v = my_vec[k];
    const float n = __fsqrt_rn(v.w);
    float q = 1.0f;
    if (norm_power > 0) {
      float pw = n;
      if (norm_power == 4) {
        pw = __fmul_rn(n, n);
        pw = __fmul_rn(pw, pw);
      } else {
        for (int e = 1; e < norm_power; ++e) pw = __fmul_rn(pw, n);
      }
      q = __frcp_rn(pw);
    }
    my_vec[k] = make_float4(__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n), q);
  }
}

// acos(c) to ~7e-5 rad (Abramowitz & Stegun 4.4.45) — only the GUESS of the bin; the table decides
__device__ __forceinline__ float acos_guess(float c) {
  const float ax = fabsf(c);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f - ax));
  const float poly = ((-0.0187293f * ax + 0.0742610f) * ax - 0.2121144f) * ax + 1.5707288f;
  const float t = r * poly;
  return c < 0.0f ? 3.14159265f - t : t;
}

// Every unordered pair (a, b), a < b, of the neighbour list, flattened over the lanes.  Row a holds
// b = a + 1 .. m - 1 (m - 1 - a entries); rows R and m - 2 - R together hold exactly m entries, so
// position t of the flattened list is entry c = t % m of combined row R = t / m: the pair
// (R, R + 1 + c) for c < m - 1 - R, else entry c - (m - 1 - R) of row m - 2 - R.  (For even m the
// middle row is its own partner; the list ends after its m / 2 entries.)
// ONE_SPECIES: a single combination, every pair counts twice — no species bytes, no table.
template <bool ONE_SPECIES>
__device__ __forceinline__ void bin_triples(const float4* my_vec, const uint8_t* my_sp, int m,
                                            int sp_i, const AdfParams& p, double* hist,
                                            double* fout, const float* edges,
                                            const int8_t* s_combo, int lane,
                                            unsigned long long& my_triples) {
  const int total = (m * (m - 1)) >> 1;
  unsigned n_triples = 0;  // <= 2 * 1024 * 1023 / 2 per centre
  int R = 0, col = lane;
  for (int t = lane; t < total; t += 32) {
    while (col >= m) { col -= m; ++R; }
    const int first = m - 1 - R;  // entries of row R in this combined row
    const int a = col < first ? R : m - 2 - R;
    const int b = col < first ? R + 1 + col : a + 1 + (col - first);
    col += 32;
    const float4 va = my_vec[a];
    const float4 vb = my_vec[b];
    float c = __fadd_rn(__fadd_rn(__fmul_rn(va.x, vb.x), __fmul_rn(va.y, vb.y)),
                        __fmul_rn(va.z, vb.z));
    c = fminf(fmaxf(c, -1.0f), 1.0f);
    int g = (int)(acos_guess(c) * p.bin_scale);
    g = max(0, min(g, p.nbins - 1));
    // edges[0] = +inf and edges[nbins] = -2 bound both steps.  The guess is within 7e-5 rad: for
    // bins wider than that (nbins <= 16384) one step either way settles it, otherwise walk.
    g += (c <= edges[g + 1]) ? 1 : 0;
    g -= (c > edges[g]) ? 1 : 0;
    if (p.nbins > 16384) {
      while (c <= edges[g + 1]) ++g;
      while (c > edges[g]) --g;
    }
    int combo = 0;
    float w = __fmul_rn(va.w, vb.w);
    if (ONE_SPECIES) {
      w = __fmul_rn(2.0f, w);
      n_triples += 2u;
    } else {
      const int sa = my_sp[a], sb = my_sp[b];
      const int lo = min(sa, sb), hi = max(sa, sb);
      combo = s_combo[(sp_i * MAX_SPECIES + lo) * MAX_SPECIES + hi];
      w = __fmul_rn((sa == sb) ? 2.0f : 1.0f, w);
      n_triples += (sa == sb) ? 2u : 1u;
    }
    if (p.hist_in_smem) atomicAdd(&hist[combo * p.nbins + g], (double)w);
    else atomicAdd(&fout[combo * p.nbins + g], (double)w);
  }
  my_triples += n_triples;
}

// Span table of the 27-cell stencil around `cell` for one warp.  The x window [cx - 1, cx + 1] of
// a cell row is one contiguous span of the sorted array, or two where it wraps around the box: 18
// spans at most.  Lane s < 18 fetches the bounds of span s (one round trip for all of them); the
// non-empty spans (half of them are the unused wrap pieces) are stored as running ends of ONE
// flattened candidate list plus the offset that maps a list position to a sorted slot.  Returns
// the length of the list.  Warp-collective.
__device__ __forceinline__ int build_span_table(int cell, const AdfGrid& g, const uint32_t* off,
                                                int lane, int* span_end, int* span_adj) {
  const int cx = cell % g.nx;
  const int cy = (cell / g.nx) % g.ny;
  const int cz = cell / (g.nx * g.ny);
  int len = 0, begin = 0;
  if (lane < N_SPANS) {
    const int r = lane >> 1, part = lane & 1;
    int z = cz + r / 3 - 1, y = cy + r % 3 - 1;
    z += (z < 0) ? g.nz : 0;
    z -= (z >= g.nz) ? g.nz : 0;
    y += (y < 0) ? g.ny : 0;
    y -= (y >= g.ny) ? g.ny : 0;
    int x0, x1;
    if (part == 0) { x0 = max(cx - 1, 0); x1 = min(cx + 1, g.nx - 1); }
    else if (cx == 0) { x0 = x1 = g.nx - 1; }
    else if (cx == g.nx - 1) { x0 = x1 = 0; }
    else { x0 = 0; x1 = -1; }
    if (x1 >= x0) {
      const int row = (z * g.ny + y) * g.nx;
      begin = (int)__ldg(off + row + x0);
      len = (int)__ldg(off + row + x1 + 1) - begin;
    }
  }
  int incl = len;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
  if (len > 0) {
    const int k = __popc(nonempty & ((1u << lane) - 1u));
    span_end[k] = incl;
    span_adj[k] = begin - (incl - len);
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, incl, 31);
}

// CELLS = false: every atom of the configuration is tested against the centre (any input; the
//   frames flagged in p.frame_flags only, if p.only_flagged).
// CELLS = true: the configuration was counting-sorted into cells of width >= 1.002 x the largest
//   norm that can pass the float16 test (x fastest), and only the 27 cells around the centre are
//   searched; frames with a coordinate beyond 8 box lengths are flagged and left to the other form
//   (the cell-assignment error bound of DESIGN.md §9 needs |x| <= 8 L).
template <bool CELLS, bool ONE_SPECIES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 10 : 5) adf_kernel(const AdfParams p) {
  constexpr int ADF_WARPS = WARPS, ADF_THREADS = WARPS * 32;
  constexpr int CENTRES_PER_ITEM = WARPS * CENTRES_PER_WARP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* nb_vec = reinterpret_cast<float4*>(smem_raw);                       // [WARPS][max_nb]
  uint8_t* nb_sp = reinterpret_cast<uint8_t*>(nb_vec + ADF_WARPS * p.max_nb);  // [WARPS][max_nb]
  double* hist = reinterpret_cast<double*>(nb_sp + ADF_WARPS * p.max_nb);      // [combos][nbins]
  float* edges = reinterpret_cast<float*>(hist + (p.hist_in_smem ? p.n_combos * p.nbins : 0));
  __shared__ unsigned long long s_item;
  __shared__ int s_span_end[ADF_WARPS][N_SPANS + 1];  // running end of every span in the flattened list
  __shared__ int s_span_adj[ADF_WARPS][N_SPANS];      // sorted slot = flattened index + adj
  __shared__ int8_t s_combo[MAX_SPECIES * MAX_SPECIES * MAX_SPECIES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hist_len = p.n_combos * p.nbins;
  if (p.hist_in_smem)
    for (int k = tid; k < hist_len; k += ADF_THREADS) hist[k] = 0.0;
  for (int k = tid; k <= p.nbins; k += ADF_THREADS) edges[k] = p.cos_edges[k];
  for (int k = tid; k < MAX_SPECIES * MAX_SPECIES * MAX_SPECIES; k += ADF_THREADS)
    s_combo[k] = p.combo_of[k];
  float4* my_vec = nb_vec + warp * p.max_nb;
  uint8_t* my_sp = nb_sp + warp * p.max_nb;
  const unsigned long long n_items = (unsigned long long)p.n_frames * p.items_per_frame;
  unsigned long long my_triples = 0, my_tests = 0;
  int my_max_nb = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(p.work_counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int frame = (int)(item / p.items_per_frame);
    const uint32_t fbits = p.frame_flags != nullptr ? p.frame_flags[frame] : 0u;
    const bool flagged = (fbits & ADF_FRAME_FAR) != 0u;
    if (CELLS ? flagged : (p.only_flagged && !flagged)) continue;
    // every coordinate inside one of the two windows: |d| <= 1.25 box, the 3-op minimum image holds
    const bool fast = CELLS && (fbits & (ADF_FRAME_OUT_A | ADF_FRAME_OUT_B)) !=
                                   (ADF_FRAME_OUT_A | ADF_FRAME_OUT_B);
    const int c0 = (int)(item % p.items_per_frame) * CENTRES_PER_ITEM;
    const int c1 = min(p.n_atoms, c0 + CENTRES_PER_ITEM);

    // A warp takes consecutive centres of the sorted order: they mostly share a cell, and the span
    // table of the 27-cell stencil is rebuilt only when the cell changes.
    constexpr int PER_WARP = CENTRES_PER_WARP;
    int prev_cell = -1, total = 0;
    for (int i = c0 + warp * PER_WARP; i < min(c1, c0 + (warp + 1) * PER_WARP); ++i) {
      int m = 0, sp_i;
      if (CELLS) {
        const float4* sorted = p.sorted + (size_t)frame * p.n_atoms;
        const uint32_t* off = p.cell_off + (size_t)frame * (p.grid.nc + 1);
        const float4 ci = __ldg(sorted + i);
        sp_i = __float_as_int(ci.w) & (MAX_SPECIES - 1);
        const int cell = __float_as_int(ci.w) >> 3;
        if (cell != prev_cell) {
          prev_cell = cell;
          total = build_span_table(cell, p.grid, off, lane, s_span_end[warp], s_span_adj[warp]);
        }
        my_tests += (unsigned long long)total;
        int span = 0, span_end = total > 0 ? s_span_end[warp][0] : 0;
        for (int t0 = 0; t0 < total; t0 += 32) {
          const int t = t0 + lane;
          const bool valid = t < total;
          float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            while (t >= span_end) span_end = s_span_end[warp][++span];
            pj = __ldg(sorted + t + s_span_adj[warp][span]);
          }
          if (fast)
            consider<true>(valid, ci.x, ci.y, ci.z, pj.x, pj.y, pj.z, sp_i,
                           __float_as_int(pj.w) & (MAX_SPECIES - 1), p, my_vec, my_sp, m, lane);
          else
            consider<false>(valid, ci.x, ci.y, ci.z, pj.x, pj.y, pj.z, sp_i,
                            __float_as_int(pj.w) & (MAX_SPECIES - 1), p, my_vec, my_sp, m, lane);
        }
        __syncwarp();
      } else {
        const float* fpos = p.pos + (size_t)frame * p.n_atoms * 3;
        const float xi = __ldg(fpos + 3 * i), yi = __ldg(fpos + 3 * i + 1), zi = __ldg(fpos + 3 * i + 2);
        sp_i = p.species[i];
        my_tests += (unsigned long long)p.n_atoms;
        for (int j0 = 0; j0 < p.n_atoms; j0 += 32) {
          const int j = j0 + lane;
          const bool valid = j < p.n_atoms;
          float xj = 0.f, yj = 0.f, zj = 0.f;
          int sp_j = 0;
          if (valid) {
            xj = __ldg(fpos + 3 * j);
            yj = __ldg(fpos + 3 * j + 1);
            zj = __ldg(fpos + 3 * j + 2);
            sp_j = p.species[j];
          }
          consider<false>(valid, xi, yi, zi, xj, yj, zj, sp_i, sp_j, p, my_vec, my_sp, m, lane);
        }
      }
      if (m > p.max_nb) {
        if (lane == 0) atomicExch(p.status, 1);
        m = p.max_nb;
      }
      my_max_nb = max(my_max_nb, m);
      __syncwarp();
      normalise(my_vec, m, lane, p.norm_power);
      __syncwarp();
      bin_triples<ONE_SPECIES>(my_vec, my_sp, m, sp_i, p, hist, p.out + (size_t)frame * hist_len, edges, s_combo,
                  lane, my_triples);
      __syncwarp();
    }
    // ---- flush this item's sums into the configuration's histogram ----
    __syncthreads();
    double* fout = p.out + (size_t)frame * hist_len;
    for (int k = tid; p.hist_in_smem && k < hist_len; k += ADF_THREADS) {
      const double v = hist[k];
      if (v != 0.0) {
        atomicAdd(fout + k, v);
        hist[k] = 0.0;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_triples += __shfl_xor_sync(0xffffffffu, my_triples, o);
    my_max_nb = max(my_max_nb, __shfl_xor_sync(0xffffffffu, my_max_nb, o));
  }
  if (lane == 0) {
    if (my_triples) atomicAdd(p.triples, my_triples);
    if (my_tests) atomicAdd(p.pair_tests, my_tests);
    atomicMax(p.status + 1, my_max_nb);
  }
}

// ---- cell build: count + rank, scan, scatter (one configuration = one segment) ----------------

// blockIdx.y walks the configurations, the x dimension the atoms of one (no integer division)
__global__ void adf_cell_count_kernel(const float* __restrict__ pos, int n_atoms, int n_frames,
                                      AdfGrid g, uint32_t* __restrict__ cell_cnt,
                                      uint32_t* __restrict__ cell_id, uint32_t* __restrict__ rank,
                                      uint32_t* __restrict__ frame_flags) {
  for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
    const size_t base = (size_t)f * n_atoms;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n_atoms; a += gridDim.x * blockDim.x) {
      const size_t idx = base + a;
      const float x = pos[3 * idx], y = pos[3 * idx + 1], z = pos[3 * idx + 2];
      uint32_t wb = mds::window_bits(x, g.box[0]) | mds::window_bits(y, g.box[1]) |
                    mds::window_bits(z, g.box[2]);  // bit0: outside window A, bit1: outside window B
      uint32_t bits = wb << 1;
      if (fabsf(x) > 8.0f * g.box[0] || fabsf(y) > 8.0f * g.box[1] || fabsf(z) > 8.0f * g.box[2])
        bits |= ADF_FRAME_FAR;
      // one atomic per warp instead of one per atom (all lanes of a warp are in configuration f)
      const unsigned act = __activemask();
      bits = __reduce_or_sync(act, bits);
      if (bits && (int)(threadIdx.x & 31) == __ffs(act) - 1 && (frame_flags[f] & bits) != bits)
        atomicOr(frame_flags + f, bits);
      const int c = mds::cell_of(x, y, z, g);
      cell_id[idx] = (uint32_t)c;
      rank[idx] = atomicAdd(cell_cnt + (size_t)f * (g.nc + 1) + c, 1u);
    }
  }
}

// one CTA per configuration: counts[0 .. nc] -> exclusive offsets, offsets[nc] = n_atoms
__global__ void __launch_bounds__(256) adf_cell_scan_kernel(uint32_t* __restrict__ cell_cnt, int nc) {
  __shared__ uint32_t s_part[256];
  uint32_t* c = cell_cnt + (size_t)blockIdx.x * (nc + 1);
  const int n = nc + 1;
  const int per = (n + 255) / 256;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  uint32_t sum = 0;
  for (int k = lo; k < hi; ++k) sum += c[k];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const uint32_t v = threadIdx.x >= (unsigned)o ? s_part[threadIdx.x - o] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = s_part[threadIdx.x] - sum;
  for (int k = lo; k < hi; ++k) {
    const uint32_t v = c[k];
    c[k] = run;
    run += v;
  }
}

__global__ void adf_cell_scatter_kernel(const float* __restrict__ pos,
                                        const uint8_t* __restrict__ species, int n_atoms,
                                        int n_frames, int nc, const uint32_t* __restrict__ cell_off,
                                        const uint32_t* __restrict__ cell_id,
                                        const uint32_t* __restrict__ rank,
                                        float4* __restrict__ sorted) {
  for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
    const size_t base = (size_t)f * n_atoms;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n_atoms; a += gridDim.x * blockDim.x) {
      const size_t idx = base + a;
      const uint32_t cell = cell_id[idx];
      const uint32_t slot = cell_off[(size_t)f * (nc + 1) + cell] + rank[idx];
      // w carries the species (3 bits) and the cell id above it (< 2^21 cells)
      sorted[base + slot] = make_float4(pos[3 * idx], pos[3 * idx + 1], pos[3 * idx + 2],
                                        __int_as_float((int)species[a] | ((int)cell << 3)));
    }
  }
}

// ---- host API ---------------------------------------------------------------------------------

struct mds_adf {
  int n_species = 0;
  std::vector<int64_t> counts;
  int nbins = 0;
  float cutoff = 0.f;
  float box[3] = {0, 0, 0};
  int norm_power = 0;
  int device = 0;
  int batch_frames = 0;
  int sm_count = 0;
  int64_t n_atoms = 0;
  int n_combos = 0;
  float s_lo = 0.f, s_hi = 0.f;
  std::vector<float> cos_edges;
  AdfParams base;
  size_t smem_bytes = 0;
  int grid = 0;
  bool use_cells = false;
  AdfGrid cells;
  int warps = 8;           // warps per CTA of the pair kernel (4 for small histograms)
  int max_nb = 0;          // capacity of a warp's neighbour list in the current configuration
  bool hist_in_smem = true;
  size_t smem_limit = 0;
  // device
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
  float* d_pos = nullptr;
  double* d_out = nullptr;
  uint8_t* d_species = nullptr;
  float* d_edges = nullptr;
  unsigned long long* d_counters = nullptr;  // [0] work counter, [1] triples, [2] pair tests
  int* d_status = nullptr;
  float4* d_sorted = nullptr;     // cell path: [batch][N]
  uint32_t* d_cell_off = nullptr; // [batch][nc + 1]
  uint32_t* d_cell_id = nullptr;  // [batch][N]
  uint32_t* d_rank = nullptr;     // [batch][N]
  uint32_t* d_flags = nullptr;    // [batch]
  std::vector<uint32_t> h_flags;
  int64_t frames_far = 0;
  int work_frames = 0;     // configurations the cell-build buffers hold
  int staging_frames = 0;  // configurations d_pos / d_out hold (host entry point only)
  // statistics
  int64_t frames = 0, launches = 0;
  double triples = 0, pair_tests = 0;
  float kernel_ms = 0, compute_ms = 0;
  int max_neighbours = 0;
};

namespace {

// Grow the per-configuration device buffers to `frames` configurations (<= batch_frames):
// allocation is tens of milliseconds per 100 MB, so a handle only pays for what it is given.
int ensure_frames(mds_adf* h, int frames, bool staging) {
  auto grow = [&](int have) { return std::min(h->batch_frames, std::max(frames, have * 2)); };
  if (h->use_cells && frames > h->work_frames) {
    const int n = grow(h->work_frames);
    ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_sorted); cudaFree(h->d_cell_id); cudaFree(h->d_rank);
    cudaFree(h->d_cell_off); cudaFree(h->d_flags);
    h->d_sorted = nullptr; h->d_cell_id = nullptr; h->d_rank = nullptr;
    h->d_cell_off = nullptr; h->d_flags = nullptr;
    h->work_frames = 0;
    const size_t atoms = (size_t)h->n_atoms * (size_t)n;
    ADF_CUDA_TRY(cudaMalloc(&h->d_sorted, sizeof(float4) * atoms));
    ADF_CUDA_TRY(cudaMalloc(&h->d_cell_id, sizeof(uint32_t) * atoms));
    ADF_CUDA_TRY(cudaMalloc(&h->d_rank, sizeof(uint32_t) * atoms));
    ADF_CUDA_TRY(cudaMalloc(&h->d_cell_off, sizeof(uint32_t) * ((size_t)h->cells.nc + 1) * (size_t)n));
    ADF_CUDA_TRY(cudaMalloc(&h->d_flags, sizeof(uint32_t) * (size_t)n));
    h->work_frames = n;
  }
  if (staging && frames > h->staging_frames) {
    const int n = grow(h->staging_frames);
    ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_pos); cudaFree(h->d_out);
    h->d_pos = nullptr; h->d_out = nullptr;
    h->staging_frames = 0;
    ADF_CUDA_TRY(cudaMalloc(&h->d_pos, sizeof(float) * 3 * (size_t)h->n_atoms * (size_t)n));
    ADF_CUDA_TRY(cudaMalloc(&h->d_out, sizeof(double) * (size_t)h->n_combos * h->nbins * (size_t)n));
    h->staging_frames = n;
  }
  return MDS_ADF_OK;
}

size_t smem_for(const mds_adf* h, int max_nb, bool hist_in_smem) {
  return (size_t)h->warps * (size_t)max_nb * (sizeof(float4) + 1) +
         (hist_in_smem ? sizeof(double) * (size_t)h->n_combos * h->nbins : 0) +
         sizeof(float) * ((size_t)h->nbins + 1);
}

typedef void (*adf_kernel_t)(const AdfParams);
adf_kernel_t adf_kernel_for(bool cells, bool one_species, int warps) {
  if (warps == 4) {
    if (cells) return one_species ? adf_kernel<true, true, 4> : adf_kernel<true, false, 4>;
    return one_species ? adf_kernel<false, true, 4> : adf_kernel<false, false, 4>;
  }
  if (cells) return one_species ? adf_kernel<true, true, 8> : adf_kernel<true, false, 8>;
  return one_species ? adf_kernel<false, true, 8> : adf_kernel<false, false, 8>;
}

// shared-memory layout, launch attributes and grid for a neighbour-list capacity
int configure(mds_adf* h, int max_nb) {
  // the CTA's histogram lives in shared memory when it leaves room for two CTAs per SM, else the
  // sums go straight to global memory (many species x many bins)
  bool hist_in_smem = smem_for(h, max_nb, true) <= h->smem_limit / 2;
  const char* env = getenv("MDS_ADF_HIST");
  if (env && !strcmp(env, "global")) hist_in_smem = false;
  const size_t bytes = smem_for(h, max_nb, hist_in_smem);
  if (bytes > h->smem_limit)
    return fail(MDS_ADF_ERR_UNSUPPORTED,
                "%d bins with %d neighbours per atom need %zu bytes of shared memory (limit %zu)",
                h->nbins, max_nb, bytes, h->smem_limit);
  // the attribute belongs to the function, not to the handle: always allow the device maximum
  const bool one = h->n_species == 1;
  int per_sm = 1 << 30;
  for (int cells = 0; cells < 2; ++cells) {
    adf_kernel_t k = adf_kernel_for(cells != 0, one, h->warps);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit) !=
        cudaSuccess)
      return fail(MDS_ADF_ERR_CUDA, "cudaFuncSetAttribute failed: %s",
                  cudaGetErrorString(cudaGetLastError()));
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, h->warps * 32, bytes) != cudaSuccess ||
        n < 1)
      return fail(MDS_ADF_ERR_CUDA, "the ADF kernel does not fit on an SM");
    per_sm = std::min(per_sm, n);
  }
  h->max_nb = max_nb;
  h->smem_bytes = bytes;
  h->grid = per_sm * h->sm_count;
  h->base.max_nb = max_nb;
  h->hist_in_smem = hist_in_smem;
  h->base.hist_in_smem = hist_in_smem ? 1 : 0;
  return MDS_ADF_OK;
}

constexpr int ADF_OVERFLOW = 1;  // run_batch_once: a neighbour list did not fit

int run_batch_once(mds_adf* h, const float* d_xyz, int n_frames, double* d_out) {
  const size_t out_bytes = (size_t)n_frames * h->n_combos * h->nbins * sizeof(double);
  ADF_CUDA_TRY(cudaMemsetAsync(d_out, 0, out_bytes, h->stream));
  ADF_CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, 3 * sizeof(unsigned long long), h->stream));
  AdfParams p = h->base;
  p.pos = d_xyz;
  p.out = d_out;
  p.n_frames = n_frames;
  const unsigned long long items = (unsigned long long)n_frames * p.items_per_frame;
  const int grid = std::max(1, (int)std::min<unsigned long long>((unsigned long long)h->grid, items));
  int64_t far = 0;
  if (h->use_cells) {
    const int nc = h->cells.nc;
    ADF_CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, sizeof(uint32_t) * (size_t)n_frames, h->stream));
    ADF_CUDA_TRY(cudaMemsetAsync(h->d_cell_off, 0, sizeof(uint32_t) * (size_t)n_frames * (nc + 1),
                                 h->stream));
    const int bx = (int)std::min<int64_t>((h->n_atoms + 255) / 256, (int64_t)h->sm_count * 16);
    const int by = std::max(1, std::min({n_frames, 65535, (h->sm_count * 16 + bx - 1) / bx}));
    const dim3 blocks((unsigned)bx, (unsigned)by);
    adf_cell_count_kernel<<<blocks, 256, 0, h->stream>>>(d_xyz, (int)h->n_atoms, n_frames, h->cells,
                                                         h->d_cell_off, h->d_cell_id, h->d_rank,
                                                         h->d_flags);
    adf_cell_scan_kernel<<<n_frames, 256, 0, h->stream>>>(h->d_cell_off, nc);
    adf_cell_scatter_kernel<<<blocks, 256, 0, h->stream>>>(d_xyz, h->d_species, (int)h->n_atoms,
                                                           n_frames, nc, h->d_cell_off,
                                                           h->d_cell_id, h->d_rank, h->d_sorted);
    p.sorted = h->d_sorted;
    p.cell_off = h->d_cell_off;
    p.frame_flags = h->d_flags;
    p.grid = h->cells;
    adf_kernel_for(true, h->n_species == 1, h->warps)<<<grid, h->warps * 32, h->smem_bytes, h->stream>>>(p);
    ADF_CUDA_TRY(cudaGetLastError());
    h->launches += 4;
    h->h_flags.resize((size_t)n_frames);
    ADF_CUDA_TRY(cudaMemcpyAsync(h->h_flags.data(), h->d_flags, sizeof(uint32_t) * (size_t)n_frames,
                                 cudaMemcpyDeviceToHost, h->stream));
    ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < n_frames; ++f) far += (h->h_flags[(size_t)f] & ADF_FRAME_FAR) != 0u;
    if (far) {  // configurations with a coordinate beyond 8 box lengths: all-atoms search
      ADF_CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long), h->stream));
      p.only_flagged = 1;
      adf_kernel_for(false, h->n_species == 1, h->warps)<<<grid, h->warps * 32, h->smem_bytes, h->stream>>>(p);
      ADF_CUDA_TRY(cudaGetLastError());
      h->launches += 1;
    }
  } else {
    adf_kernel_for(false, h->n_species == 1, h->warps)<<<grid, h->warps * 32, h->smem_bytes, h->stream>>>(p);
    ADF_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
  }
  unsigned long long counters[3];
  int status[2];
  ADF_CUDA_TRY(cudaMemcpyAsync(counters, h->d_counters, sizeof(counters), cudaMemcpyDeviceToHost,
                               h->stream));
  ADF_CUDA_TRY(cudaMemcpyAsync(status, h->d_status, sizeof(status), cudaMemcpyDeviceToHost,
                               h->stream));
  ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (status[0]) {
    ADF_CUDA_TRY(cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream));
    return ADF_OVERFLOW;
  }
  h->triples += (double)counters[1];
  h->pair_tests += (double)counters[2];
  h->max_neighbours = std::max(h->max_neighbours, status[1]);
  h->frames += n_frames;
  h->frames_far += far;
  return MDS_ADF_OK;
}

// A neighbour list that does not fit makes the batch invalid: double the capacity (fewer CTAs per
// SM from then on) and run the batch again, up to MDS_ADF_MAX_NEIGHBOURS.
int run_batch(mds_adf* h, const float* d_xyz, int n_frames, double* d_out) {
  for (;;) {
    const int rc = run_batch_once(h, d_xyz, n_frames, d_out);
    if (rc != ADF_OVERFLOW) return rc;
    if (h->max_nb >= MAX_NB_LIMIT)
      return fail(MDS_ADF_ERR_UNSUPPORTED,
                  "an atom has more than %d neighbours within the cutoff %g; reduce the cutoff",
                  MAX_NB_LIMIT, (double)h->cutoff);
    const int rc2 = configure(h, h->max_nb * 2);
    if (rc2 != MDS_ADF_OK) return rc2;
  }
}

}  // namespace

extern "C" {

const char* mds_adf_last_error(void) { return g_adf_error.c_str(); }
int mds_adf_abi_version(void) { return MDS_ADF_ABI_VERSION; }

int mds_adf_build_threshold_table(int nbins, float* cos_edges) {
  if (nbins < 1 || nbins > (1 << 20) || !cos_edges)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_build_threshold_table: bad nbins %d or NULL", nbins);
  std::vector<float> e;
  build_cos_edges(nbins, e);
  memcpy(cos_edges, e.data(), sizeof(float) * ((size_t)nbins + 1));
  return MDS_ADF_OK;
}

int mds_adf_neighbour_window(float cutoff, float* s_lo, float* s_hi) {
  if (!s_lo || !s_hi) return fail(MDS_ADF_ERR_INVALID, "mds_adf_neighbour_window: NULL argument");
  float n_lo, n_hi;
  const int rc = neighbour_window_norm(cutoff, &n_lo, &n_hi);
  if (rc != MDS_ADF_OK) return rc;
  *s_lo = smallest_s_with_sqrt_ge(n_lo);
  *s_hi = smallest_s_with_sqrt_ge(n_hi);
  return MDS_ADF_OK;
}

int mds_adf_create(mds_adf_t** out, const mds_adf_config* cfg) {
  if (!out || !cfg) return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: NULL argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(mds_adf_config))
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: struct_size %u != %zu", cfg->struct_size,
                sizeof(mds_adf_config));
  if (cfg->n_species < 1 || cfg->n_species > MAX_SPECIES || !cfg->counts)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: n_species %d (1..%d)", cfg->n_species,
                MAX_SPECIES);
  if (cfg->nbins < 1 || cfg->nbins > 65536)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: nbins %d (1..65536)", cfg->nbins);
  if (cfg->norm_power < 0 || cfg->norm_power > 16)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: norm_power %d (0..16)", cfg->norm_power);
  for (int k = 0; k < 3; ++k)
    if (!(cfg->box[k] > 0.0f) || !std::isfinite(cfg->box[k]))
      return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: bad box length %g", (double)cfg->box[k]);
  int64_t n_atoms = 0;
  for (int s = 0; s < cfg->n_species; ++s) {
    if (cfg->counts[s] < 0) return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: negative count");
    n_atoms += cfg->counts[s];
  }
  if (n_atoms < 1 || n_atoms > (1ll << 30))
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: %lld atoms", (long long)n_atoms);
  float s_lo, s_hi;
  int rc = mds_adf_neighbour_window(cfg->cutoff, &s_lo, &s_hi);
  if (rc != MDS_ADF_OK) return rc;

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
    return fail(MDS_ADF_ERR_CUDA, "no CUDA device: this library has no CPU path");
  if (cfg->device < 0 || cfg->device >= n_dev)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_create: device %d of %d", cfg->device, n_dev);
  mds_adf* h = new (std::nothrow) mds_adf();
  if (!h) return fail(MDS_ADF_ERR_NOMEM, "mds_adf_create: out of memory");
  h->n_species = cfg->n_species;
  h->counts.assign(cfg->counts, cfg->counts + cfg->n_species);
  h->nbins = cfg->nbins;
  h->cutoff = cfg->cutoff;
  memcpy(h->box, cfg->box, sizeof(h->box));
  h->norm_power = cfg->norm_power;
  h->device = cfg->device;
  h->n_atoms = n_atoms;
  h->s_lo = s_lo;
  h->s_hi = s_hi;
  const int S = h->n_species;
  h->n_combos = S * (S + 1) * (S + 2) / 6;
  build_cos_edges(h->nbins, h->cos_edges);

  auto cleanup = [&](int code) { mds_adf_destroy(h); return code; };
  if (cudaSetDevice(h->device) != cudaSuccess) return cleanup(fail(MDS_ADF_ERR_CUDA, "cudaSetDevice failed"));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess)
    return cleanup(fail(MDS_ADF_ERR_CUDA, "cudaGetDeviceProperties failed"));
  h->sm_count = prop.multiProcessorCount;
  {
    // dynamic shared memory available to the kernels: the opt-in maximum minus their static part
    cudaFuncAttributes fa, fb;
    h->warps = h->n_combos * h->nbins <= SMALL_HIST_LEN ? 4 : 8;
    const char* wenv = getenv("MDS_ADF_WARPS");
    if (wenv && (atoi(wenv) == 4 || atoi(wenv) == 8)) h->warps = atoi(wenv);
    if (cudaFuncGetAttributes(&fa, adf_kernel_for(true, cfg->n_species == 1, h->warps)) != cudaSuccess ||
        cudaFuncGetAttributes(&fb, adf_kernel_for(false, cfg->n_species == 1, h->warps)) != cudaSuccess)
      return cleanup(fail(MDS_ADF_ERR_CUDA, "cudaFuncGetAttributes failed: %s",
                          cudaGetErrorString(cudaGetLastError())));
    h->smem_limit = (size_t)prop.sharedMemPerBlockOptin - std::max(fa.sharedSizeBytes, fb.sharedSizeBytes);
  }
  {
    // capacity of the neighbour lists: 2.5 x the mean number of atoms inside the cutoff sphere
    // (+ 32), as a power of two in [64, MDS_ADF_MAX_NEIGHBOURS]; doubled on demand (run_batch)
    const double volume = (double)h->box[0] * h->box[1] * h->box[2];
    const double expected = (double)n_atoms / volume * 4.18879 * std::pow((double)s_hi, 1.5);
    int cap = 64;
    while (cap < MAX_NB_LIMIT && (double)cap < 2.5 * expected + 32.0) cap *= 2;
    const char* env = getenv("MDS_ADF_NEIGHBOUR_CAPACITY");
    if (env && atoi(env) >= 32 && atoi(env) <= MAX_NB_LIMIT) cap = atoi(env);
    while (cap > 64 && smem_for(h, cap, false) > h->smem_limit) cap /= 2;
    rc = configure(h, cap);
    if (rc != MDS_ADF_OK) return cleanup(rc);
  }

  // Cell grid: width >= 1.002 x the largest norm that passes the float16 test, at most 128 cells
  // per dimension, at least 3 (otherwise the 27-cell stencil would visit a cell twice); used from
  // 27 cells and 256 atoms on, and switched off with MDS_ADF_PATH=allatoms (=cells insists).
  {
    const double reach = 1.002 * std::sqrt((double)s_hi);
    int n[3];
    for (int k = 0; k < 3; ++k) n[k] = (int)std::min(128.0, std::floor((double)h->box[k] / reach));
    h->use_cells = n[0] >= 3 && n[1] >= 3 && n[2] >= 3 && n_atoms >= 256;
    const char* env = getenv("MDS_ADF_PATH");
    if (env && !strcmp(env, "allatoms")) h->use_cells = false;
    if (env && !strcmp(env, "cells")) {
      if (n[0] < 3 || n[1] < 3 || n[2] < 3)
        return cleanup(fail(MDS_ADF_ERR_UNSUPPORTED, "the cell path needs box >= 3.006 x cutoff"));
      h->use_cells = true;
    }
    h->cells.nx = n[0]; h->cells.ny = n[1]; h->cells.nz = n[2];
    h->cells.nc = h->use_cells ? n[0] * n[1] * n[2] : 0;
    memcpy(h->cells.box, h->box, sizeof(h->box));
  }

  const size_t frame_bytes = (size_t)n_atoms * 3 * sizeof(float);
  const size_t out_frame_bytes = (size_t)h->n_combos * h->nbins * sizeof(double);
  const size_t cell_frame_bytes =
      h->use_cells ? (size_t)n_atoms * (sizeof(float4) + 8) + sizeof(uint32_t) * ((size_t)h->cells.nc + 2)
                   : 0;
  int64_t bf = cfg->max_frames_in_flight > 0
                   ? cfg->max_frames_in_flight
                   : (int64_t)((512u << 20) / (frame_bytes + out_frame_bytes + cell_frame_bytes));
  bf = std::max<int64_t>(1, std::min<int64_t>(bf, 4096));
  h->batch_frames = (int)bf;

#define ADF_TRY_CREATE(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return cleanup(fail(_e == cudaErrorMemoryAllocation ? MDS_ADF_ERR_NOMEM : MDS_ADF_ERR_CUDA, \
                          "%s failed: %s", #expr, cudaGetErrorString(_e)));                   \
  } while (0)
  ADF_TRY_CREATE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  ADF_TRY_CREATE(cudaEventCreate(&h->ev0));
  ADF_TRY_CREATE(cudaEventCreate(&h->ev1));
  ADF_TRY_CREATE(cudaEventCreate(&h->ev_k0));
  ADF_TRY_CREATE(cudaEventCreate(&h->ev_k1));
  ADF_TRY_CREATE(cudaMalloc(&h->d_species, (size_t)n_atoms));
  ADF_TRY_CREATE(cudaMalloc(&h->d_edges, sizeof(float) * ((size_t)h->nbins + 1)));
  ADF_TRY_CREATE(cudaMalloc(&h->d_counters, 3 * sizeof(unsigned long long)));
  // the per-configuration buffers are sized by the calls that come (ensure_frames)
  ADF_TRY_CREATE(cudaMalloc(&h->d_status, 2 * sizeof(int)));
  ADF_TRY_CREATE(cudaMemset(h->d_status, 0, 2 * sizeof(int)));
  {
    std::vector<uint8_t> sp((size_t)n_atoms);
    size_t o = 0;
    for (int s = 0; s < S; ++s)
      for (int64_t k = 0; k < h->counts[(size_t)s]; ++k) sp[o++] = (uint8_t)s;
    ADF_TRY_CREATE(cudaMemcpy(h->d_species, sp.data(), sp.size(), cudaMemcpyHostToDevice));
    ADF_TRY_CREATE(cudaMemcpy(h->d_edges, h->cos_edges.data(), sizeof(float) * ((size_t)h->nbins + 1),
                              cudaMemcpyHostToDevice));
  }
#undef ADF_TRY_CREATE

  AdfParams& p = h->base;
  memset(&p, 0, sizeof(p));
  p.max_nb = h->max_nb;
  p.hist_in_smem = h->hist_in_smem ? 1 : 0;
  p.species = h->d_species;
  p.cos_edges = h->d_edges;
  p.work_counter = h->d_counters;
  p.triples = h->d_counters + 1;
  p.pair_tests = h->d_counters + 2;
  p.status = h->d_status;
  p.n_atoms = (int)n_atoms;
  {
    const int64_t per_item = (int64_t)h->warps * CENTRES_PER_WARP;
    p.items_per_frame = (int)((n_atoms + per_item - 1) / per_item);
  }
  p.nbins = h->nbins;
  p.n_combos = h->n_combos;
  p.norm_power = h->norm_power;
  memcpy(p.box, h->box, sizeof(p.box));
  p.s_lo = s_lo;
  p.s_hi = s_hi;
  p.bin_scale = (float)((double)h->nbins / MDS_ADF_RANGE_HI);
  memset(p.combo_of, -1, sizeof(p.combo_of));
  {
    int c = 0;  // itertools.combinations_with_replacement(range(S), 3) order
    for (int a = 0; a < S; ++a)
      for (int b = a; b < S; ++b)
        for (int d = b; d < S; ++d) p.combo_of[(a * MAX_SPECIES + b) * MAX_SPECIES + d] = (int8_t)c++;
  }
  *out = h;
  return MDS_ADF_OK;
}

void mds_adf_destroy(mds_adf_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_pos);
  cudaFree(h->d_out);
  cudaFree(h->d_species);
  cudaFree(h->d_edges);
  cudaFree(h->d_counters);
  cudaFree(h->d_status);
  cudaFree(h->d_sorted);
  cudaFree(h->d_cell_id);
  cudaFree(h->d_rank);
  cudaFree(h->d_cell_off);
  cudaFree(h->d_flags);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_k0) cudaEventDestroy(h->ev_k0);
  if (h->ev_k1) cudaEventDestroy(h->ev_k1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int mds_adf_compute(mds_adf_t* h, const float* xyz, int64_t n_frames, double* out) {
  if (!h || (n_frames > 0 && (!xyz || !out)) || n_frames < 0)
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_compute: bad argument");
  ADF_CUDA_TRY(cudaSetDevice(h->device));
  const size_t frame_floats = (size_t)h->n_atoms * 3;
  const size_t out_frame = (size_t)h->n_combos * h->nbins;
  ADF_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  float kernel_ms = 0.f;
  {
    const int rc = ensure_frames(h, (int)std::min<int64_t>(h->batch_frames, n_frames), true);
    if (rc != MDS_ADF_OK) return rc;
  }
  for (int64_t f0 = 0; f0 < n_frames; f0 += h->batch_frames) {
    const int nf = (int)std::min<int64_t>(h->batch_frames, n_frames - f0);
    ADF_CUDA_TRY(cudaMemcpyAsync(h->d_pos, xyz + (size_t)f0 * frame_floats,
                                 sizeof(float) * frame_floats * (size_t)nf, cudaMemcpyHostToDevice,
                                 h->stream));
    ADF_CUDA_TRY(cudaEventRecord(h->ev_k0, h->stream));
    const int rc = run_batch(h, h->d_pos, nf, h->d_out);
    if (rc != MDS_ADF_OK) return rc;
    ADF_CUDA_TRY(cudaEventRecord(h->ev_k1, h->stream));
    ADF_CUDA_TRY(cudaMemcpyAsync(out + (size_t)f0 * out_frame, h->d_out,
                                 sizeof(double) * out_frame * (size_t)nf, cudaMemcpyDeviceToHost,
                                 h->stream));
    ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    ADF_CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1));
    kernel_ms += ms;
  }
  ADF_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  ADF_CUDA_TRY(cudaEventSynchronize(h->ev1));
  ADF_CUDA_TRY(cudaEventElapsedTime(&h->compute_ms, h->ev0, h->ev1));
  h->kernel_ms = kernel_ms;
  return MDS_ADF_OK;
}

int mds_adf_compute_device(mds_adf_t* h, const float* d_xyz, int64_t n_frames, double* d_out,
                           void* stream) {
  if (!h || (n_frames > 0 && (!d_xyz || !d_out)) || n_frames < 0 || n_frames > (1 << 30))
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_compute_device: bad argument");
  ADF_CUDA_TRY(cudaSetDevice(h->device));
  ADF_CUDA_TRY(cudaEventRecord(h->ev0, (cudaStream_t)stream));
  ADF_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev0, 0));
  ADF_CUDA_TRY(cudaEventRecord(h->ev_k0, h->stream));
  const size_t frame_floats = (size_t)h->n_atoms * 3;
  const size_t out_frame = (size_t)h->n_combos * h->nbins;
  {
    const int rc = ensure_frames(h, (int)std::min<int64_t>(h->batch_frames, n_frames), false);
    if (rc != MDS_ADF_OK) return rc;
  }
  for (int64_t f0 = 0; f0 < n_frames; f0 += h->batch_frames) {
    const int nf = (int)std::min<int64_t>(h->batch_frames, n_frames - f0);
    const int rc = run_batch(h, d_xyz + (size_t)f0 * frame_floats, nf, d_out + (size_t)f0 * out_frame);
    if (rc != MDS_ADF_OK) return rc;
  }
  ADF_CUDA_TRY(cudaEventRecord(h->ev_k1, h->stream));
  ADF_CUDA_TRY(cudaEventSynchronize(h->ev_k1));
  ADF_CUDA_TRY(cudaEventElapsedTime(&h->kernel_ms, h->ev_k0, h->ev_k1));
  h->compute_ms = h->kernel_ms;
  return MDS_ADF_OK;
}

int mds_adf_stats_get(mds_adf_t* h, mds_adf_stats* out) {
  if (!h || !out) return fail(MDS_ADF_ERR_INVALID, "mds_adf_stats_get: NULL argument");
  if (out->struct_size != sizeof(mds_adf_stats))
    return fail(MDS_ADF_ERR_INVALID, "mds_adf_stats_get: struct_size %u != %zu", out->struct_size,
                sizeof(mds_adf_stats));
  out->n_combinations = h->n_combos;
  out->frames = h->frames;
  out->triples = h->triples;
  out->pair_tests = h->pair_tests;
  out->kernel_launches = h->launches;
  out->kernel_ms = h->kernel_ms;
  out->compute_ms = h->compute_ms;
  out->max_neighbours = h->max_neighbours;
  out->sm_count = h->sm_count;
  out->cells[0] = h->use_cells ? h->cells.nx : 0;
  out->cells[1] = h->use_cells ? h->cells.ny : 0;
  out->cells[2] = h->use_cells ? h->cells.nz : 0;
  out->frames_far = h->frames_far;
  return MDS_ADF_OK;
}

}  // extern "C"
