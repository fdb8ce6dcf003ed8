// Cell-list RDF path for sm_100a (north-star kernel (b)): large N, cutoff much smaller than the
// box.  The reference has no such path — it evaluates all N(N-1)/2 pairs of every frame
// (SURVEY.md §5 "Long-context"); this path evaluates only pairs of atoms in adjacent cells and
// must return the SAME integer histogram, bit for bit.
//
// Why it is exact: a cell list only decides WHICH pairs are distance-tested.  Every tested pair
// runs the reference's float32 arithmetic on the raw coordinates (rdf_common.cuh; reference
// utils/linalg.py:84-99, calculators/radial_distribution_function.py:667-689), and the squared
// distance is symmetric in (i, j) bit for bit.  The grid is conservative: cell width >= 1.001 x
// cutoff, so for |coordinate| <= 8 box lengths every pair the reference would bin sits in
// adjacent cells (DESIGN.md §9 derives the margin).  Frames with coordinates further out are
// flagged by the pack kernel and handed to the all-pairs kernel instead.
//
// Per batch of frames:
//   rdf_pack_kernel        (rdf_pack.cu) raw xyz -> quad-block positions; per-(frame, species,
//                          cell) counts and ranks with one global atomic per atom; frame flags
//   rdf_cell_scan_kernel   exclusive scan of the counts of a frame -> cell offsets (one CTA/frame)
//   rdf_cell_scatter_kernel counting-sort scatter -> atoms ordered by (species, cell), x fastest
//   rdf_cells_kernel       persistent warps; one work item = (frame, species pair, cell): the
//                          cell's atoms are register rows (8 row groups x 4 column phases per
//                          warp, R = ceil(n/8) rows per lane), the neighbour cells' atoms are a
//                          flattened column stream (x-adjacent cells are contiguous in the sorted
//                          array, so 27 cells are <= 18 spans) staged 32 atoms at a time through
//                          a per-warp double-buffered shared tile; binning as in the all-pairs
//                          kernel (threshold table + ATOMS.POPC.INC, branch-free with a sink slot).
#include "rdf_common.cuh"
#include "rdf_launch.h"
#include "mds_rdf.h"

namespace mds {

// Exclusive scan of the n_keys counts of one frame (in place) + total at [n_keys].
__global__ void __launch_bounds__(1024) rdf_cell_scan_kernel(uint32_t* cell_off, int64_t off_stride,
                                                             int n_keys) {
  uint32_t* a = cell_off + (long long)blockIdx.x * off_stride;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t chunk_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < n_keys; base += 4096) {
    const int i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n_keys) ? a[i0 + k] : 0u;
    const uint32_t t = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = warp_sums[lane];
      uint32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += n;
      }
      warp_sums[lane] = winc - w;
      if (lane == 31) chunk_total = winc;
    }
    __syncthreads();
    uint32_t e = carry + warp_sums[warp] + inc - t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n_keys) a[i0 + k] = e;
      e += v[k];
    }
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n_keys] = carry;
}

__global__ void __launch_bounds__(256) rdf_cell_scatter_kernel(const PackParams p,
                                                               float4* __restrict__ sorted) {
  const long long total = (long long)p.n_frames * p.n_packed;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long f = t / p.n_packed;
    const int q = (int)(t % p.n_packed);
    if (p.src_index[q] < 0) continue;  // padding slot
    const float* src = p.out + (f * p.frame_stride + (q & ~3)) * 3 + (q & 3);
    const float x = src[0], y = src[4], z = src[8];
    const int key = p.species[q] * p.grid.nc + cell_of(x, y, z, p.grid);
    const uint32_t pos = p.cell_off[f * p.off_stride + key] + p.rank[f * p.frame_stride + q];
    sorted[f * p.frame_stride + pos] = make_float4(x, y, z, __int_as_float(q));
  }
}

// ---- the pair kernel -------------------------------------------------------------------------

constexpr int CELL_THREADS = 256;
constexpr int CELL_WARPS = CELL_THREADS / 32;
constexpr int CH = 32;         // column atoms per staged chunk: CH / 32 float4 loads per lane
constexpr int CPL = CH / 32;   // columns staged per lane and chunk
constexpr int MAX_SPANS = 20;  // >= 18 (9 cell rows x 2 pieces when the x window wraps)
constexpr int RG = 8;          // row groups per warp (lanes with the same lane % RG share rows)
constexpr int PH = 32 / RG;    // column phases per warp
constexpr int RMAX = 4;        // rows per lane: one pass covers RG * RMAX = 32 rows of a cell

// One chunk of staged columns (quad blocks in the warp's shared tile) against the register rows
// of a warp.  Lane layout: row group rg = lane % RG (rows rg, rg + RG, ...), column phase
// ph = lane / RG (column pairs 2 ph, 2 ph + 2 PH, ...): three LDS.64 with PH distinct addresses
// serve 2 R pair evaluations per lane through the packed FADD2 / FMUL2 sequence of pair2_s.
// With RG = 4 a cell of n atoms occupies 4 ceil(n / 4) row slots, i.e. at most 3 idle ones.
template <int R, bool FAST, bool DIAG, bool HIST_SMEM>
__device__ __forceinline__ void cell_chunk(const float* __restrict__ tl, int cnt, int q0, int ph,
                                           const float (&xi)[RMAX], const float (&yi)[RMAX],
                                           const float (&zi)[RMAX], const int (&il)[RMAX],
                                           float bx, float by, float bz, float s_cut,
                                           float inv_step, uint32_t c_edge, uint32_t c_hist,
                                           uint32_t four, const float* __restrict__ edges_g,
                                           unsigned long long* hist_g,
                                           unsigned long long& issued) {
  (void)issued;
  float hi = 0.f;
#pragma unroll 1
  for (int c0 = ph * 2; c0 < cnt; c0 += 2 * PH) {
    const float* b = tl + (c0 >> 2) * QUAD_FLOATS + (c0 & 2);
    const float2 X = *reinterpret_cast<const float2*>(b);
    const float2 Y = *reinterpret_cast<const float2*>(b + 4);
    const float2 Z = *reinterpret_cast<const float2*>(b + 8);
    const int q = q0 + c0;  // position in the column stream; the own cell comes first
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s[2];
      pair2_s<FAST>(X.x, X.y, Y.x, Y.y, Z.x, Z.y, xi[r], yi[r], zi[r], bx, by, bz, s[0], s[1]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (HIST_SMEM) {
          bin_count_dense<DIAG, false>(s[c], s_cut, inv_step, 0.f, c_edge, c_hist, four, q + c,
                                       il[r], hi);
          MDS_CHK(issued += 1;)
        } else {
          bool in = s[c] < s_cut;  // false for NaN (padding rows / columns, NaN input)
          if (DIAG) in = in && (q + c > il[r]);
          if (in) atomicAdd(hist_g + bin_of_s(s[c], inv_step, edges_g), 1ull);
        }
      }
    }
  }
}

// Everything a warp needs to walk the column stream of its current work item.
struct WarpStream {
  float* tile;             // [2][CH * 3] this warp's staging tile (quad blocks)
  const int* t_qend;       // span table: stream position where span k ends ...
  const int* t_off;        // ... and (first sorted index of span k) - (stream position of its start)
  const float4* fpos;      // sorted atoms of the frame
  int total;               // stream length
  int lane, ph, st_off;
  float bx, by, bz, s_cut, inv_step;
  uint32_t c_edge, c_hist, four;
  const float* edges_g;
  unsigned long long* hist_g;
  unsigned long long* issued;  // histogram atomics issued by this thread (checked builds)
};

// Columns [qs, qe) of the stream against the register rows of the warp: lanes gather 32 columns
// at a time (one LDG.128 each, issued a whole chunk ahead of its use), transpose them into the
// quad-block tile, and every lane then evaluates its (row group, column phase) share.  DIAG
// applies the col > row mask (only the pass's own block of the own cell needs it).
template <int R, bool FAST, bool DIAG, bool HIST_SMEM>
__device__ __forceinline__ void cell_stream(const WarpStream& w, int qs, int qe,
                                            const float (&xi)[RMAX], const float (&yi)[RMAX],
                                            const float (&zi)[RMAX], const int (&il)[RMAX]) {
  const float4 nan4 = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000),
                                  __int_as_float(0x7fc00000), 0.f);
  int k = 0, cur_end = w.t_qend[0], cur_off = w.t_off[0];
  auto fetch = [&](int q) {  // q is visited in increasing order
    float4 v = nan4;  // slots past the end hold NaN atoms (never inside the cutoff)
    if (q < qe) {
      while (q >= cur_end) {
        ++k;
        cur_end = w.t_qend[k];
        cur_off = w.t_off[k];
      }
      v = __ldg(w.fpos + cur_off + q);
    }
    return v;
  };
  auto stage = [&](float* t, const float4 (&v)[CPL]) {
#pragma unroll
    for (int u = 0; u < CPL; ++u) {  // column lane + 32 u -> quad block 8 u + lane / 4
      float* o = t + u * (8 * QUAD_FLOATS) + w.st_off;
      o[0] = v[u].x; o[4] = v[u].y; o[8] = v[u].z;
    }
  };
  float4 nxt[CPL];
#pragma unroll
  for (int u = 0; u < CPL; ++u) nxt[u] = fetch(qs + 32 * u + w.lane);
  __syncwarp();  // the previous stream's readers are done with the tile
  stage(w.tile, nxt);
  __syncwarp();
  int st = 0;
#pragma unroll 1
  for (int q0 = qs; q0 < qe; q0 += CH, st ^= 1) {
    const int cnt = min(CH, qe - q0);
#pragma unroll
    for (int u = 0; u < CPL; ++u) nxt[u] = fetch(q0 + CH + 32 * u + w.lane);
    const float* tl = w.tile + st * (CH * 3);
    cell_chunk<R, FAST, DIAG, HIST_SMEM>(tl, cnt, q0, w.ph, xi, yi, zi, il, w.bx, w.by, w.bz,
                                         w.s_cut, w.inv_step, w.c_edge, w.c_hist, w.four,
                                         w.edges_g, w.hist_g, *w.issued);
    if (q0 + CH < qe) stage(w.tile + (st ^ 1) * (CH * 3), nxt);  // warp-uniform condition
    __syncwarp();
  }
}

// Move what has accumulated in the CTA's shared histogram to the global counts while other warps
// keep counting: atomicExch takes each bin's value and zeroes it in one step, so nothing is lost.
__device__ __forceinline__ void flush_hist_warp(uint32_t* s_hist, const CellsParams& p, int lane,
                                                unsigned long long* s_chk = nullptr) {
  (void)s_chk;
  const int stride = p.nbins + 1;
  for (int pr = 0; pr < p.pair_cnt; ++pr) {
    unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
    uint32_t* h = s_hist + pr * stride;
    for (int k = lane; k < p.nbins; k += 32) {
      const uint32_t v = atomicExch(h + k, 0u);
      if (v) {
        atomicAdd(g + k, (unsigned long long)v);
        MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)v);)
      }
    }
    if (lane == 0) {  // the sink may wrap; nobody reads it
      const uint32_t v = atomicExch(h + p.nbins, 0u);
      (void)v;
      MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)v);)
    }
  }
}

template <bool HIST_SMEM>
__global__ void __launch_bounds__(CELL_THREADS, 4) rdf_cells_kernel(const CellsParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);                // [WARPS][2][CH * 3] quad blocks
  int* span_tab = reinterpret_cast<int*>(tiles + CELL_WARPS * 2 * CH * 3);   // [WARPS][2][MAX_SPANS]
  uint32_t* s_evals = reinterpret_cast<uint32_t*>(span_tab + CELL_WARPS * 2 * MAX_SPANS);  // [8]
  float* s_edges = reinterpret_cast<float*>(s_evals + 8);
  const int n_edges = p.nbins + 2;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_edges + ((n_edges + 3) & ~3));
  const int hist_stride = p.nbins + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = lane % RG, ph = lane / RG;
  unsigned long long issued = 0;  // only counted in checked builds
  MDS_CHK(__shared__ unsigned long long s_chk[2]; if (tid == 0) { s_chk[0] = 0; s_chk[1] = 0; })

  if (HIST_SMEM) {
    for (int k = tid; k < n_edges; k += CELL_THREADS) s_edges[k] = p.edges[k];
    for (int k = tid; k < p.pair_cnt * hist_stride; k += CELL_THREADS) s_hist[k] = 0u;
  }
  if (tid == 0) {
    s_evals[0] = 0u;
    // address constants of the binning sequence, made opaque to ptxas (see rdf_allpairs.cu)
    s_evals[1] = smem_u32(s_edges + 1) - kMagic4;
    s_evals[2] = smem_u32(s_hist) - kMagic4;
    s_evals[4] = 4u;
    s_evals[5] = 4u;
  }
  __syncthreads();

  float* tile = tiles + warp * 2 * CH * 3;
  const int st_off = (lane >> 2) * QUAD_FLOATS + (lane & 3);  // this lane's slot in a staged chunk
  const float4 nan4 = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000),
                                  __int_as_float(0x7fc00000), 0.f);
  int* t_qend = span_tab + warp * 2 * MAX_SPANS;
  int* t_off = t_qend + MAX_SPANS;
  const uint32_t c_edge = *reinterpret_cast<volatile uint32_t*>(s_evals + 1);
  const uint32_t hist0 = *reinterpret_cast<volatile uint32_t*>(s_evals + 2);
  const uint32_t four = *reinterpret_cast<volatile uint32_t*>(s_evals + 4 + (lane & 1));
  const float bx = p.box[0], by = p.box[1], bz = p.box[2];
  const float s_cut = p.s_cut, inv_step = p.inv_step;
  const int nx = p.nx, ny = p.ny, nz = p.nz, nc = p.nc;
  const unsigned long long total_items =
      (unsigned long long)p.n_frames * (unsigned long long)p.pair_cnt * (unsigned long long)nc;
  const uint32_t publish_every = max(1u, p.flush_threshold >> 6);
  unsigned long long evaluated = 0;  // warp-uniform statistics
  uint32_t unpublished = 0;          // candidate pairs since this warp last told the CTA

  WarpStream ws;
  ws.tile = tile;
  ws.t_qend = t_qend;
  ws.t_off = t_off;
  ws.lane = lane;
  ws.ph = ph;
  ws.st_off = st_off;
  ws.bx = bx; ws.by = by; ws.bz = bz;
  ws.s_cut = s_cut;
  ws.inv_step = inv_step;
  ws.c_edge = c_edge;
  ws.four = four;
  ws.edges_g = p.edges;
  ws.issued = &issued;

  unsigned long long nxt_raw = 0;
  if (lane == 0) nxt_raw = atomicAdd(p.work_counter, 1ull);
  unsigned long long idx = __shfl_sync(0xffffffffu, nxt_raw, 0) * p.shard_count + p.shard_index;

  while (idx < total_items) {
    if (lane == 0) nxt_raw = atomicAdd(p.work_counter, 1ull);  // in flight during this item
    do {
      const int c = (int)(idx % (unsigned long long)nc);
      const unsigned long long t = idx / (unsigned long long)nc;
      const int pr = (int)(t % (unsigned long long)p.pair_cnt);
      const int frame = (int)(t / (unsigned long long)p.pair_cnt);
      const uint32_t fl = p.frame_flags[frame];
      if (fl & MDS_FRAME_FAR) break;  // the all-pairs kernel takes this frame
      const bool fast = (fl & 3u) != 3u;
      const int2 ab = p.pair_ab[p.pair_lo + pr];
      const bool same = ab.x == ab.y;
      const uint32_t* __restrict__ off = p.cell_off + (long long)frame * p.off_stride;
      const int r0 = (int)off[ab.x * nc + c];
      const int n = (int)off[ab.x * nc + c + 1] - r0;
      if (n == 0) break;
      const float4* __restrict__ fpos = p.spos + (long long)frame * p.frame_stride;

      // ---- span table of the column stream (lane l builds span l) ----
      const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
      int start = 0, len = 0;
      {
        const int n_spans = same ? 10 : 18;
        if (lane < n_spans) {
          int dy, dz, part, x0, x1;
          bool whole_row_piece = true;
          if (same) {
            if (lane < 2) {  // own cell (the diagonal block), then the +x neighbour
              dy = 0; dz = 0; part = 0; whole_row_piece = false;
              x0 = (lane == 0) ? cx : (cx + 1 == nx ? 0 : cx + 1);
              x1 = x0 + 1;
            } else {
              const int m = (lane - 2) >> 1;  // (dy, dz): (1,0) (-1,1) (0,1) (1,1)
              part = lane & 1;
              dy = (m == 0) ? 1 : m - 2;
              dz = (m == 0) ? 0 : 1;
            }
          } else {
            const int m = lane >> 1;
            part = lane & 1;
            dy = m % 3 - 1;
            dz = m / 3 - 1;
          }
          if (whole_row_piece) {
            // cells cx-1 .. cx+1 of the row, split in two where the window wraps around
            if (cx == 0)           { x0 = part ? 0 : nx - 1; x1 = part ? 2 : nx; }
            else if (cx == nx - 1) { x0 = part ? 0 : nx - 2; x1 = part ? 1 : nx; }
            else                   { x0 = part ? 0 : cx - 1; x1 = part ? 0 : cx + 2; }
          }
          int yy = cy + dy, zz = cz + dz;
          yy += (yy < 0) ? ny : 0; yy -= (yy >= ny) ? ny : 0;
          zz += (zz < 0) ? nz : 0; zz -= (zz >= nz) ? nz : 0;
          const int kb = ab.y * nc + (zz * ny + yy) * nx;
          if (x1 > x0) {
            start = (int)off[kb + x0];
            len = (int)off[kb + x1] - start;
          }
        }
      }
      int qend = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, qend, o);
        if (lane >= o) qend += v;
      }
      const int total = __shfl_sync(0xffffffffu, qend, 31);
      __syncwarp();  // previous item's readers of the table are done
      if (lane < MAX_SPANS) {
        t_qend[lane] = qend;
        t_off[lane] = start - (qend - len);
      }
      __syncwarp();

      ws.fpos = fpos;
      ws.total = total;
      ws.c_hist = hist0 + (uint32_t)(pr * hist_stride) * 4u;
      ws.hist_g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;

      // ---- row passes of up to 32 atoms of the cell ----
      for (int rp = 0; rp < n; rp += RG * RMAX) {
        const int nr = min(RG * RMAX, n - rp);
        const int R = fast ? (nr + RG - 1) / RG : RMAX;
        float xi[RMAX], yi[RMAX], zi[RMAX];
        int il[RMAX];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          il[r] = rp + rg + RG * r;
          if (il[r] < n) {
            const float4 v = __ldg(fpos + r0 + il[r]);
            xi[r] = v.x; yi[r] = v.y; zi[r] = v.z;
          } else {
            xi[r] = yi[r] = zi[r] = __int_as_float(0x7fc00000);  // NaN: never in cutoff
          }
        }
        // same species: columns q <= row are masked, so the stream starts at the pass's own block
        // and only that block needs the mask
        const int qs = same ? rp : 0;
        // sub-ranges keep the evaluation counter of the histogram-flush logic fine-grained even
        // for absurdly crowded neighbourhoods (normally one range = the whole stream)
        for (int qa = qs; qa < total; qa += (1 << 20)) {
          const int qe = min(total, qa + (1 << 20));
          int q1 = qa;
          if (same && qa == qs) {
            // the pass's own block of the own cell: the only columns that need the col > row
            // mask.  One masked body for every R (idle row slots hold NaN) keeps the code small —
            // the kernel is sensitive to instruction-cache footprint.
            q1 = min(qe, qa + RG * RMAX);
            if (fast) cell_stream<RMAX, true, true, HIST_SMEM>(ws, qa, q1, xi, yi, zi, il);
            else      cell_stream<RMAX, false, true, HIST_SMEM>(ws, qa, q1, xi, yi, zi, il);
          }
          if (q1 < qe) {
            if (fast) {
              switch (R) {
#define MDS_STREAM(RV) cell_stream<RV, true, false, HIST_SMEM>(ws, q1, qe, xi, yi, zi, il)
                case 1: MDS_STREAM(1); break;
                case 2: MDS_STREAM(RMAX >= 2 ? 2 : RMAX); break;
                case 3: MDS_STREAM(RMAX >= 3 ? 3 : RMAX); break;
                case 4: MDS_STREAM(RMAX >= 4 ? 4 : RMAX); break;
                case 5: MDS_STREAM(RMAX >= 5 ? 5 : RMAX); break;
                case 6: MDS_STREAM(RMAX >= 6 ? 6 : RMAX); break;
                case 7: MDS_STREAM(RMAX >= 7 ? 7 : RMAX); break;
                default: MDS_STREAM(RMAX); break;
#undef MDS_STREAM
              }
            } else {
              cell_stream<RMAX, false, false, HIST_SMEM>(ws, q1, qe, xi, yi, zi, il);
            }
          }
          if (HIST_SMEM) {
            unpublished += (uint32_t)(qe - qa) * (uint32_t)nr;
            if (unpublished >= publish_every) {
              uint32_t old = 0;
              if (lane == 0) old = atomicAdd(s_evals, unpublished);
              old = __shfl_sync(0xffffffffu, old, 0);
              // the warp whose contribution crosses the threshold empties the histogram
              if (old < p.flush_threshold && old + unpublished >= p.flush_threshold) {
                if (lane == 0) atomicSub(s_evals, p.flush_threshold);
                flush_hist_warp(s_hist, p, lane MDS_CHK(, s_chk));
              }
              unpublished = 0;
            }
          }
        }
      }
      // pairs distance-tested and not masked by the diagonal
      evaluated += same ? (unsigned long long)n * (unsigned long long)(n - 1) / 2ull +
                              (unsigned long long)n * (unsigned long long)(total - n)
                        : (unsigned long long)n * (unsigned long long)total;
    } while (0);
    idx = __shfl_sync(0xffffffffu, nxt_raw, 0) * p.shard_count + p.shard_index;
  }

  if (lane == 0 && evaluated) atomicAdd(p.evaluated, evaluated);
  if (HIST_SMEM) {
    __syncthreads();
    MDS_CHK(checked_conservation(s_chk, issued, s_hist, p.pair_cnt * hist_stride, tid, CELL_THREADS,
                                 "rdf_cells_kernel");)
    const int stride = p.nbins + 1;
    for (int pr = 0; pr < p.pair_cnt; ++pr) {
      unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
      const uint32_t* h = s_hist + pr * stride;
      for (int k = tid; k < p.nbins; k += CELL_THREADS) {
        const uint32_t v = h[k];
        if (v) atomicAdd(g + k, (unsigned long long)v);
      }
    }
  }
}

// ---- host-side launchers ------------------------------------------------------------------

size_t cells_smem_bytes(int nbins, int pair_cnt, bool hist_in_smem) {
  size_t b = (size_t)CELL_WARPS * 2 * CH * 3 * sizeof(float) +
             (size_t)CELL_WARPS * 2 * MAX_SPANS * sizeof(int) + 8 * sizeof(uint32_t);
  if (hist_in_smem) {
    b += (size_t)((nbins + 2 + 3) & ~3) * sizeof(float);
    b += (size_t)pair_cnt * (size_t)(nbins + 1) * sizeof(uint32_t);
  }
  return b;
}

cudaError_t launch_cell_build(const PackParams& pp, int n_species, float4* sorted,
                              cudaStream_t stream) {
  const long long total = (long long)pp.n_frames * pp.n_packed;
  if (total <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(pp.cell_off, 0,
                                  sizeof(uint32_t) * (size_t)pp.off_stride * (size_t)pp.n_frames,
                                  stream);
  if (e != cudaSuccess) return e;
  e = launch_pack(pp, stream);  // quad-block positions + per-cell counts and ranks + frame flags
  if (e != cudaSuccess) return e;
  rdf_cell_scan_kernel<<<(unsigned)pp.n_frames, 1024, 0, stream>>>(pp.cell_off, pp.off_stride,
                                                                   n_species * pp.grid.nc);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  rdf_cell_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pp, sorted);
  return cudaGetLastError();
}

cudaError_t launch_cells(const CellsParams& p, int sm_count, cudaStream_t stream, int* grid_out) {
  const size_t smem = cells_smem_bytes(p.nbins, p.pair_cnt, p.hist_in_smem != 0);
  auto kern = p.hist_in_smem ? rdf_cells_kernel<true> : rdf_cells_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CELL_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  unsigned long long items =
      (unsigned long long)p.n_frames * (unsigned long long)p.pair_cnt * (unsigned long long)p.nc;
  items = (items + p.shard_count - 1) / p.shard_count;
  unsigned long long grid = (unsigned long long)per_sm * sm_count;  // persistent
  const unsigned long long need = (items + CELL_WARPS - 1) / CELL_WARPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  kern<<<(unsigned)grid, CELL_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mds
