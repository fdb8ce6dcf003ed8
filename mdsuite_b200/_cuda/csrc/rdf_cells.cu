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
//   rdf_cell_pack_kernel   raw xyz -> float4 (x, y, z, rank in cell); per-(frame, species, cell)
//                          counts with one global atomic per atom; frame classification
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

// ---- cell assignment (shared by pack and scatter; exact IEEE ops so it is reproducible) -------
__device__ __forceinline__ int cell_coord(float x, float box, int n) {
  float u = __fdiv_rn(x, box);
  u = __fsub_rn(u, floorf(u));                        // [0, 1]; NaN / inf -> NaN
  const int c = __float2int_rz(__fmul_rn(u, (float)n));  // NaN -> 0
  return min(max(c, 0), n - 1);
}

__device__ __forceinline__ int cell_of(float x, float y, float z, const CellGrid& g) {
  const int cx = cell_coord(x, g.box[0], g.nx);
  const int cy = cell_coord(y, g.box[1], g.ny);
  const int cz = cell_coord(z, g.box[2], g.nz);
  return (cz * g.ny + cy) * g.nx + cx;
}

__global__ void __launch_bounds__(256) rdf_cell_pack_kernel(const PackParams p, const CellGrid g,
                                                            uint32_t* cell_off,
                                                            int64_t off_stride) {
  const long long total = (long long)p.n_frames * p.n_packed;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long f, q;
    if (p.layout == MDS_LAYOUT_ATOM_MAJOR_XYZ) {
      q = t / p.n_frames;
      f = t % p.n_frames;
    } else {
      f = t / p.n_packed;
      q = t % p.n_packed;
    }
    const long long a = p.src_index[q];
    const float* src = (p.layout == MDS_LAYOUT_ATOM_MAJOR_XYZ)
                           ? p.xyz + (a * p.n_frames_total + p.frame0 + f) * 3
                           : p.xyz + ((p.frame0 + f) * p.n_atoms + a) * 3;
    const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
    const int key = p.species[q] * g.nc + cell_of(x, y, z, g);
    const uint32_t rank = atomicAdd(cell_off + f * off_stride + key, 1u);
    p.out[f * p.frame_stride + q] = make_float4(x, y, z, __uint_as_float(rank));
    uint32_t bits = window_bits(x, p.box[0]) | window_bits(y, p.box[1]) | window_bits(z, p.box[2]);
    // bit2: a coordinate further than 8 box lengths out (or infinite): the cell margin no longer
    // covers float32 rounding of the reference's own arithmetic -> all-pairs kernel for this frame
    if (fabsf(x) > 8.0f * p.box[0] || fabsf(y) > 8.0f * p.box[1] || fabsf(z) > 8.0f * p.box[2])
      bits |= MDS_FRAME_FAR;
    if (bits && (p.frame_flags[f] & bits) != bits) atomicOr(p.frame_flags + f, bits);
  }
}

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

__global__ void __launch_bounds__(256) rdf_cell_scatter_kernel(const float4* __restrict__ packed,
                                                               float4* __restrict__ sorted,
                                                               int64_t frame_stride,
                                                               const int32_t* __restrict__ species,
                                                               int32_t n_packed, int64_t n_frames,
                                                               const uint32_t* __restrict__ cell_off,
                                                               int64_t off_stride, const CellGrid g) {
  const long long total = n_frames * n_packed;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long f = t / n_packed;
    const int q = (int)(t % n_packed);
    const float4 v = packed[f * frame_stride + q];
    const int key = species[q] * g.nc + cell_of(v.x, v.y, v.z, g);
    const uint32_t pos = cell_off[f * off_stride + key] + __float_as_uint(v.w);
    sorted[f * frame_stride + pos] = make_float4(v.x, v.y, v.z, __int_as_float(q));
  }
}

// ---- the pair kernel -------------------------------------------------------------------------

constexpr int CELL_THREADS = 256;
constexpr int CELL_WARPS = CELL_THREADS / 32;
constexpr int CH = 32;         // column atoms per staged chunk: one float4 per lane
constexpr int MAX_SPANS = 20;  // >= 18 (9 cell rows x 2 pieces when the x window wraps)
constexpr int RMAX = 4;        // rows per lane: one pass covers 8 * RMAX = 32 rows of a cell

// One chunk of staged columns against the register rows of a warp.  Lane layout: row group
// rg = lane & 7 (rows rg, rg + 8, ...), column phase ph = lane >> 3 (columns ph, ph + 4, ...):
// one LDS.128 with 4 distinct addresses serves R pair evaluations per lane.
template <int R, bool FAST, bool DIAG, bool HIST_SMEM>
__device__ __forceinline__ void cell_chunk(const float4* __restrict__ tile, int cnt, int q0, int ph,
                                           const float (&xi)[RMAX], const float (&yi)[RMAX],
                                           const float (&zi)[RMAX], const int (&il)[RMAX],
                                           float bx, float by, float bz, float s_cut,
                                           float inv_step, uint32_t c_edge, uint32_t d_hist,
                                           const float* __restrict__ edges_g,
                                           unsigned long long* hist_g) {
#pragma unroll 2
  for (int jj = ph; jj < cnt; jj += 4) {
    const float4 pj = tile[jj];
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float dx = __fsub_rn(pj.x, xi[r]);
      const float dy = __fsub_rn(pj.y, yi[r]);
      const float dz = __fsub_rn(pj.z, zi[r]);
      float mx, my, mz;
      if (FAST) {
        mx = minimg_fast(dx, bx);
        my = minimg_fast(dy, by);
        mz = minimg_fast(dz, bz);
      } else {
        mx = minimg_general(dx, bx);
        my = minimg_general(dy, by);
        mz = minimg_general(dz, bz);
      }
      s[r] = sq_norm(mx, my, mz);
    }
    const int q = q0 + jj;  // position in the column stream; the own cell comes first
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (HIST_SMEM) {
        bin_count_dense<DIAG>(s[r], s_cut, inv_step, c_edge, d_hist, q, il[r]);
      } else {
        bool in = s[r] < s_cut;  // false for NaN (padding rows, NaN input)
        if (DIAG) in = in && (q > il[r]);
        if (in) atomicAdd(hist_g + bin_of_s(s[r], inv_step, edges_g), 1ull);
      }
    }
  }
}

// Move what has accumulated in the CTA's shared histogram to the global counts while other warps
// keep counting: atomicExch takes each bin's value and zeroes it in one step, so nothing is lost.
__device__ __forceinline__ void flush_hist_warp(uint32_t* s_hist, const CellsParams& p, int lane) {
  const int stride = p.nbins + 1;
  for (int pr = 0; pr < p.pair_cnt; ++pr) {
    unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
    uint32_t* h = s_hist + pr * stride;
    for (int k = lane; k < p.nbins; k += 32) {
      const uint32_t v = atomicExch(h + k, 0u);
      if (v) atomicAdd(g + k, (unsigned long long)v);
    }
    if (lane == 0) atomicExch(h + p.nbins, 0u);  // the sink may wrap; nobody reads it
  }
}

template <bool HIST_SMEM>
__global__ void __launch_bounds__(CELL_THREADS, 3) rdf_cells_kernel(const CellsParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);                       // [WARPS][2][CH]
  int* span_tab = reinterpret_cast<int*>(tiles + CELL_WARPS * 2 * CH);       // [WARPS][2][MAX_SPANS]
  uint32_t* s_evals = reinterpret_cast<uint32_t*>(span_tab + CELL_WARPS * 2 * MAX_SPANS);  // [4]
  float* s_edges = reinterpret_cast<float*>(s_evals + 4);
  const int n_edges = p.nbins + 2;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_edges + ((n_edges + 3) & ~3));
  const int hist_stride = p.nbins + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = lane & 7, ph = lane >> 3;

  if (HIST_SMEM) {
    for (int k = tid; k < n_edges; k += CELL_THREADS) s_edges[k] = p.edges[k];
    for (int k = tid; k < p.pair_cnt * hist_stride; k += CELL_THREADS) s_hist[k] = 0u;
  }
  if (tid == 0) s_evals[0] = 0u;
  __syncthreads();

  float4* tile = tiles + warp * 2 * CH;
  int* t_qend = span_tab + warp * 2 * MAX_SPANS;
  int* t_off = t_qend + MAX_SPANS;
  const uint32_t edges1_sa = smem_u32(s_edges + 1);
  const uint32_t c_edge = edges1_sa - kMagic4;
  const float bx = p.box[0], by = p.box[1], bz = p.box[2];
  const float s_cut = p.s_cut, inv_step = p.inv_step;
  const int nx = p.nx, ny = p.ny, nz = p.nz, nc = p.nc;
  const unsigned long long total_items =
      (unsigned long long)p.n_frames * (unsigned long long)p.pair_cnt * (unsigned long long)nc;
  const uint32_t publish_every = max(1u, p.flush_threshold >> 6);
  unsigned long long evaluated = 0;  // warp-uniform statistics
  uint32_t unpublished = 0;          // candidate pairs since this warp last told the CTA

  unsigned long long nxt_raw = 0;
  if (lane == 0) nxt_raw = atomicAdd(p.work_counter, 1ull);
  unsigned long long idx = __shfl_sync(0xffffffffu, nxt_raw, 0);

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

      const uint32_t d_hist = smem_u32(s_hist + pr * hist_stride) - edges1_sa;
      unsigned long long* hist_g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;

      // ---- row passes of up to 32 atoms of the cell ----
      for (int rp = 0; rp < n; rp += 8 * RMAX) {
        const int nr = min(8 * RMAX, n - rp);
        const int R = fast ? (nr + 7) >> 3 : RMAX;
        float xi[RMAX], yi[RMAX], zi[RMAX];
        int il[RMAX];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          il[r] = rp + rg + 8 * r;
          if (il[r] < n) {
            const float4 v = __ldg(fpos + r0 + il[r]);
            xi[r] = v.x; yi[r] = v.y; zi[r] = v.z;
          } else {
            xi[r] = yi[r] = zi[r] = __int_as_float(0x7fc00000);  // NaN: never in cutoff
          }
        }
        // same species: columns q <= row are masked, so the stream starts at the pass's own block
        const int qs = same ? rp : 0;
        int k = 0;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        {
          const int q = qs + lane;
          if (q < total) {
            while (q >= t_qend[k]) ++k;
            nxt = __ldg(fpos + t_off[k] + q);
          }
        }
        __syncwarp();
        tile[lane] = nxt;
        __syncwarp();
        int st = 0;
        for (int q0 = qs; q0 < total; q0 += CH, st ^= 1) {
          const int cnt = min(CH, total - q0);
          const int qn = q0 + CH + lane;
          const bool have = qn < total;
          if (have) {
            while (qn >= t_qend[k]) ++k;
            nxt = __ldg(fpos + t_off[k] + qn);
          }
          const float4* tl = tile + st * CH;
          const bool diag = same && (q0 < rp + 8 * RMAX);
#define MDS_CHUNK(RV, FASTV, DIAGV)                                                             \
  cell_chunk<RV, FASTV, DIAGV, HIST_SMEM>(tl, cnt, q0, ph, xi, yi, zi, il, bx, by, bz, s_cut,       \
                                          inv_step, c_edge, d_hist, p.edges, hist_g)
#define MDS_CHUNK_R(RV)                                                                         \
  do { if (diag) MDS_CHUNK(RV, true, true); else MDS_CHUNK(RV, true, false); } while (0)
          if (fast) {
            switch (R) {
              case 1: MDS_CHUNK_R(1); break;
              case 2: MDS_CHUNK_R(2); break;
              case 3: MDS_CHUNK_R(3); break;
              default: MDS_CHUNK_R(4); break;
            }
          } else {
            if (diag) MDS_CHUNK(RMAX, false, true); else MDS_CHUNK(RMAX, false, false);
          }
#undef MDS_CHUNK_R
#undef MDS_CHUNK
          if (have) tile[(st ^ 1) * CH + lane] = nxt;
          __syncwarp();
          if (HIST_SMEM) {
            unpublished += (uint32_t)(cnt * nr);
            if (unpublished >= publish_every) {
              uint32_t old = 0;
              if (lane == 0) old = atomicAdd(s_evals, unpublished);
              old = __shfl_sync(0xffffffffu, old, 0);
              // the warp whose contribution crosses the threshold empties the histogram
              if (old < p.flush_threshold && old + unpublished >= p.flush_threshold) {
                if (lane == 0) atomicSub(s_evals, p.flush_threshold);
                flush_hist_warp(s_hist, p, lane);
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
    idx = __shfl_sync(0xffffffffu, nxt_raw, 0);
  }

  if (lane == 0 && evaluated) atomicAdd(p.evaluated, evaluated);
  if (HIST_SMEM) {
    __syncthreads();
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
  size_t b = (size_t)CELL_WARPS * 2 * CH * sizeof(float4) +
             (size_t)CELL_WARPS * 2 * MAX_SPANS * sizeof(int) + 4 * sizeof(uint32_t);
  if (hist_in_smem) {
    b += (size_t)((nbins + 2 + 3) & ~3) * sizeof(float);
    b += (size_t)pair_cnt * (size_t)(nbins + 1) * sizeof(uint32_t);
  }
  return b;
}

cudaError_t launch_cell_build(const PackParams& pp, const CellGrid& g, int n_species,
                              uint32_t* cell_off, int64_t off_stride, float4* sorted,
                              cudaStream_t stream) {
  const long long total = (long long)pp.n_frames * pp.n_packed;
  if (total <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(cell_off, 0, sizeof(uint32_t) * (size_t)off_stride * (size_t)pp.n_frames,
                                  stream);
  if (e != cudaSuccess) return e;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  rdf_cell_pack_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pp, g, cell_off, off_stride);
  rdf_cell_scan_kernel<<<(unsigned)pp.n_frames, 1024, 0, stream>>>(cell_off, off_stride,
                                                                   n_species * g.nc);
  rdf_cell_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      pp.out, sorted, pp.frame_stride, pp.species, pp.n_packed, pp.n_frames, cell_off, off_stride, g);
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
  const unsigned long long items =
      (unsigned long long)p.n_frames * (unsigned long long)p.pair_cnt * (unsigned long long)p.nc;
  unsigned long long grid = (unsigned long long)per_sm * sm_count;  // persistent
  const unsigned long long need = (items + CELL_WARPS - 1) / CELL_WARPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  kern<<<(unsigned)grid, CELL_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mds
