// Cell-list RDF path for sm_100a (north-star kernel (b)): large N, cutoff much smaller than the
// box.  The reference has no such path — it evaluates all N(N-1)/2 pairs of every frame
// (SURVEY.md §5 "Long-context"); this path evaluates only pairs of atoms in nearby cells and
// must return the SAME integer histogram, bit for bit.
//
// Why it is exact: a cell list only decides WHICH pairs are distance-tested.  Every tested pair
// runs the reference's float32 arithmetic on the raw coordinates (rdf_common.cuh; reference
// utils/linalg.py:84-99, calculators/radial_distribution_function.py:667-689), and the squared
// distance is symmetric in (i, j) bit for bit.  The grid is conservative: K cells span at least
// 1.001 x cutoff in every dimension (K = 1 or 2, the stencil half-width), so for |coordinate| <=
// 8 box lengths every pair the reference would bin sits in cells at most K apart (DESIGN.md §9
// derives the margin).  Frames with coordinates further out are flagged by the pack kernel and
// handed to the all-pairs kernel instead.
//
// Per batch of frames, either
//   rdf_cell_build_kernel  a CTA per frame does the whole build (counts in shared memory, scan,
//                          placement); see "the whole build of a frame in one CTA" below
// or (atom-major input, key tables beyond shared memory, batches shorter than a wave of CTAs)
//   rdf_pack_kernel        (rdf_pack.cu) raw xyz -> quad-block positions; per-(frame, species,
//                          cell) counts and ranks with one global atomic per atom; frame flags
//   rdf_cell_scan_kernel   exclusive scan of the counts of a frame -> cell offsets (one CTA/frame);
//                          flags frames with an absurdly crowded cell for the all-pairs kernel
//   rdf_cell_scatter_kernel counting-sort scatter -> atoms ordered by (species, cell), x fastest
// and then
//   rdf_cells_kernel       persistent CTAs; one work item = (frame, species pair, pencil of cells
//                          along x).  See "the pair kernel" below.
#include "rdf_common.cuh"
#include "rdf_launch.h"
#include "mds_rdf.h"

namespace mds {

// Exclusive scan of the n_keys counts of one frame (in place) + total at [n_keys].  Also the
// largest count of the frame: above `crowded_limit` the pencil kernel could not stage even one
// row cell with one neighbour pencil, so the frame goes to the all-pairs kernel (MDS_FRAME_FAR).
__global__ void __launch_bounds__(1024) rdf_cell_scan_kernel(uint32_t* cell_off, int64_t off_stride,
                                                             int n_keys, uint32_t* frame_flags,
                                                             uint32_t crowded_limit) {
  uint32_t* a = cell_off + (long long)blockIdx.x * off_stride;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t chunk_total;
  __shared__ uint32_t s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_max = 0u;
  uint32_t carry = 0, mx = 0;
  for (int base = 0; base < n_keys; base += 4096) {
    const int i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (i0 + k < n_keys) ? a[i0 + k] : 0u;
      mx = max(mx, v[k]);
    }
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
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0 && mx > crowded_limit) atomicMax(&s_max, mx);
  __syncthreads();
  if (threadIdx.x == 0) {
    a[n_keys] = carry;
    if (s_max > crowded_limit) atomicOr(frame_flags + blockIdx.x, MDS_FRAME_FAR);
  }
}

// The same scan for long key tables (a million-atom frame has > 10^5 keys, and a single frame —
// item-sharded over the GPUs — has nothing else to hide one serial CTA behind): pass 1 scans
// blocks of 4096 keys in place and leaves their totals, pass 2 adds the totals of the blocks
// before it.  grid = (blocks per frame, frames).
__global__ void __launch_bounds__(1024) rdf_cell_scan_blocks_kernel(uint32_t* cell_off,
                                                                    int64_t off_stride, int n_keys,
                                                                    uint32_t* block_sums,
                                                                    uint32_t* frame_max) {
  uint32_t* a = cell_off + (long long)blockIdx.y * off_stride;
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 4096 + threadIdx.x * 4;
  uint32_t v[4], mx = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = (i0 + k < n_keys) ? a[i0 + k] : 0u;
    mx = max(mx, v[k]);
  }
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
    if (lane == 31) block_sums[(long long)blockIdx.y * gridDim.x + blockIdx.x] = winc;
  }
  __syncthreads();
  uint32_t e = warp_sums[warp] + inc - t;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i0 + k < n_keys) a[i0 + k] = e;
    e += v[k];
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0 && mx) atomicMax(frame_max + blockIdx.y, mx);
}

__global__ void __launch_bounds__(1024) rdf_cell_scan_finish_kernel(uint32_t* cell_off,
                                                                    int64_t off_stride, int n_keys,
                                                                    const uint32_t* block_sums,
                                                                    uint32_t* frame_max,
                                                                    uint32_t* frame_flags,
                                                                    uint32_t crowded_limit) {
  uint32_t* a = cell_off + (long long)blockIdx.y * off_stride;
  const uint32_t* sums = block_sums + (long long)blockIdx.y * gridDim.x;
  __shared__ uint32_t s_before;
  if (threadIdx.x < 32) {  // totals of the blocks before this one (and of all, for the last)
    uint32_t acc = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) acc += sums[b];
    acc = __reduce_add_sync(0xffffffffu, acc);
    if (threadIdx.x == 0) s_before = acc;
  }
  __syncthreads();
  const uint32_t before = s_before;
  const int i0 = blockIdx.x * 4096 + threadIdx.x * 4;
  if (before) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k < n_keys) a[i0 + k] += before;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    a[n_keys] = before + sums[blockIdx.x];
    if (frame_max[blockIdx.y] > crowded_limit) atomicOr(frame_flags + blockIdx.y, MDS_FRAME_FAR);
    frame_max[blockIdx.y] = 0u;  // ready for the next batch
  }
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
    const int cell = cell_of(x, y, z, p.grid);
    if (!cell_in_slab(cell, p.grid)) continue;  // item shards: not in this handle's slab
    const int key = p.species[q] * p.grid.nc + cell;
    const uint32_t pos = p.cell_off[f * p.off_stride + key] + p.rank[f * p.frame_stride + q];
    sorted[f * p.frame_stride + pos] = make_float4(x, y, z, __int_as_float(q));
  }
}

// ---- the whole build of a frame in one CTA ---------------------------------------------------
//
// The three kernels above move every atom through global memory three times (pack: raw ->
// quad blocks + rank, one global atomic per atom; scatter: quad blocks + rank -> sorted) and are
// bound by the global atomics, not by HBM.  When the key table of a frame fits in shared memory
// and the input is frame-major (a frame's atoms are contiguous), one CTA does all of it for one
// frame: count into a shared-memory table (ATOMS.POPC.INC), scan the table in place, write the
// offsets, then read the frame a second time and place every atom with a returning shared atomic
// on the running offsets — no rank array, no global atomics except the frame's flag word.
// Order inside a cell is arbitrary, as it was (the histogram is a sum over pairs).
// The quad-block copy of the frame is only needed by the all-pairs kernel, for frames flagged
// MDS_FRAME_FAR: normally (PACKED = false) the kernel does not write it, pass 2 reads the raw frame
// again, and the caller packs the flagged frames of a batch afterwards (launch_pack with a gate).
#ifndef MDS_BUILD_THREADS
#define MDS_BUILD_THREADS 1024
#endif
#ifndef MDS_BUILD_MINB
#define MDS_BUILD_MINB 1
#endif
#ifndef MDS_BUILD_UNROLL
#define MDS_BUILD_UNROLL 4
#endif
constexpr int BUILD_THREADS = MDS_BUILD_THREADS;  // a multiple of 128
constexpr int BUILD_MAX_KEYS = 48 * 1024;  // 192 KB of counters

constexpr int BUILD_UNROLL = MDS_BUILD_UNROLL;  // atoms per thread whose loads are in flight together

// One CTA works on one frame, so the time of a batch is the time of a thread's loop: the loads of
// BUILD_UNROLL atoms (slots q0 + u * BUILD_THREADS) are issued back to back, branch-free, and the
// per-atom instruction count is kept down — the frame classification is taken from per-thread
// minima / maxima instead of eight comparisons per atom, and pass 2 reads the key pass 1 computed
// (three IEEE divisions per atom) instead of computing it again.
template <bool PACKED>  // PACKED: also write the quad-block copy of the frame (see above)
__global__ void __launch_bounds__(BUILD_THREADS, MDS_BUILD_MINB) rdf_cell_build_kernel(const PackParams p,
                                                                          float4* __restrict__ sorted,
                                                                          int n_keys,
                                                                          uint32_t crowded_limit) {
  extern __shared__ uint32_t s_key[];  // [n_keys] counts, then running offsets; [n_keys]: sink
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t chunk_total, s_bits, s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float nan = __int_as_float(0x7fc00000);
  const float inf = __int_as_float(0x7f800000);
  const bool one_species = n_keys == p.grid.nc;
  for (long long f = blockIdx.x; f < p.n_frames; f += gridDim.x) {
    for (int i = threadIdx.x; i < n_keys; i += BUILD_THREADS) s_key[i] = 0u;
    if (threadIdx.x == 0) { s_bits = 0u; s_max = 0u; }
    __syncthreads();
    const float* src_f = p.xyz + (p.frame0 + f) * p.n_atoms * 3;
    float* out_f = p.out + f * p.frame_stride * 3;
    int* key_f = reinterpret_cast<int*>(p.rank) + f * p.frame_stride;  // key of every slot, -1: none

    // pass 1: quad-block copy, frame classification, keys and counts
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};  // fminf / fmaxf skip NaN
    for (int q0 = threadIdx.x; q0 < p.n_packed; q0 += BUILD_UNROLL * BUILD_THREADS) {
      int idx[BUILD_UNROLL];  // raw atom index, -1: padding slot, -2: beyond the end of the frame
      float x[BUILD_UNROLL], y[BUILD_UNROLL], z[BUILD_UNROLL];
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u) {
        const int q = q0 + u * BUILD_THREADS;
        idx[u] = (q < p.n_packed) ? __ldg(p.src_index + q) : -2;
      }
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u) {
        const float* src = src_f + max(idx[u], 0) * 3;
        x[u] = __ldg(src); y[u] = __ldg(src + 1); z[u] = __ldg(src + 2);
      }
      // (BUILD_THREADS is a multiple of 4: slot q0 + u * BUILD_THREADS sits 3 * u * BUILD_THREADS
      // floats after slot q0 in the quad-block array)
      float* dst0 = out_f + ((q0 & ~3) * 3 + (q0 & 3));
      int* key0 = key_f + q0;
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u) {
        if (idx[u] == -2) continue;
        float* dst = dst0 + u * (3 * BUILD_THREADS);
        if (idx[u] < 0) {  // padding slot at the end of a species segment
          if (PACKED) { dst[0] = nan; dst[4] = nan; dst[8] = nan; }
          key0[u * BUILD_THREADS] = -1;
          continue;
        }
        if (PACKED) { dst[0] = x[u]; dst[4] = y[u]; dst[8] = z[u]; }
        lo[0] = fminf(lo[0], x[u]); hi[0] = fmaxf(hi[0], x[u]);
        lo[1] = fminf(lo[1], y[u]); hi[1] = fmaxf(hi[1], y[u]);
        lo[2] = fminf(lo[2], z[u]); hi[2] = fmaxf(hi[2], z[u]);
        const int cell = cell_of(x[u], y[u], z[u], p.grid);
        // (item shards: atoms outside this handle's slab of z layers are not part of its lists)
        int key = -1;
        if (cell_in_slab(cell, p.grid)) {
          key = one_species ? cell : __ldg(p.species + q0 + u * BUILD_THREADS) * p.grid.nc + cell;
          atomicAdd(&s_key[key], 1u);
        }
        key0[u * BUILD_THREADS] = key;
      }
    }
    // the flags rdf_pack_kernel derives atom by atom (window_bits, MDS_FRAME_FAR), from the extrema:
    // "some coordinate below a" is "the minimum is below a", and NaN coordinates set no bit there
    uint32_t bits = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float box = p.box[d];  // (a thread without atoms keeps lo = +inf, hi = -inf: no bit)
      if (lo[d] < -0.125f * box || hi[d] > 1.125f * box) bits |= MDS_FRAME_OUT_A;
      if (lo[d] < -0.625f * box || hi[d] > 0.625f * box) bits |= MDS_FRAME_OUT_B;
      if (lo[d] < -(8.0f * box) || hi[d] > 8.0f * box) bits |= MDS_FRAME_FAR;
    }
    bits = __reduce_or_sync(0xffffffffu, bits);
    if (lane == 0 && bits) atomicOr(&s_bits, bits);
    __syncthreads();

    // exclusive scan of the table in place; the offsets also go to global memory for the pair kernel
    uint32_t* off_f = p.cell_off + f * p.off_stride;
    uint32_t carry = 0, mx = 0;
    for (int base = 0; base < n_keys; base += 4 * BUILD_THREADS) {
      const int i0 = base + threadIdx.x * 4;
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = (i0 + k < n_keys) ? s_key[i0 + k] : 0u;
        mx = max(mx, v[k]);
      }
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
        const uint32_t w = lane < BUILD_THREADS / 32 ? warp_sums[lane] : 0u;
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
        if (i0 + k < n_keys) {
          s_key[i0 + k] = e;
          off_f[i0 + k] = e;
        }
        e += v[k];
      }
      carry += chunk_total;
      __syncthreads();
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0 && mx > crowded_limit) atomicMax(&s_max, mx);
    __syncthreads();  // (also: this CTA's quad-block copy and keys are visible to all its threads)
    if (threadIdx.x == 0) {
      off_f[n_keys] = carry;
      const uint32_t flags = s_bits | (s_max > crowded_limit ? MDS_FRAME_FAR : 0u);
      if (flags) atomicOr(p.frame_flags + f, flags);
    }

    // pass 2: place the atoms — key and coordinates of a slot as pass 1 left them, the position
    // from the running offset of the key
    float4* sorted_f = sorted + f * p.frame_stride;
    for (int q0 = threadIdx.x; q0 < p.n_packed; q0 += BUILD_UNROLL * BUILD_THREADS) {
      int key[BUILD_UNROLL];
      float x[BUILD_UNROLL], y[BUILD_UNROLL], z[BUILD_UNROLL];
      if (PACKED) {
        const float* src0 = out_f + ((q0 & ~3) * 3 + (q0 & 3));
#pragma unroll
        for (int u = 0; u < BUILD_UNROLL; ++u) {
          const bool in = q0 + u * BUILD_THREADS < p.n_packed;
          key[u] = in ? key_f[q0 + u * BUILD_THREADS] : -1;
          const float* src = in ? src0 + u * (3 * BUILD_THREADS) : src0;
          x[u] = src[0]; y[u] = src[4]; z[u] = src[8];
        }
      } else {
        int idx[BUILD_UNROLL];
#pragma unroll
        for (int u = 0; u < BUILD_UNROLL; ++u) {
          const bool in = q0 + u * BUILD_THREADS < p.n_packed;
          key[u] = in ? key_f[q0 + u * BUILD_THREADS] : -1;
          idx[u] = in ? __ldg(p.src_index + q0 + u * BUILD_THREADS) : 0;
        }
#pragma unroll
        for (int u = 0; u < BUILD_UNROLL; ++u) {
          const float* src = src_f + max(idx[u], 0) * 3;
          x[u] = __ldg(src); y[u] = __ldg(src + 1); z[u] = __ldg(src + 2);
        }
      }
      // the returning atomics of the unrolled slots are issued together, then the stores
      uint32_t pos[BUILD_UNROLL];
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u)
        pos[u] = atomicAdd(&s_key[key[u] >= 0 ? key[u] : n_keys], key[u] >= 0 ? 1u : 0u);
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u)
        if (key[u] >= 0)
          sorted_f[pos[u]] = make_float4(x[u], y[u], z[u], __int_as_float(q0 + u * BUILD_THREADS));
    }
    __syncthreads();  // the table is zeroed again for the next frame
  }
}

// ---- the pair kernel -------------------------------------------------------------------------
//
// Geometry.  Cells are at least 1.001 cutoff / K wide (K = kx, ky, kz per dimension), so the
// partners of an atom lie in the (2kx+1)(2ky+1)(2kz+1) cells around its own.  With K = 2 that
// block has 125/8 = 15.6 times the volume of a cutoff-sized cell instead of the 27 of the
// classical K = 1 scheme: 42 % fewer candidate pairs for the same binned pairs.  Small cells only
// pay if few atoms can share one column set efficiently, which decides the layout:
//
// * One work item is a PENCIL: the nx cells with the same (y, z).  The atoms of a pencil are
//   contiguous in the sorted array (x fastest), and so are those of each of the (2ky+1)(2kz+1)
//   pencils around it (same-species pairs: the half shell, kz (2ky+1) + ky pencils, plus the
//   pencil itself for the +x half).
// * The CTA stages those pencils ONCE in shared memory, re-ordered slab by slab: slab c holds the
//   atoms of x-cell c of all neighbour pencils.  The column set of row cell cx is then ONE
//   contiguous range, slabs cx-kx .. cx+kx (ghost slabs at both ends take care of the periodic
//   wrap), found with two table reads.  Every atom is read from global memory once per
//   neighbouring pencil (25 times per frame for K = 2) instead of once per neighbouring cell.
// * Columns sit on lanes, rows are broadcast: a warp takes one row cell, each lane loads two
//   columns of the range into registers (64 per chunk) and loops over the n atoms of the cell,
//   one LDS.128 per row with every lane on the same address; each (row, two columns) goes through
//   the packed FADD2 / FMUL2 sequence.  A row cell with 3 atoms keeps all 32 lanes busy, which a
//   rows-in-registers layout cannot do.
// * Staging is a table of (slab, pencil) entries: counts from the cell-offset table, one
//   block-wide exclusive scan, then 16-byte cp.async copies that are all in flight at once.
// * Same-species items: the own cell's atoms come first in the column range; the col > row mask
//   of that block is applied by poisoning column r with NaN when row r is reached (rows are
//   visited in increasing order, and column r is not needed by any later row).
// * Binning as in the all-pairs kernel: threshold table + ATOMS.POPC.INC, branch-free, sink slot.
//
// Shared memory per CTA is fixed (cap_atoms staged positions); a pencil is processed in x
// segments of at most `seg` row cells (chosen per species pair at handle creation from the
// average density) and, where a frame is denser than average, in fewer cells or with a subset of
// the neighbour pencils per pass — the kernel measures every candidate segment against the
// capacity before staging it, so no density distribution can overflow the buffers.

#ifndef CELL_UNROLL
#define CELL_UNROLL 8  // row pairs per pass of the inner loop (measured on cfg3: 4 beats 2 by 3-4 %, 8 beats 4 by 1.4 %, 6 = 4)
#endif
#ifndef CELL_MINB
#define CELL_MINB 4    // resident CTAs per SM the register budget is set for
#endif
constexpr int kCellUnroll = CELL_UNROLL;
constexpr int CELL_THREADS = 256;
constexpr int CELL_WARPS = CELL_THREADS / 32;
constexpr int MAXW = 72;        // x cells of a staged segment incl. ghost cells (seg <= 64, kx <= 4)
constexpr int MAX_ENT = 1024;   // (slab, pencil) entries of a staged segment
constexpr int EPT = MAX_ENT / CELL_THREADS;  // entries per thread in the scan
// fixed-size tables at the start of the CTA's shared memory (E, Esrc, ro, ao, s_pen, misc)
constexpr int MAX_CAP = 4096;                    // staged atoms per CTA, at most
constexpr int MAX_UNITS = 2 * (MAX_CAP / 32) + 2;  // 32-column chunks of a staged segment
constexpr int ROWS_CAP = 512;                    // row atoms of a staged segment (+ pair padding)
constexpr int CELLS_FIXED_BYTES =
    ((2 * MAX_ENT + 2 * (MAXW + 1) + 32 + 24 + 2 * MAX_UNITS) * 4 + 127) & ~127;

struct BinConsts {
  float bx, by, bz, s_cut, inv_step;
  uint32_t c_edge, c_hist, four;
  const float* edges_g;
  unsigned long long* hist_g;
};

// One chunk of 32 columns (one per lane, in registers) against n2 consecutive PAIRS of row atoms
// (structure-of-arrays in shared memory, pair-aligned): three LDS.64 with every lane on the same
// address give the register pairs of the packed FADD2 / FMUL2 sequence, two rows per pass.
// MASK: the column only counts against rows [lo, lo + len) of the 2 n2 (same-species chunks of the
// own pencil: rows before the column itself; chunks whose rows could alias a periodic image).
// INTER: the CTA's threshold table and histogram are one interleaved array (rdf_common.cuh,
// bin_count_dense_interleaved; launches with one species pair), k.c_edge / k.four then hold the
// slot constants.
template <bool FAST, bool MASK, bool HIST_SMEM, bool INTER>
__device__ __forceinline__ void rows_vs_column(const float* __restrict__ rx,
                                               const float* __restrict__ ry,
                                               const float* __restrict__ rz, int n2, float4 c, int lo,
                                               uint32_t len, const BinConsts& k,
                                               unsigned long long& issued) {
  (void)issued;
  float hi = 0.f;
  const float nan = __int_as_float(0x7fc00000);
  const f32x2 xj = pack2(c.x, c.x), yj = pack2(c.y, c.y), zj = pack2(c.z, c.z);
#pragma unroll kCellUnroll
  for (int r = 0; r < n2; ++r) {
    f32x2 xi = *reinterpret_cast<const f32x2*>(rx + 2 * r);
    const f32x2 yi = *reinterpret_cast<const f32x2*>(ry + 2 * r);
    const f32x2 zi = *reinterpret_cast<const f32x2*>(rz + 2 * r);
    if (MASK) {
      float x0, x1;
      unpack2(xi, x0, x1);
      x0 = ((uint32_t)(2 * r - lo) < len) ? x0 : nan;
      x1 = ((uint32_t)(2 * r + 1 - lo) < len) ? x1 : nan;
      xi = pack2(x0, x1);
    }
    float s[2];
    if (FAST) {
      pair2_fast_packed(xi, yi, zi, xj, yj, zj, k.bx, k.by, k.bz, s[0], s[1]);
    } else {
      float x0, x1, y0, y1, z0, z1;
      unpack2(xi, x0, x1); unpack2(yi, y0, y1); unpack2(zi, z0, z1);
      pair2_s<false>(x0, x1, y0, y1, z0, z1, c.x, c.y, c.z, k.bx, k.by, k.bz, s[0], s[1]);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (HIST_SMEM && INTER) {
        bin_count_dense_interleaved(s[q], k.s_cut, k.inv_step, k.c_edge, k.four);
        MDS_CHK(issued += 1;)
      } else if (HIST_SMEM) {
        bin_count_dense<false, false>(s[q], k.s_cut, k.inv_step, 0.f, k.c_edge, k.c_hist, k.four, 0,
                                      0, hi);
        MDS_CHK(issued += 1;)
      } else {
        if (s[q] < k.s_cut)  // false for NaN (slots past the end, masked rows, NaN input)
          atomicAdd(k.hist_g + bin_of_s(s[q], k.inv_step, k.edges_g), 1ull);
      }
    }
  }
}

// Largest c in [0, n) with t[c * stride] <= q, for a non-decreasing table with t[0] = 0 and
// q < (the total, which is not stored when stride > 1): the cell / slab that holds position q.
__device__ __forceinline__ int find_cell(const int* t, int stride, int n, int q) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t[mid * stride] <= q) lo = mid; else hi = mid;
  }
  return lo;
}

// Inclusive scan in place of t[1 .. n] (t[0] = 0 set by the caller), by one warp.
__device__ __forceinline__ void warp_scan_table(int* t, int n, int lane) {
  int carry = 0;
  for (int base = 1; base <= n; base += 32) {
    const int i = base + lane;
    int v = (i <= n) ? t[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (i <= n) t[i] = v + carry;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
}

// 16-byte asynchronous global -> shared copy (LDGSTS): the staging copies of a thread are all in
// flight at once instead of one round trip per atom.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// Move what has accumulated in the CTA's shared histogram to the global counts while other warps
// keep counting: atomicExch takes each bin's value and zeroes it in one step, so nothing is lost.
// hs: words between consecutive bins (2 for the interleaved table, where pair_cnt is 1)
__device__ __forceinline__ void flush_hist_warp(uint32_t* s_hist, int hs, const CellsParams& p, int lane,
                                                unsigned long long* s_chk = nullptr) {
  (void)s_chk;
  const int stride = p.nbins + 1;
  for (int pr = 0; pr < p.pair_cnt; ++pr) {
    unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
    uint32_t* h = s_hist + pr * stride;
    for (int k = lane; k < p.nbins; k += 32) {
      const uint32_t v = atomicExch(h + k * hs, 0u);
      if (v) {
        atomicAdd(g + k, (unsigned long long)v);
        MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)v);)
      }
    }
    if (lane == 0) {  // the sink may wrap; nobody reads it
      const uint32_t v = atomicExch(h + p.nbins * hs, 0u);
      (void)v;
      MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)v);)
    }
  }
}

// misc slots of the CTA
enum { M_EVALS = 0, M_CEDGE = 1, M_HIST0 = 2, M_FOUR = 4, M_FRAME = 6, M_PEN = 7, M_PR = 8,
       M_NEXT_CELL = 9, M_SEG = 10, M_WSUM = 16, M_MISC_WORDS = 24 };

// 128-bit load from a shared-memory byte address
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}

template <bool HIST_SMEM, bool INTER>
__global__ void __launch_bounds__(CELL_THREADS, CELL_MINB) rdf_cells_kernel(const CellsParams p) {
  // fixed-size tables first (constant offsets), then the staged atoms, then table + histograms
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* E = reinterpret_cast<int*>(smem_raw);           // [MAX_ENT] atoms per (slab, pencil) -> offsets
  int* Esrc = E + MAX_ENT;                             // [MAX_ENT] first sorted index of the entry
  int* ro = Esrc + MAX_ENT;                            // [MAXW + 1] row-pencil offsets per x cell
  int* s_pen = ro + 2 * (MAXW + 1);                        // [32] offset-table position of neighbour pencil q
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_pen + 32);  // [M_MISC_WORDS]
  int2* s_unit = reinterpret_cast<int2*>(s_misc + M_MISC_WORDS);  // [MAX_UNITS] (first row, rows | masked << 30)
  float* RX = reinterpret_cast<float*>(smem_raw + CELLS_FIXED_BYTES);  // [3][ROWS_CAP] row atoms, SoA
  float* RY = RX + ROWS_CAP;
  float* RZ = RY + ROWS_CAP;
  float4* S = reinterpret_cast<float4*>(RZ + ROWS_CAP);  // [cap] the rows' pencil, then the slabs
  float* s_edges = reinterpret_cast<float*>(S + p.cap_atoms);
  const int n_edges = p.nbins + 2;
  // INTER (pair_cnt == 1): the same bytes hold nbins + 1 slots {edges[k + 1], hist[k]}
  uint32_t* s_hist = INTER ? reinterpret_cast<uint32_t*>(s_edges) + 1
                           : reinterpret_cast<uint32_t*>(s_edges + ((n_edges + 3) & ~3));
  constexpr int hs = INTER ? 2 : 1;  // words between consecutive bins
  const int hist_stride = p.nbins + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long issued = 0;  // only counted in checked builds
  MDS_CHK(__shared__ unsigned long long s_chk[2]; if (tid == 0) { s_chk[0] = 0; s_chk[1] = 0; })

  const int nx = p.nx, ny = p.ny, nc = p.nc;
  const int kx = p.kx, ky = p.ky, kz = p.kz;
  const int npen = p.n_pen;  // pencils of this handle per frame (all ny * nz without item shards)
  const int wy = 2 * ky + 1;
  const unsigned long long total_items = (unsigned long long)p.n_frames *
                                         (unsigned long long)p.pair_cnt *
                                         (unsigned long long)npen * (unsigned long long)p.n_seg;
  // thread 0: take the next work item off the global counter and leave it decoded in the slots
  auto fetch_item = [&]() {
    const unsigned long long idx = atomicAdd(p.work_counter, 1ull) * p.shard_count + p.shard_index;
    int frame = -1, pen = 0, pr = 0, seg = 0;
    if (idx < total_items) {
      // consecutive items: the segments of a pencil, then the next pencil of the frame
      seg = (int)(idx % (unsigned long long)p.n_seg);
      const unsigned long long u = idx / (unsigned long long)p.n_seg;
      pen = p.pen_lo + (int)(u % (unsigned long long)npen);
      const unsigned long long t = u / (unsigned long long)npen;
      pr = (int)(t % (unsigned long long)p.pair_cnt);
      frame = (int)(t / (unsigned long long)p.pair_cnt);
    }
    s_misc[M_FRAME] = (uint32_t)frame;
    s_misc[M_PEN] = (uint32_t)pen;
    s_misc[M_PR] = (uint32_t)pr;
    s_misc[M_SEG] = (uint32_t)seg;
  };

  if (HIST_SMEM && INTER) {
    for (int k = tid; k < hist_stride; k += CELL_THREADS) {
      s_edges[2 * k] = p.edges[k + 1];
      s_hist[2 * k] = 0u;
    }
  } else if (HIST_SMEM) {
    for (int k = tid; k < n_edges; k += CELL_THREADS) s_edges[k] = p.edges[k];
    for (int k = tid; k < p.pair_cnt * hist_stride; k += CELL_THREADS) s_hist[k] = 0u;
  }
  if (tid == 0) {
    s_misc[M_EVALS] = 0u;
    // address constants of the binning sequence, made opaque to ptxas (see rdf_allpairs.cu)
    s_misc[M_CEDGE] = INTER ? smem_u32(s_edges) - kMagic8 : smem_u32(s_edges + 1) - kMagic4;
    s_misc[M_HIST0] = smem_u32(s_hist) - kMagic4;
    s_misc[M_FOUR] = INTER ? 8u : 4u;
    s_misc[M_FOUR + 1] = INTER ? 8u : 4u;
    fetch_item();
  }
  __syncthreads();

  BinConsts bc;
  bc.c_edge = *reinterpret_cast<volatile uint32_t*>(s_misc + M_CEDGE);
  const uint32_t hist0 = *reinterpret_cast<volatile uint32_t*>(s_misc + M_HIST0);
  bc.four = *reinterpret_cast<volatile uint32_t*>(s_misc + M_FOUR + (lane & 1));
  bc.bx = p.box[0]; bc.by = p.box[1]; bc.bz = p.box[2];
  bc.s_cut = p.s_cut;
  bc.inv_step = p.inv_step;
  bc.edges_g = p.edges;
  const uint32_t publish_every = max(1u, p.flush_threshold >> 6);
  const uint32_t s_addr = smem_u32(S);
  unsigned long long evaluated = 0;  // per-thread share of the statistics: candidate pairs ...
  unsigned long long tested = 0;     // ... and pairs distance-tested (lane 0 of every warp)
  uint32_t unpublished = 0;          // candidate pairs since this warp last told the CTA
  const float4 nan4 = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000),
                                  __int_as_float(0x7fc00000), 0.f);

  for (;;) {
    __syncthreads();  // previous item finished; its successor is in the slots
    const int frame = (int)*reinterpret_cast<volatile uint32_t*>(s_misc + M_FRAME);
    const int pen = (int)*reinterpret_cast<volatile uint32_t*>(s_misc + M_PEN);
    const int pr = (int)*reinterpret_cast<volatile uint32_t*>(s_misc + M_PR);
    const int seg_i = (int)*reinterpret_cast<volatile uint32_t*>(s_misc + M_SEG);
    int4 ab = make_int4(0, 0, 1, 0);
    if (frame >= 0) ab = p.pair_ab[p.pair_lo + pr];  // row species, column species, segment length
    const bool same = ab.x == ab.y;
    const int cy = pen % ny, cz = pen / ny;
    // neighbour pencils (column species): the half shell for same-species pairs (dz = 0: dy = 1 ..
    // ky, then dz = 1 .. kz with every dy), all (2ky+1)(2kz+1) of them otherwise
    const int P = same ? kz * wy + ky : wy * (2 * kz + 1);
    if (tid < P) {
      int dy, dz;
      if (same) {
        if (tid < ky) { dy = tid + 1; dz = 0; }
        else { const int m = tid - ky; dz = 1 + m / wy; dy = m % wy - ky; }
      } else {
        dz = tid / wy - kz;
        dy = tid % wy - ky;
      }
      int yy = cy + dy, zz = cz + dz;
      yy += (yy < 0) ? ny : 0; yy -= (yy >= ny) ? ny : 0;
      zz += (zz < 0) ? p.nz : 0; zz -= (zz >= p.nz) ? p.nz : 0;
      s_pen[tid] = ab.y * nc + (zz * ny + yy) * nx;
    }
    __syncthreads();  // everyone has read the slots; the pencil table is written
    if (frame < 0) break;
    if (tid == 0) fetch_item();  // in flight during this item
    const uint32_t fl = p.frame_flags[frame];
    if (fl & MDS_FRAME_FAR) continue;  // the all-pairs kernel takes this frame
    const bool fast = (fl & 3u) != 3u;
    const uint32_t* __restrict__ off = p.cell_off + (long long)frame * p.off_stride;
    const float4* __restrict__ fpos = p.spos + (long long)frame * p.frame_stride;
    const uint32_t* __restrict__ Frow = off + ab.x * nc + (cz * ny + cy) * nx;  // own pencil, row species
    bc.c_hist = hist0 + (uint32_t)(pr * hist_stride) * 4u;
    bc.hist_g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;

    // this item: x cells [seg_i * ab.z, (seg_i + 1) * ab.z) of the pencil (nothing if the pair's
    // segments are longer than those of the pair that sets n_seg)
    int x0 = seg_i * ab.z, p_lo = 0;
    const int x_end = min(nx, x0 + ab.z);
    while (x0 < x_end) {
      // ---- stage Sx row cells against neighbour pencils [p_lo, p_hi): count, scan, and make it
      // smaller while it does not fit (cells are given up first, so a pencil subset, p_lo > 0,
      // only ever follows a one-cell pass)
      int Sx = p_lo > 0 ? 1 : x_end - x0, p_hi = P;
      const bool with_own = same && p_lo == 0;  // the pass that also takes the own pencil's +x half
      int Wa, Wr, Pn, n_r, n_a, n_ent;
      for (;;) {
        Pn = p_hi - p_lo;
        Wa = Sx + 2 * kx;
        Wr = with_own ? Sx + kx : Sx;
        while (Wa * Pn > MAX_ENT) {  // (uniform) keep the entry table in bounds
          if (Sx > 1) Sx >>= 1; else p_hi = p_lo + (Pn >> 1);
          Pn = p_hi - p_lo; Wa = Sx + 2 * kx; Wr = with_own ? Sx + kx : Sx;
        }
        n_ent = Wa * Pn;
        // phase 1: atoms per (pencil, x cell) entry — warps walk pencils, lanes x cells (neighbouring
        // lanes read neighbouring offsets); stored slab-major (x cell, pencil) for the scan
        for (int pi = warp; pi < Pn; pi += CELL_WARPS) {
          const uint32_t* F = off + s_pen[p_lo + pi];
          for (int c = lane; c < Wa; c += 32) {
            int xw = x0 - kx + c;
            xw += (xw < 0) ? nx : 0; xw -= (xw >= nx) ? nx : 0;
            const int src = (int)F[xw];
            E[c * Pn + pi] = (int)F[xw + 1] - src;
            Esrc[c * Pn + pi] = src;
          }
        }
        for (int c = tid; c < Wr; c += CELL_THREADS) {
          int xw = x0 + c;
          xw -= (xw >= nx) ? nx : 0;
          ro[c + 1] = (int)(Frow[xw + 1] - Frow[xw]);
        }
        if (tid == 0) { ro[0] = 0; s_misc[M_NEXT_CELL] = 0u; }
        __syncthreads();
        // phase 2: exclusive scan of the entry table (4 consecutive entries per thread)
        int v[EPT], tsum = 0;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          const int e = tid * EPT + j;
          v[j] = (e < n_ent) ? E[e] : 0;
          tsum += v[j];
        }
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        if (lane == 31) s_misc[M_WSUM + warp] = (uint32_t)inc;
        if (warp == 0) warp_scan_table(ro, Wr, lane);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CELL_WARPS; ++w) {
          const int ws = (int)s_misc[M_WSUM + w];
          before += (w < warp) ? ws : 0;
          total += ws;
        }
        int run = before + inc - tsum;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          const int e = tid * EPT + j;
          if (e < n_ent) E[e] = run;
          run += v[j];
        }
        n_a = total;
        n_r = ro[Wr];
        __syncthreads();  // offsets visible; also: everyone has read the warp sums and ro
        if (n_r + n_a <= p.cap_atoms && ro[Sx] <= ROWS_CAP - 2) break;
        if (Sx > 1) Sx >>= 1;
        else if (Pn > 1) p_hi = p_lo + (Pn >> 1);
        else __trap();  // the scan kernel hands such frames to the all-pairs kernel
      }
      // ---- phase 3: copy (asynchronous, 16 B per atom) ----------------------------------------
      for (int e = tid; e < n_ent; e += CELL_THREADS) {
        const int d0 = E[e], cnt = ((e + 1 < n_ent) ? E[e + 1] : n_a) - d0;
        float4* dst = S + n_r + d0;
        const float4* src = fpos + Esrc[e];
        for (int i = 0; i < cnt; ++i) cp_async16(dst + i, src + i);
      }
      for (int c = tid; c < Wr; c += CELL_THREADS) {
        int xw = x0 + c;
        xw -= (xw >= nx) ? nx : 0;
        const int src = (int)Frow[xw], cnt = ro[c + 1] - ro[c];
        float4* dst = S + ro[c];
        for (int i = 0; i < cnt; ++i) cp_async16(dst + i, fpos + src + i);
      }
      // ---- work units: chunks of 32 consecutive columns, first of the own pencil (same-species
      // pairs: columns after the row, up to kx cells ahead), then of the slab array.  Both arrays
      // are ordered by x cell, so the rows a chunk can pair with are a contiguous range: the cells
      // from (first cell of the chunk - stencil width) to its last cell.  Rows of that range that
      // are further than the stencil from a particular column are tested in vain (they can never be
      // inside the cutoff) — unless the pencil is so short that "further" wraps around to a
      // periodic image, in which case the chunk runs with per-column row ranges (masked).
      const int n_rows = ro[Sx];
      const int U_own = with_own ? (n_r + 31) >> 5 : 0;
      const int U = U_own + ((n_a + 31) >> 5);
      for (int u = tid; u < U; u += CELL_THREADS) {
        int rb, re, masked;
        if (u < U_own) {
          const int q0 = u << 5, q1 = min(q0 + 31, n_r - 1);
          const int s_lo = find_cell(ro, 1, Wr, q0), s_hi = find_cell(ro, 1, Wr, q1);
          rb = ro[max(0, s_lo - kx)] & ~1;      // whole row pairs
          re = min(ro[min(Sx, s_hi + 1)], q1);  // a row pairs with the columns after it
          masked = 1;
        } else {
          const int q0 = (u - U_own) << 5, q1 = min(q0 + 31, n_a - 1);
          const int s_lo = find_cell(E, Pn, Wa, q0), s_hi = find_cell(E, Pn, Wa, q1);
          rb = ro[max(0, s_lo - 2 * kx)] & ~1;
          re = ro[min(Sx, s_hi + 1)];
          // Row cell x' and slab c' (x cell x' + c' - kx - x') are partners when 0 <= c' - x' <= 2 kx;
          // any other combination in the range is tested in vain — harmless unless the offset
          // c' - kx - x' reaches a periodic image of a partner cell, |c' - kx - x'| >= nx - kx.
          // The extreme rows (incl. the ones that complete a pair) against the extreme slabs:
          masked = 0;
          if (re > rb) {
            const int last = min(((re - rb + 1) & ~1) + rb, n_rows) - 1;  // last real row touched
            const int xr_lo = find_cell(ro, 1, Sx, rb), xr_hi = find_cell(ro, 1, Sx, last);
            masked = (s_hi - kx - xr_lo >= nx - kx || xr_hi + kx - s_lo >= nx - kx) ? 1 : 0;
          }
        }
        s_unit[u] = make_int2(rb, (re > rb ? (re - rb + 1) >> 1 : 0) | (masked << 30));
      }
      // candidate pairs of this segment (statistics): pairs whose cells are within the stencil
      for (int c = tid; c < Sx; c += CELL_THREADS) {
        const int n = ro[c + 1] - ro[c];
        const int a0 = E[c * Pn], a1 = (c + 2 * kx + 1 < Wa) ? E[(c + 2 * kx + 1) * Pn] : n_a;
        long long ev = (long long)n * (a1 - a0);
        if (with_own) ev += (long long)n * (n - 1) / 2 + (long long)n * (ro[c + kx + 1] - ro[c + 1]);
        evaluated += (unsigned long long)ev;
      }
      cp_async_wait_all();
      __syncthreads();
      // the row atoms once more as structure-of-arrays (+ a NaN pad so that pairs are whole)
      for (int i = tid; i <= (n_rows | 1); i += CELL_THREADS) {
        float4 v = nan4;
        if (i < n_rows) v = S[i];
        RX[i] = v.x; RY[i] = v.y; RZ[i] = v.z;
      }
      __syncthreads();

      // ---- compute: one chunk per warp at a time ---------------------------------------------
      for (;;) {
        int u = 0;
        if (lane == 0) u = (int)atomicAdd(s_misc + M_NEXT_CELL, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= U) break;
        const int2 un = s_unit[u];
        const int rb = un.x, n2 = un.y & 0x3fffffff;
        if (n2 == 0) continue;
        const bool masked = (un.y >> 30) != 0;
        const bool own = u < U_own;
        const int q = ((own ? u : u - U_own) << 5) + lane;
        const int lim = own ? n_r : n_a;
        float4 c = nan4;  // slots past the end hold NaN atoms (never inside the cutoff)
        if (q < lim) c = lds128(s_addr + ((uint32_t)((own ? 0 : n_r) + q) << 4));
        if (!masked && fast) {
          rows_vs_column<true, false, HIST_SMEM, INTER>(RX + rb, RY + rb, RZ + rb, n2, c, 0, 0u, bc, issued);
        } else {
          // rows [lo, lo + len) of this chunk's row range are the ones column q may pair with
          int lo = 0;
          uint32_t len = 0x7fffffffu;
          if (masked) {
            if (own) {
              const int cell = find_cell(ro, 1, Wr, min(q, n_r - 1));
              lo = ro[max(0, cell - kx)] - rb;
              len = (uint32_t)max(0, q - rb - lo);
            } else {
              const int cell = find_cell(E, Pn, Wa, min(q, n_a - 1));
              lo = ro[max(0, cell - 2 * kx)] - rb;
              len = (uint32_t)max(0, ro[min(Sx, cell + 1)] - rb - lo);
            }
          }
          if (fast) rows_vs_column<true, true, HIST_SMEM, INTER>(RX + rb, RY + rb, RZ + rb, n2, c, lo, len, bc,
                                                         issued);
          else rows_vs_column<false, true, HIST_SMEM, INTER>(RX + rb, RY + rb, RZ + rb, n2, c, lo, len, bc,
                                                      issued);
        }
        tested += (uint32_t)n2 * 64u;
        if (HIST_SMEM) {
          unpublished += (uint32_t)n2 * 64u;
          if (unpublished >= publish_every) {
            uint32_t old = 0;
            if (lane == 0) old = atomicAdd(s_misc + M_EVALS, unpublished);
            old = __shfl_sync(0xffffffffu, old, 0);
            // the warp whose contribution crosses the threshold empties the histogram
            if (old < p.flush_threshold && old + unpublished >= p.flush_threshold) {
              if (lane == 0) atomicSub(s_misc + M_EVALS, p.flush_threshold);
              flush_hist_warp(s_hist, hs, p, lane MDS_CHK(, s_chk));
            }
            unpublished = 0;
          }
        }
      }
      __syncthreads();  // the staged segment is consumed
      if (p_hi < P) p_lo = p_hi;
      else { p_lo = 0; x0 += Sx; }
    }
  }

  evaluated = __reduce_add_sync(0xffffffffu, (unsigned)(evaluated & 0xffffffffull)) +
              ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)(evaluated >> 32)) << 32);
  if (lane == 0 && evaluated) atomicAdd(p.evaluated, evaluated);
  if (lane == 0 && tested) atomicAdd(p.evaluated + 1, tested);
  if (HIST_SMEM) {
    __syncthreads();
    MDS_CHK(checked_conservation(s_chk, issued, s_hist, p.pair_cnt * hist_stride, tid, CELL_THREADS,
                                 "rdf_cells_kernel", hs);)
    const int stride = p.nbins + 1;
    for (int pr = 0; pr < p.pair_cnt; ++pr) {
      unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
      const uint32_t* h = s_hist + pr * stride;
      for (int k = tid; k < p.nbins; k += CELL_THREADS) {
        const uint32_t v = h[k * hs];
        if (v) atomicAdd(g + k, (unsigned long long)v);
      }
    }
  }
}

// ---- host-side launchers ------------------------------------------------------------------

int cells_max_seg() { return MAXW - 2 * 4; }
int cells_max_cap() { return MAX_CAP; }
// words of launch_cell_build's scratch: a fixed region for the per-frame maxima (MAX_SCAN_FRAMES
// frames), then the block totals
constexpr int MAX_SCAN_FRAMES = 4096;
size_t cells_scan_tmp_words(int n_frames, int n_keys) {
  return (size_t)MAX_SCAN_FRAMES + (size_t)n_frames * (size_t)((n_keys + 4095) / 4096);
}
int cells_max_entries() { return MAX_ENT; }

size_t cells_smem_bytes(int nbins, int pair_cnt, bool hist_in_smem, int cap_atoms) {
  size_t b = (size_t)CELLS_FIXED_BYTES + 3 * (size_t)ROWS_CAP * sizeof(float) +
             (size_t)cap_atoms * sizeof(float4);
  if (hist_in_smem) {
    b += (size_t)((nbins + 2 + 3) & ~3) * sizeof(float);
    b += (size_t)pair_cnt * (size_t)(nbins + 1) * sizeof(uint32_t);
  }
  return b;
}

bool cells_build_fused(int layout, int n_keys) {
  return layout == MDS_LAYOUT_FRAME_MAJOR_XYZ && n_keys <= BUILD_MAX_KEYS;
}

// frames the one-kernel build works on at a time: its resident CTAs on the device (0: not eligible)
int cells_build_frames_per_wave(int layout, int n_keys, int sm_count) {
  if (!cells_build_fused(layout, n_keys)) return 0;
  const size_t smem = ((size_t)n_keys + 1) * sizeof(uint32_t);
  auto kern = rdf_cell_build_kernel<false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BUILD_THREADS, smem) != cudaSuccess)
    return 0;
  return per_sm * sm_count;
}

cudaError_t launch_cell_build(const PackParams& pp, int n_species, float4* sorted,
                              uint32_t crowded_limit, uint32_t* scan_tmp, int sm_count,
                              bool allow_fused, bool write_packed, cudaStream_t stream) {
  const long long total = (long long)pp.n_frames * pp.n_packed;
  if (total <= 0) return cudaSuccess;
  const int n_keys = n_species * pp.grid.nc;
  if (allow_fused && cells_build_fused(pp.layout, n_keys)) {
    // one CTA per frame, persistent: counts and running offsets live in shared memory
    const size_t smem = ((size_t)n_keys + 1) * sizeof(uint32_t);
    auto kern = write_packed ? rdf_cell_build_kernel<true> : rdf_cell_build_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BUILD_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)per_sm * sm_count;
    if (grid > pp.n_frames) grid = pp.n_frames;
    kern<<<(unsigned)grid, BUILD_THREADS, smem, stream>>>(pp, sorted, n_keys, crowded_limit);
    return cudaGetLastError();
  }
  cudaError_t e = cudaMemsetAsync(pp.cell_off, 0,
                                  sizeof(uint32_t) * (size_t)pp.off_stride * (size_t)pp.n_frames,
                                  stream);
  if (e != cudaSuccess) return e;
  e = launch_pack(pp, sm_count, stream);  // quad-block positions + per-cell counts and ranks + frame flags
  if (e != cudaSuccess) return e;
  const int n_blocks = (n_keys + 4095) / 4096;
  if (n_blocks <= 2 || !scan_tmp) {
    rdf_cell_scan_kernel<<<(unsigned)pp.n_frames, 1024, 0, stream>>>(
        pp.cell_off, pp.off_stride, n_keys, pp.frame_flags, crowded_limit);
  } else {
    // scan_tmp: [frames] per-frame maxima (kept zero between batches), then [frames][n_blocks] totals
    uint32_t* frame_max = scan_tmp;
    uint32_t* block_sums = scan_tmp + cells_scan_tmp_words(1, 0);
    const dim3 grid((unsigned)n_blocks, (unsigned)pp.n_frames);
    rdf_cell_scan_blocks_kernel<<<grid, 1024, 0, stream>>>(pp.cell_off, pp.off_stride, n_keys,
                                                           block_sums, frame_max);
    rdf_cell_scan_finish_kernel<<<grid, 1024, 0, stream>>>(pp.cell_off, pp.off_stride, n_keys,
                                                           block_sums, frame_max, pp.frame_flags,
                                                           crowded_limit);
  }
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sm_count * 64) blocks = (long long)sm_count * 64;
  rdf_cell_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pp, sorted);
  return cudaGetLastError();
}

cudaError_t launch_cells(const CellsParams& p, int sm_count, cudaStream_t stream, int* grid_out) {
  const size_t smem = cells_smem_bytes(p.nbins, p.pair_cnt, p.hist_in_smem != 0, p.cap_atoms);
  // one species pair per launch: threshold table and histogram interleaved (one address per pair)
  auto kern = !p.hist_in_smem ? rdf_cells_kernel<false, false>
              : (p.pair_cnt == 1 && !getenv("MDS_RDF_CELL_TWO_TABLES")) ? rdf_cells_kernel<true, true>
                                                                        : rdf_cells_kernel<true, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CELL_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  unsigned long long items = (unsigned long long)p.n_frames * (unsigned long long)p.pair_cnt *
                             (unsigned long long)p.n_pen * (unsigned long long)p.n_seg;
  items = (items + p.shard_count - 1) / p.shard_count;
  if (items == 0) return cudaSuccess;
  unsigned long long grid = (unsigned long long)per_sm * sm_count;  // persistent
  if (grid > items) grid = items;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  kern<<<(unsigned)grid, CELL_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mds
