// All-pairs RDF kernel for sm_100a: persistent CTAs, TMA-staged column tiles, block-private
// shared-memory histograms, exact float32 arithmetic.
//
// Replaces, fused into one kernel, the reference's K1-K9 TensorFlow ops:
//   get_partial_triu_indices   mdsuite/utils/linalg.py:102-122      -> the (row, col) tile schedule
//   get_dij                    calculators/radial_distribution_function.py:647-689
//   apply_minimum_image        mdsuite/utils/linalg.py:84-99
//   bin_minibatch              calculators/radial_distribution_function.py:616-645
//   apply_system_cutoff        mdsuite/utils/linalg.py:125-136
//   tf.histogram_fixed_width   calculators/radial_distribution_function.py:642-644
//   combine_dictionaries       calculators/radial_distribution_function.py:601-614, :884-885
//
// Layout: one work item = (frame, species pair, row tile of THREADS*R atoms, column range).
// Each thread keeps R row atoms in registers and walks the column range through a
// double-buffered shared-memory tile (quad blocks: x[4] y[4] z[4]) that one elected thread fills
// with cp.async.bulk (TMA 1-D) and an mbarrier; three broadcast LDS.128 bring four column atoms,
// and each (row, column pair) is evaluated with the packed FADD2 / FMUL2 ops of sm_100a.
// In-cutoff pairs are binned against the squared-distance threshold table and counted with a
// shared-memory atomic; the CTA's histogram is merged into the global uint64 counts once, at
// the end (or before 2^31 increments could overflow a 32-bit bin).
#include "rdf_common.cuh"
#include "rdf_launch.h"

namespace mds {

#ifndef MDS_COND
#define MDS_COND false
#endif
constexpr int TJ = 512;                          // column atoms per shared-memory tile
constexpr int TJ_F4 = TJ / 4 * 3;                // float4 rows per tile stage (quad blocks), 6 KB

// One tile of columns (quad blocks in shared memory) against the R register rows of a thread.
// Every block of four columns is three broadcast LDS.128; each (row, column pair) goes through
// the packed FADD2 / FMUL2 sequence of pair2_s.
// RA <= R is the number of row registers that can hold a counted pair for these columns: row
// tiles at the end of a species fill fewer than R registers, and below the diagonal whole
// registers are masked — they are skipped instead of being evaluated into the sink.
template <int R, int RA, bool FAST, bool DIAG, bool HIST_SMEM, bool DENSE>
__device__ __forceinline__ void pair_tile(const float4* __restrict__ jt, int cnt, int jbase,
                                          const float (&xi)[R], const float (&yi)[R],
                                          const float (&zi)[R], const int (&ig)[R], float bx,
                                          float by, float bz, float s_cut, float inv_step,
                                          float inv_hi, uint32_t c_edge, uint32_t c_hist,
                                          uint32_t four,
                                          const float* __restrict__ edges_g,
                                          unsigned long long* hist_g,
                                          unsigned long long& issued) {
  (void)issued;
  float hi = 0.f;  // last table entry a lane looked up (see bin_count_dense)
#pragma unroll 1
  for (int jj = 0; jj < cnt; jj += 4, jt += 3) {
    const float4 X = jt[0], Y = jt[1], Z = jt[2];
    const int jg = jbase + jj;
#pragma unroll
    for (int r = 0; r < RA; ++r) {
      float s[4];
      pair2_s<FAST>(X.x, X.y, Y.x, Y.y, Z.x, Z.y, xi[r], yi[r], zi[r], bx, by, bz, s[0], s[1]);
      pair2_s<FAST>(X.z, X.w, Y.z, Y.w, Z.z, Z.w, xi[r], yi[r], zi[r], bx, by, bz, s[2], s[3]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (HIST_SMEM && DENSE) {
          bin_count_dense<DIAG, MDS_COND>(s[c], s_cut, inv_step, inv_hi, c_edge, c_hist, four, jg + c,
                                      ig[r], hi);
          MDS_CHK(issued += 1;)
        } else {
          bool in = s[c] < s_cut;  // false for NaN (padding rows / columns, NaN input)
          if (DIAG) in = in && (jg + c > ig[r]);
          if (in) {
            if (HIST_SMEM) {
              bin_count(s[c], inv_step, c_edge, c_hist, four);
              MDS_CHK(issued += 1;)
            } else {
              atomicAdd(hist_g + bin_of_s(s[c], inv_step, edges_g), 1ull);
            }
          }
        }
      }
    }
  }
}

// Coalesced merge of the block-private histogram into the global uint64 counts (skips the sink
// slot of every pair).
template <int THREADS>
__device__ __forceinline__ void merge_hist(uint32_t* s_hist, const AllPairsParams& p, int tid,
                                           bool clear, unsigned long long* s_chk = nullptr) {
  (void)s_chk;
  const int stride = p.nbins + 1;
  for (int pr = 0; pr < p.pair_cnt; ++pr) {
    unsigned long long* g = p.counts + (long long)(p.pair_lo + pr) * p.nbins;
    uint32_t* h = s_hist + pr * stride;
    for (int k = tid; k < p.nbins; k += THREADS) {
      const uint32_t v = h[k];
      if (v) {
        atomicAdd(g + k, (unsigned long long)v);
        if (clear) {
          h[k] = 0u;
          MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)v);)
        }
      }
    }
    if (clear && tid == 0) {
      MDS_CHK(atomicAdd(&s_chk[1], (unsigned long long)h[p.nbins]);)
      h[p.nbins] = 0u;
    }
  }
}

template <int THREADS, int R, int MINB, bool HIST_SMEM>
__global__ void __launch_bounds__(THREADS, MINB) rdf_allpairs_kernel(const AllPairsParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* jt = reinterpret_cast<float4*>(smem_raw);                  // [2][TJ_F4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(jt + 2 * TJ_F4);      // [2]
  volatile int* s_item = reinterpret_cast<volatile int*>(bars + 2);  // 32 B: item + constants
  float* s_edges = reinterpret_cast<float*>(bars + 6);
  const int n_edges = p.nbins + 2;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_edges + ((n_edges + 3) & ~3));
  const int hist_stride = p.nbins + 1;  // slot nbins of every pair is the out-of-cutoff sink
  const int n_hist = p.pair_cnt * hist_stride;
  const int tid = threadIdx.x;
  MDS_CHK(__shared__ unsigned long long s_chk[2]; unsigned long long issued = 0;
          if (tid == 0) { s_chk[0] = 0; s_chk[1] = 0; })
#ifndef MDS_CHECKED
  unsigned long long issued = 0;  // only counted in checked builds
#endif

  if (HIST_SMEM) {
    for (int k = tid; k < n_edges; k += THREADS) s_edges[k] = p.edges[k];
    for (int k = tid; k < n_hist; k += THREADS) s_hist[k] = 0u;
  }
  if (p.gate && *p.gate == 0ull) return;  // nothing was handed over by the cell path
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    // The two address constants of the binning sequence go through shared memory so that ptxas
    // sees opaque registers: folded into the address arithmetic as known constants they cost two
    // extra integer instructions per pair (LEA + IADD3 with the split magic constant).
    s_item[1] = (int)(smem_u32(s_edges + 1) - kMagic4);
    s_item[2] = (int)(smem_u32(s_hist) - kMagic4);
    s_item[4] = 4;
    s_item[5] = 4;
  }
  __syncthreads();

  const uint32_t c_edge = (uint32_t)s_item[1];
  const uint32_t hist0 = (uint32_t)s_item[2];
  // read through a lane-dependent address so that ptxas keeps it in a vector register (as a
  // uniform register it would need a MOV in front of every IMAD that also takes c_edge / c_hist)
  const uint32_t four = (uint32_t)s_item[4 + (tid & 1)];
  const float bx = p.box[0], by = p.box[1], bz = p.box[2];
  const float s_cut = p.s_cut, inv_step = p.inv_step, inv_hi = p.inv_step_hi;
  const bool dense = HIST_SMEM && ((p.dense & 1) != 0);
  const long long total = (long long)p.n_items * (long long)p.n_frames;
  uint32_t ph0 = 0, ph1 = 0;
  unsigned long long since_flush = 0;
  unsigned long long my_pairs = 0;  // thread 0: reference pairs of the units this CTA processed

  for (;;) {
    if (tid == 0) *s_item = (int)atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const long long idx = (long long)(unsigned int)*s_item * p.shard_count + p.shard_index;
    __syncthreads();  // everyone has read the slot before thread 0 overwrites it
    if (idx >= total) break;
    // big items first: consecutive indices walk the frames of one item
    const PairItem it = p.items[idx / p.n_frames];
    const int frame = (int)(idx % p.n_frames);
    if (p.frame_need_bits && !(p.frame_flags[frame] & p.frame_need_bits)) continue;  // CTA-uniform
    if (tid == 0) my_pairs += ((unsigned long long)it.pairs_hi << 32) | it.pairs_lo;
    const float* __restrict__ fpos = p.pos + (long long)frame * p.frame_stride * 3;
    const bool fast = ((p.frame_flags[frame] & 3u) != 3u);  // see rdf_pack.cu

    float xi[R], yi[R], zi[R];
    int ig[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int li = tid + r * THREADS;
      ig[r] = it.i0 + li;
      if (li < it.i_cnt) {
        const float* q = fpos + (ig[r] >> 2) * QUAD_FLOATS + (ig[r] & 3);
        xi[r] = __ldg(q); yi[r] = __ldg(q + 4); zi[r] = __ldg(q + 8);
      } else {
        xi[r] = yi[r] = zi[r] = __int_as_float(0x7fc00000);  // NaN: never in cutoff
      }
    }

    uint32_t c_hist = hist0 + (uint32_t)((it.pair - p.pair_lo) * hist_stride) * 4u;
    // fault injection for the self-check's own test: aim the atomics at the column tiles
    MDS_CHK(if (p.dense & 2) c_hist = smem_u32(jt) - kMagic4;)
    const int r_item = min(R, (it.i_cnt + THREADS - 1) / THREADS);  // row registers in use
    unsigned long long* hist_g = p.counts + (long long)it.pair * p.nbins;
    const int n_cols = it.j1 - it.j0;
    const int n_tiles = (n_cols + TJ - 1) / TJ;

    if (tid == 0) {
      const uint32_t bytes = (uint32_t)min(TJ, n_cols) * 12u;  // j0, j1 are multiples of 4
      fence_proxy_async();
      mbar_expect_tx(&bars[0], bytes);
      tma_load_1d(jt, fpos + (long long)it.j0 * 3, bytes, &bars[0]);
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t & 1;
      if (tid == 0 && t + 1 < n_tiles) {
        const int nb = it.j0 + (t + 1) * TJ;
        const uint32_t bytes = (uint32_t)min(TJ, it.j1 - nb) * 12u;
        fence_proxy_async();
        mbar_expect_tx(&bars[st ^ 1], bytes);
        tma_load_1d(jt + (st ^ 1) * TJ_F4, fpos + (long long)nb * 3, bytes, &bars[st ^ 1]);
      }
      if (st == 0) { mbar_wait(&bars[0], ph0); ph0 ^= 1u; }
      else         { mbar_wait(&bars[1], ph1); ph1 ^= 1u; }
      const int jbase = it.j0 + t * TJ;
      const int cnt = min(TJ, it.j1 - jbase);
      const float4* tile = jt + st * TJ_F4;
      // Segments of THREADS columns: row register r holds rows [i0 + r THREADS, i0 + (r+1) THREADS),
      // so on a same-species item the columns of segment g = (col - i0) / THREADS can only pair
      // (col > row) with registers r <= g, and need the mask only while g < r_item.
#define MDS_TILE(RAV, FASTV, DIAGV, DENSEV)                                                    \
  pair_tile<R, RAV, FASTV, DIAGV, HIST_SMEM, DENSEV>(seg, scnt, sbase, xi, yi, zi, ig, bx, by, bz, \
                                                     s_cut, inv_step, inv_hi, c_edge, c_hist,   \
                                                     four, p.edges, hist_g, issued)
#define MDS_TILE_FD(RAV)                                                                       \
  do {                                                                                         \
    if (dense) { if (dg) MDS_TILE(RAV, true, true, true); else MDS_TILE(RAV, true, false, true); } \
    else { if (dg) MDS_TILE(RAV, true, true, false); else MDS_TILE(RAV, true, false, false); }  \
  } while (0)
      const int seg_cols = (it.diag && jbase < it.i0 + r_item * THREADS) ? THREADS : TJ;
      for (int c0 = 0; c0 < cnt; c0 += seg_cols) {
        const float4* seg = tile + (c0 >> 2) * 3;
        const int scnt = min(seg_cols, cnt - c0);
        const int sbase = jbase + c0;
        int ra = r_item;
        bool dg = false;
        if (it.diag) {
          const int g = (sbase - it.i0) / THREADS;
          ra = min(r_item, g + 1);
          dg = g < r_item;
        }
        if (!fast) {  // literal minimum image (rare): always the full-R bodies
          if (dense) { if (dg) MDS_TILE(R, false, true, true); else MDS_TILE(R, false, false, true); }
          else { if (dg) MDS_TILE(R, false, true, false); else MDS_TILE(R, false, false, false); }
        } else if (R >= 4 && ra >= 4) {
          MDS_TILE_FD(R >= 4 ? 4 : R);
        } else if (R >= 3 && ra == 3) {
          MDS_TILE_FD(R >= 3 ? 3 : R);
        } else if (R >= 2 && ra == 2) {
          MDS_TILE_FD(R >= 2 ? 2 : R);
        } else {
          MDS_TILE_FD(1);
        }
      }
#undef MDS_TILE_FD
#undef MDS_TILE
      __syncthreads();  // tile `st` fully consumed before it is refilled two iterations later
    }

    if (HIST_SMEM) {
      since_flush += (unsigned long long)it.i_cnt * (unsigned long long)n_cols;
      if (since_flush >= (1ull << 31)) {  // uniform across the CTA
        merge_hist<THREADS>(s_hist, p, tid, true MDS_CHK(, s_chk));
        since_flush = 0;
        __syncthreads();
      }
    }
  }

  if (tid == 0 && my_pairs && p.evaluated) atomicAdd(p.evaluated, my_pairs);
  if (HIST_SMEM) {
    MDS_CHK(__syncthreads();
            checked_conservation(s_chk, issued, s_hist, n_hist, tid, THREADS, "rdf_allpairs_kernel");)
    merge_hist<THREADS>(s_hist, p, tid, false);
  }
}

// ---- host-side launch table ----------------------------------------------------------------

size_t allpairs_smem_bytes(int nbins, int pair_cnt, bool hist_in_smem) {
  size_t b = 2 * TJ_F4 * sizeof(float4) + 6 * sizeof(uint64_t);
  if (hist_in_smem) {
    b += (size_t)((nbins + 2 + 3) & ~3) * sizeof(float);
    b += (size_t)pair_cnt * (size_t)(nbins + 1) * sizeof(uint32_t);
  }
  return b;
}

template <int THREADS, int R, int MINB>
static cudaError_t launch_variant(const AllPairsParams& p, int sm_count, cudaStream_t stream,
                                  int* grid_out) {
  const size_t smem = allpairs_smem_bytes(p.nbins, p.pair_cnt, p.hist_in_smem != 0);
  auto kern = p.hist_in_smem ? rdf_allpairs_kernel<THREADS, R, MINB, true>
                             : rdf_allpairs_kernel<THREADS, R, MINB, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  long long total = (long long)p.n_items * p.n_frames;
  total = (total - p.shard_index + p.shard_count - 1) / p.shard_count;  // this shard's units
  long long grid = (long long)per_sm * sm_count;  // persistent: a multiple of the SM count
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  kern<<<(unsigned)grid, THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_allpairs(const AllPairsParams& p, int variant, int sm_count,
                            cudaStream_t stream, int* grid_out) {
  switch (variant) {
    case 0: return launch_variant<256, 1, 3>(p, sm_count, stream, grid_out);
    case 1: return launch_variant<256, 2, 3>(p, sm_count, stream, grid_out);
    case 2: return launch_variant<256, 4, 3>(p, sm_count, stream, grid_out);
    case 3: return launch_variant<512, 2, 2>(p, sm_count, stream, grid_out);
    case 4: return launch_variant<512, 4, 2>(p, sm_count, stream, grid_out);
    case 5: return launch_variant<256, 4, 2>(p, sm_count, stream, grid_out);
    case 6: return launch_variant<256, 4, 4>(p, sm_count, stream, grid_out);
    default: return cudaErrorInvalidValue;
  }
}

int allpairs_variant_tile(int variant) {
  static const int tiles[] = {256, 512, 1024, 1024, 2048, 1024, 1024};
  return (variant >= 0 && variant < 7) ? tiles[variant] : -1;
}

}  // namespace mds
