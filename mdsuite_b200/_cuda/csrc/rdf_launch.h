// Host-visible launch entry points of the kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mds {

struct AllPairsParams;

constexpr int MDS_MAX_BINS = 65536;  // bin guess stays within +-1 of the exact bin below this

// ---- rdf_allpairs.cu ----
size_t allpairs_smem_bytes(int nbins, int pair_cnt, bool hist_in_smem);
cudaError_t launch_allpairs(const AllPairsParams& p, int variant, int sm_count,
                            cudaStream_t stream, int* grid_out);
int allpairs_variant_tile(int variant);  // rows per work item of a variant, -1 if unknown
constexpr int ALLPAIRS_N_VARIANTS = 7;

// ---- rdf_pack.cu ----
struct CellGrid {
  int nx, ny, nz, nc;  // cells per dimension, nc = nx * ny * nz (each >= 3)
  float box[3];
  // item shards of the cell path (mds_rdf_set_item_shard): this handle only works on a slab of z
  // layers, so only atoms in layers [z_lo, z_lo + z_len) (periodic; the slab plus its stencil)
  // are counted, ranked and sorted.  z_len >= nz: every atom.
  int z_lo, z_len;
};
__host__ __device__ inline bool cell_in_slab(int cell, const CellGrid& g) {
  if (g.z_len >= g.nz) return true;
  int dz = cell / (g.nx * g.ny) - g.z_lo;
  dz += (dz < 0) ? g.nz : 0;
  return dz < g.z_len;
}
struct PackParams {
  const float* xyz;        // raw input, layout below
  int layout;              // MDS_LAYOUT_*
  int64_t n_atoms;         // atoms per frame in the raw input
  int64_t n_frames;        // frames in this batch
  int64_t frame0;          // first frame of the batch inside the raw buffer (atom-major only)
  int64_t n_frames_total;  // frames in the raw buffer (stride of the atom-major layout)
  const int32_t* src_index;  // [n_packed] raw atom index of each packed slot, -1: padding
  const int32_t* species;    // [n_packed] species id of each packed slot
  int32_t n_packed;        // packed slots per frame (species segments padded to multiples of 4)
  float* out;              // [n_frames][frame_stride / 4][3][4] quad blocks
  int64_t frame_stride;    // atom slots between frames
  // cell-list path only (cell_off == nullptr otherwise)
  uint32_t* cell_off;      // [n_frames][off_stride] zeroed counts of key = species * nc + cell
  int64_t off_stride;
  uint32_t* rank;          // [n_frames][frame_stride] rank of the atom inside its (species, cell)
  CellGrid grid;
  uint32_t* frame_flags;   // [n_frames], zeroed by the kernel's caller
  float box[3];
  // pack only what a fallback needs: nothing if *gate == 0, else the frames whose flags have one
  // of need_bits (nullptr / 0: every frame)
  const unsigned long long* gate;
  uint32_t need_bits;
};
cudaError_t launch_pack(const PackParams& p, int sm_count, cudaStream_t stream);
// out[0] += frames on the general minimum-image path, out[1] += frames flagged MDS_FRAME_FAR,
// *batch_far (nullable) = frames flagged MDS_FRAME_FAR in this batch
cudaError_t launch_count_general(const uint32_t* flags, int n, unsigned long long* out,
                                 unsigned long long* batch_far, cudaStream_t stream);

// frame_flags bits written by the pack kernels
constexpr uint32_t MDS_FRAME_OUT_A = 1u;  // a coordinate outside [-L/8, 9L/8]
constexpr uint32_t MDS_FRAME_OUT_B = 2u;  // a coordinate outside [-5L/8, 5L/8]
constexpr uint32_t MDS_FRAME_FAR = 4u;    // a coordinate beyond 8 box lengths: cell path declines

// ---- rdf_cells.cu ----
struct CellsParams {
  const float4* spos;         // [n_frames][frame_stride] atoms sorted by (species, cell)
  int64_t frame_stride;
  int32_t n_frames;
  const uint32_t* cell_off;   // [n_frames][off_stride]: offsets of key = species * nc + cell
  int64_t off_stride;         // >= n_species * nc + 1
  const uint32_t* frame_flags;
  int32_t nx, ny, nz, nc;
  int32_t kx, ky, kz;         // stencil half-widths: partners lie at most k cells away (1 or 2)
  int32_t cap_atoms;          // staged positions per CTA (shared memory), see rdf_cells.cu
  const int4* pair_ab;        // [n_pairs_total] (row species, column species, x cells per segment, 0)
  const float* edges;
  unsigned long long* counts;     // [n_pairs_total][nbins]
  unsigned long long* evaluated;  // [0] += candidate pairs (cells within the stencil), [1] += pairs distance-tested
  unsigned long long* work_counter;  // zeroed before launch
  int32_t nbins, pair_lo, pair_cnt, hist_in_smem;
  float s_cut, inv_step;
  float box[3];
  uint32_t flush_threshold;   // candidate pairs per CTA between histogram flushes (<= 2^30)
  // item shards: with at least as many z layers as shards the handle takes the pencils [pen_lo,
  // pen_lo + n_pen) of every frame (a slab of z layers: only that slab and its stencil have to be
  // sorted); otherwise every shard_count-th (frame, pair, pencil, segment) item from shard_index
  int32_t pen_lo, n_pen;
  uint32_t shard_index, shard_count;
  int32_t n_seg;              // x segments per pencil = work items per (frame, pair, pencil): the
                              // largest ceil(nx / segment length) over the species pairs
};
size_t cells_smem_bytes(int nbins, int pair_cnt, bool hist_in_smem, int cap_atoms);
int cells_max_seg();   // longest segment (x cells) the kernel's tables can hold
int cells_max_cap();   // largest cap_atoms the kernel's tables can hold
int cells_max_entries();  // (x cell, neighbour pencil) entries of a staged segment, at most
// pp must have cell_off / rank / grid set: cell offsets into pp.cell_off, atoms ordered by
// (species, cell) into `sorted`; frames with a cell of more than crowded_limit atoms are flagged
// MDS_FRAME_FAR
// scan_tmp: cells_scan_tmp_words(frames per batch, keys) zeroed uint32 words (nullptr: one CTA
// scans a frame's keys on its own)
size_t cells_scan_tmp_words(int n_frames, int n_keys);
// allow_fused: frame-major input whose key table fits in shared memory (cells_build_fused) is
// built by ONE kernel, a CTA per frame; otherwise (and when false) pack / scan / scatter.
// write_packed = false (one-kernel build only): pp.out is left alone — the caller packs the frames
// that need the all-pairs kernel afterwards (launch_pack with gate / need_bits)
bool cells_build_fused(int layout, int n_keys);
int cells_build_frames_per_wave(int layout, int n_keys, int sm_count);
cudaError_t launch_cell_build(const PackParams& pp, int n_species, float4* sorted,
                              uint32_t crowded_limit, uint32_t* scan_tmp, int sm_count,
                              bool allow_fused, bool write_packed, cudaStream_t stream);
cudaError_t launch_cells(const CellsParams& p, int sm_count, cudaStream_t stream, int* grid_out);

// ---- microbench.cu ----
cudaError_t run_shared_atomic_bench(int nbins, int iters, int mode, int sm_count, float* ms,
                                    double* lanes);
}  // namespace mds
