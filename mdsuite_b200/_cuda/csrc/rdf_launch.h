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
constexpr int ALLPAIRS_N_VARIANTS = 6;

// ---- rdf_pack.cu ----
struct PackParams {
  const float* xyz;        // raw input, layout below
  int layout;              // MDS_LAYOUT_*
  int64_t n_atoms;         // atoms per frame in the raw input
  int64_t n_frames;        // frames in this batch
  int64_t frame0;          // first frame of the batch inside the raw buffer (atom-major only)
  int64_t n_frames_total;  // frames in the raw buffer (stride of the atom-major layout)
  const int32_t* src_index;  // [n_packed] raw atom index of each packed atom
  const int32_t* species;    // [n_packed] species id of each packed atom
  int32_t n_packed;
  float4* out;             // [n_frames][frame_stride]
  int64_t frame_stride;
  uint32_t* frame_flags;   // [n_frames], zeroed by the kernel's caller
  float box[3];
};
cudaError_t launch_pack(const PackParams& p, cudaStream_t stream);
cudaError_t launch_count_general(const uint32_t* flags, int n, unsigned long long* out,
                                 cudaStream_t stream);

// ---- microbench.cu ----
cudaError_t run_shared_atomic_bench(int nbins, int iters, int mode, int sm_count, float* ms,
                                    double* lanes);
}  // namespace mds
