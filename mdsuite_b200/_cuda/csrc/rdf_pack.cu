// Pack kernel: raw float32 xyz frames (either layout of include/mds_rdf.h) -> the internal
// per-frame quad-block array the pair kernels read (rdf_common.cuh: x[4] y[4] z[4] per four
// atoms, 12 B per atom; species segments padded to multiples of four with NaN atoms).
//
// Follows the data preparation of the reference's batch loop:
//   _format_data  calculators/radial_distribution_function.py:535-563  (species concatenated)
//   quirk Q1      calculators/radial_distribution_function.py:637-638  (strict '>' drops the atom at
//                 each species' offset) — in parity mode those atoms are simply not packed
//   Database.load_data database/simulation_database.py:594-639 produces the atom-major
//                 (N, C, 3) layout that MDS_LAYOUT_ATOM_MAJOR_XYZ accepts as is.
// It also classifies every frame: if all coordinates lie in one window of length 1.25 box
// the 2-op minimum image is bit-identical to the reference's div/rint form;
// otherwise the pair kernel uses the literal form for that frame.
#include "rdf_common.cuh"
#include "rdf_launch.h"
#include "mds_rdf.h"

namespace mds {

// Thread layout without integer division: blockIdx.y walks the OUTER index, the x dimension the
// INNER one — frame-major input: outer = frame, inner = packed slot; atom-major input: outer =
// packed slot, inner = frame (consecutive threads walk the frames of one atom: contiguous 12 B
// records in the source).
__global__ void __launch_bounds__(256, 8) rdf_pack_kernel(const PackParams p) {
  const bool atom_major = p.layout == MDS_LAYOUT_ATOM_MAJOR_XYZ;
  const long long inner_n = atom_major ? p.n_frames : (long long)p.n_packed;
  const long long outer_n = atom_major ? (long long)p.n_packed : p.n_frames;
  const int lane = threadIdx.x & 31;
  if (p.gate && *p.gate == 0ull) return;  // no frame of this batch needs the copy
  for (long long outer = blockIdx.y; outer < outer_n; outer += gridDim.y) {
    for (long long inner = (long long)blockIdx.x * blockDim.x + threadIdx.x; inner < inner_n;
         inner += (long long)gridDim.x * blockDim.x) {
      const long long f = atom_major ? inner : outer;
      const long long q = atom_major ? outer : inner;
      if (p.need_bits && !(p.frame_flags[f] & p.need_bits)) continue;
      float* dst = p.out + (f * p.frame_stride + (q & ~3LL)) * 3 + (q & 3);
      const long long a = p.src_index[q];
      if (a < 0) {  // padding slot at the end of a species segment: never inside the cutoff
        const float nan = __int_as_float(0x7fc00000);
        dst[0] = nan; dst[4] = nan; dst[8] = nan;
        continue;
      }
      const float* src = atom_major ? p.xyz + (a * p.n_frames_total + p.frame0 + f) * 3
                                    : p.xyz + ((p.frame0 + f) * p.n_atoms + a) * 3;
      const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
      dst[0] = x; dst[4] = y; dst[8] = z;
      uint32_t bits = window_bits(x, p.box[0]) | window_bits(y, p.box[1]) | window_bits(z, p.box[2]);
      const unsigned active = __activemask();
      if (p.cell_off) {
        // cell-list path: rank of the atom inside its (frame, species, cell).  Lanes of a warp that
        // hit the same counter (neighbouring atoms of an ordered trajectory share a cell) are
        // merged with __match_any_sync: one global atomic per distinct counter and warp.
        const int cell = cell_of(x, y, z, p.grid);
        // (item shards: atoms outside this handle's slab of z layers are not part of its lists)
        uint32_t* counter = cell_in_slab(cell, p.grid)
                                ? p.cell_off + f * p.off_stride + p.species[q] * p.grid.nc + cell
                                : nullptr;
        const unsigned peers = __match_any_sync(active, (unsigned long long)counter);
        if (counter) {
          const int leader = __ffs(peers) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(peers));
          base = __shfl_sync(peers, base, leader);
          p.rank[f * p.frame_stride + q] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        }
        // a coordinate further than 8 box lengths out (or infinite): the cell margin no longer
        // covers the float32 rounding of the reference's own arithmetic -> all-pairs for this frame
        if (fabsf(x) > 8.0f * p.box[0] || fabsf(y) > 8.0f * p.box[1] || fabsf(z) > 8.0f * p.box[2])
          bits |= MDS_FRAME_FAR;
      }
      // frame flags: one atomic per warp and frame (a wrapped frame sets a window bit for most of
      // its atoms; per-atom atomics on the frame's word serialise)
      const unsigned same_frame = __match_any_sync(active, f);
      bits = __reduce_or_sync(same_frame, bits);
      if (bits && lane == __ffs(same_frame) - 1 && (p.frame_flags[f] & bits) != bits)
        atomicOr(p.frame_flags + f, bits);
    }
  }
}

cudaError_t launch_pack(const PackParams& p, int sm_count, cudaStream_t stream) {
  const bool atom_major = p.layout == MDS_LAYOUT_ATOM_MAJOR_XYZ;
  const long long inner_n = atom_major ? p.n_frames : (long long)p.n_packed;
  const long long outer_n = atom_major ? (long long)p.n_packed : p.n_frames;
  if (inner_n <= 0 || outer_n <= 0) return cudaSuccess;
  long long bx = (inner_n + 255) / 256;
  const long long fill = (long long)sm_count * 64;  // 8 resident CTAs per SM, 8 waves
  if (bx > fill) bx = fill;
  // enough CTAs to fill the machine, at most 65535 in y
  long long by = (fill + bx - 1) / bx;
  if (by > outer_n) by = outer_n;
  if (by > 65535) by = 65535;
  rdf_pack_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) rdf_count_general_kernel(const uint32_t* flags, int n,
                                                                unsigned long long* out,
                                                                unsigned long long* batch_far) {
  __shared__ unsigned int s_c, s_far;
  if (threadIdx.x == 0) { s_c = 0; s_far = 0; }
  __syncthreads();
  unsigned int c = 0, far = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t f = flags[i];
    c += ((f & 3u) == 3u) ? 1u : 0u;
    far += (f & MDS_FRAME_FAR) ? 1u : 0u;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  far = __reduce_add_sync(0xffffffffu, far);
  if ((threadIdx.x & 31) == 0) {
    if (c) atomicAdd(&s_c, c);
    if (far) atomicAdd(&s_far, far);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_c) atomicAdd(out, (unsigned long long)s_c);
    if (s_far) atomicAdd(out + 1, (unsigned long long)s_far);
    if (batch_far) *batch_far = s_far;  // gate of the all-pairs fallback launch of this batch
  }
}

cudaError_t launch_count_general(const uint32_t* flags, int n, unsigned long long* out,
                                 unsigned long long* batch_far, cudaStream_t stream) {
  rdf_count_general_kernel<<<1, 256, 0, stream>>>(flags, n, out, batch_far);
  return cudaGetLastError();
}

}  // namespace mds
