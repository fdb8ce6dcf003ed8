// Shared declarations of the sm_100a RDF kernels (all-pairs, pack, cell list).
//
// Arithmetic contract (reference utils/linalg.py:84-99 and
// calculators/radial_distribution_function.py:667-689): float32, every op rounded on its own,
// no FMA contraction — hence the explicit __f*_rn intrinsics everywhere a rounding counts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// -DMDS_CHECKED builds a self-checking library (libmds_rdf_checked.so; compute-sanitizer is closed
// on the GPU pool this was developed on).  Every CTA then proves a conservation law at exit: the
// number of shared-memory histogram atomics its threads issued equals what sits in its histogram
// slots (bins + sink slots) plus what it already flushed to the global counts.  An atomic that
// landed anywhere outside the CTA's histogram region — a wild shared-memory address — breaks the
// equality; the CTA prints the two numbers and traps, which fails the launch.
#ifdef MDS_CHECKED
#define MDS_CHK(...) __VA_ARGS__
#else
#define MDS_CHK(...)
#endif

namespace mds {

#ifdef MDS_CHECKED
// called by every thread of the CTA after all counting is done; s_chk[0] = issued, s_chk[1] = found
__device__ __forceinline__ void checked_conservation(unsigned long long* s_chk,
                                                     unsigned long long issued,
                                                     const uint32_t* s_hist, int n_slots, int tid,
                                                     int n_threads, const char* kernel,
                                                     int slot_stride = 1) {
  if (issued) atomicAdd(&s_chk[0], issued);
  unsigned long long part = 0;
  for (int k = tid; k < n_slots; k += n_threads) part += s_hist[k * slot_stride];
  if (part) atomicAdd(&s_chk[1], part);
  __syncthreads();
  if (tid == 0 && s_chk[0] != s_chk[1]) {
    printf("MDS_CHECKED %s block %d: %llu histogram atomics issued, %llu found\n", kernel,
           (int)blockIdx.x, s_chk[0], s_chk[1]);
    __trap();
  }
}
#endif

// One unit of work of the all-pairs kernel: rows [i0, i0+i_cnt) of species a against columns
// [j0, j1) of species b (packed atom indices inside one frame).  When `diag` is set the item
// belongs to a same-species pair and only col > row contributes.
struct PairItem {
  int32_t pair;   // species-pair index (combinations_with_replacement order)
  int32_t i0;
  int32_t i_cnt;
  int32_t j0;
  int32_t j1;
  int32_t diag;
  uint32_t pairs_lo, pairs_hi;  // reference pairs of this item in one frame (64-bit, for statistics)
};
static_assert(sizeof(PairItem) == 32, "PairItem must stay 32 bytes");

// Packed position layout ("quad blocks", SoA in groups of four atoms, 12 B per atom):
//   float pos[n_frames][frame_stride / 4][3][4]   block b = atoms 4b .. 4b+3: x[4], y[4], z[4]
// One LDS.128 of a block row yields two aligned register pairs (x0,x1), (x2,x3) for the packed
// FADD2 / FMUL2 instructions of sm_100a.  Species segments start at multiples of 4 and are
// padded with NaN atoms (never inside the cutoff).
constexpr int QUAD_FLOATS = 12;  // floats per block of four atoms

struct AllPairsParams {
  const float* pos;           // [n_frames][frame_stride / 4][3][4] quad-block positions
  int64_t frame_stride;       // atom slots between frames (multiple of 4)
  int32_t n_frames;
  int32_t n_items;            // items per frame
  const PairItem* items;      // [n_items], sorted by decreasing cost
  const float* edges;         // [nbins + 2] threshold table, see "binning" below
  const uint32_t* frame_flags;  // [n_frames] nonzero: frame is not wrapped -> div/rint minimum image
  unsigned long long* counts; // [n_pairs_total][nbins] global accumulators
  unsigned int* work_counter; // zeroed before launch
  int32_t nbins;
  int32_t pair_lo;            // first species pair held in shared memory by this launch
  int32_t pair_cnt;           // number of species pairs held by this launch
  int32_t hist_in_smem;       // 0: histogram bins live in global memory (very large nbins)
  float s_cut;                // in-cutoff  <=>  s < s_cut
  float inv_step;             // float32(nbins / cutoff * (1 - 2^-20)): seeds the bin guess k~ <= k
  float inv_step_hi;          // float32(nbins / cutoff * (1 + 2^-20)): the opposite guess k^ >= k
  int32_t dense;              // 1: predicated binning (most pairs in cutoff), 0: branch
  float box[3];
  uint32_t frame_need_bits;   // nonzero: only frames whose flags have one of these bits
  const unsigned long long* gate;  // non-null: the whole launch is a no-op when *gate == 0
  unsigned long long* evaluated;   // += reference pairs of the (item, frame) units processed
  // Item sharding (strong scaling within frames, for fewer frames than GPUs): this launch takes
  // the units u with u % shard_count == shard_index of the (item, frame) enumeration.
  uint32_t shard_index, shard_count;
};

// frame classification of the pack kernels
__device__ __forceinline__ uint32_t window_bits(float v, float box) {
  // bit0: outside [-box/8, 9 box/8]   bit1: outside [-5 box/8, 5 box/8]   (NaN: inside both)
  uint32_t b = 0;
  if (v < -0.125f * box || v > 1.125f * box) b |= 1u;
  if (v < -0.625f * box || v > 0.625f * box) b |= 2u;
  return b;
}

// ---- exact float32 building blocks --------------------------------------------------------

// Minimum image for wrapped coordinates (|d| <= 1.25 box): bit-identical to
// d - rint(d / box) * box up to sign, 2 ops instead of 4.
__device__ __forceinline__ float minimg_fast(float d, float box) {
  const float a = fabsf(d);
  return fminf(a, __fsub_rn(box, a));
}

// The reference's literal sequence: RealDiv, Rint, Mul, Sub.
__device__ __forceinline__ float minimg_general(float d, float box) {
  const float q = __fdiv_rn(d, box);
  const float n = rintf(q);
  return __fsub_rn(d, __fmul_rn(n, box));
}

// (x*x + y*y) + z*z with materialised products (tf.linalg.norm before the sqrt).
__device__ __forceinline__ float sq_norm(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// ---- cell assignment of the cell-list path (exact IEEE ops so it is reproducible) -----------
__device__ __forceinline__ int cell_coord(float x, float box, int n) {
  float u = __fdiv_rn(x, box);
  u = __fsub_rn(u, floorf(u));                            // [0, 1]; NaN / inf -> NaN
  const int c = __float2int_rz(__fmul_rn(u, (float)n));   // NaN -> 0
  return min(max(c, 0), n - 1);
}
template <typename Grid>
__device__ __forceinline__ int cell_of(float x, float y, float z, const Grid& g) {
  const int cx = cell_coord(x, g.box[0], g.nx);
  const int cy = cell_coord(y, g.box[1], g.ny);
  const int cz = cell_coord(z, g.box[2], g.nz);
  return (cz * g.ny + cy) * g.nx + cx;
}

// ---- packed float32 pairs (sm_100a FADD2 / FMUL2: two IEEE-rounded ops per issue slot) ------
// ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 (observed with CUDA 12.9,
// even with -fmad=false), which would break the reference's separately rounded products — so
// packed ops are used for the subtractions and the squares only, and the two additions of
// (x*x + y*y) + z*z stay scalar add.rn.f32.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Squared minimum-image distances of ONE row atom (xi, yi, zi) against TWO column atoms, every
// op rounded exactly as in the scalar building blocks above.  Per pair: 1.5 FADD2 (differences)
// + 1.5 FADD2 (box - |d|, abs/neg are operand modifiers) + 3 FMNMX + 1.5 FMUL2 + 2 FADD.
// The row atom comes as register pairs (xi0, xi1) etc.: pair2_s passes the same register twice
// (rows held in registers), pair2_sd two registers that hold the same value (rows broadcast
// from a shared-memory layout that stores every coordinate twice, so that one LDS.128 yields
// the aligned pairs the packed instructions want; rdf_cells.cu).
template <bool FAST>
__device__ __forceinline__ void pair2_sd(float xj0, float xj1, float yj0, float yj1, float zj0,
                                         float zj1, float xi0, float xi1, float yi0, float yi1,
                                         float zi0, float zi1, float bx, float by, float bz,
                                         float& s0, float& s1) {
  float dx0, dx1, dy0, dy1, dz0, dz1;
  unpack2(sub2(pack2(xj0, xj1), pack2(xi0, xi1)), dx0, dx1);
  unpack2(sub2(pack2(yj0, yj1), pack2(yi0, yi1)), dy0, dy1);
  unpack2(sub2(pack2(zj0, zj1), pack2(zi0, zi1)), dz0, dz1);
  float mx0, mx1, my0, my1, mz0, mz1;
  if (FAST) {
    float tx0, tx1, ty0, ty1, tz0, tz1;
    unpack2(sub2(pack2(bx, bx), pack2(fabsf(dx0), fabsf(dx1))), tx0, tx1);
    unpack2(sub2(pack2(by, by), pack2(fabsf(dy0), fabsf(dy1))), ty0, ty1);
    unpack2(sub2(pack2(bz, bz), pack2(fabsf(dz0), fabsf(dz1))), tz0, tz1);
    mx0 = fminf(fabsf(dx0), tx0); mx1 = fminf(fabsf(dx1), tx1);
    my0 = fminf(fabsf(dy0), ty0); my1 = fminf(fabsf(dy1), ty1);
    mz0 = fminf(fabsf(dz0), tz0); mz1 = fminf(fabsf(dz1), tz1);
  } else {
    mx0 = minimg_general(dx0, bx); mx1 = minimg_general(dx1, bx);
    my0 = minimg_general(dy0, by); my1 = minimg_general(dy1, by);
    mz0 = minimg_general(dz0, bz); mz1 = minimg_general(dz1, bz);
  }
  float xx0, xx1, yy0, yy1, zz0, zz1;
  const f32x2 mx = pack2(mx0, mx1), my = pack2(my0, my1), mz = pack2(mz0, mz1);
  unpack2(mul2(mx, mx), xx0, xx1);
  unpack2(mul2(my, my), yy0, yy1);
  unpack2(mul2(mz, mz), zz0, zz1);
  s0 = __fadd_rn(__fadd_rn(xx0, yy0), zz0);
  s1 = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
}

// Same with everything already in packed registers: columns (xj, yj, zj) = two atoms, row
// (xi, yi, zi) = one atom stored twice.  Keeps ptxas from re-assembling the pairs per iteration.
__device__ __forceinline__ void pair2_fast_packed(f32x2 xj, f32x2 yj, f32x2 zj, f32x2 xi, f32x2 yi,
                                                  f32x2 zi, float bx, float by, float bz, float& s0,
                                                  float& s1) {
  float dx0, dx1, dy0, dy1, dz0, dz1;
  unpack2(sub2(xj, xi), dx0, dx1);
  unpack2(sub2(yj, yi), dy0, dy1);
  unpack2(sub2(zj, zi), dz0, dz1);
  float tx0, tx1, ty0, ty1, tz0, tz1;
  unpack2(sub2(pack2(bx, bx), pack2(fabsf(dx0), fabsf(dx1))), tx0, tx1);
  unpack2(sub2(pack2(by, by), pack2(fabsf(dy0), fabsf(dy1))), ty0, ty1);
  unpack2(sub2(pack2(bz, bz), pack2(fabsf(dz0), fabsf(dz1))), tz0, tz1);
  const f32x2 mx = pack2(fminf(fabsf(dx0), tx0), fminf(fabsf(dx1), tx1));
  const f32x2 my = pack2(fminf(fabsf(dy0), ty0), fminf(fabsf(dy1), ty1));
  const f32x2 mz = pack2(fminf(fabsf(dz0), tz0), fminf(fabsf(dz1), tz1));
  float xx0, xx1, yy0, yy1, zz0, zz1;
  unpack2(mul2(mx, mx), xx0, xx1);
  unpack2(mul2(my, my), yy0, yy1);
  unpack2(mul2(mz, mz), zz0, zz1);
  s0 = __fadd_rn(__fadd_rn(xx0, yy0), zz0);
  s1 = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
}

template <bool FAST>
__device__ __forceinline__ void pair2_s(float xj0, float xj1, float yj0, float yj1, float zj0,
                                        float zj1, float xi, float yi, float zi, float bx, float by,
                                        float bz, float& s0, float& s1) {
  pair2_sd<FAST>(xj0, xj1, yj0, yj1, zj0, zj1, xi, xi, yi, yi, zi, zi, bx, by, bz, s0, s1);
}

__device__ __forceinline__ float sqrt_approx(float s) {
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
  return d;
}

// ---- binning ---------------------------------------------------------------------------------
// Threshold table (host-built, rdf_api.cu), nbins + 2 floats:
//   edges[0] = -inf, edges[k] = smallest float32 s that the reference's sqrt +
//   histogram_fixed_width puts in bin >= k (1 <= k < nbins), edges[nbins] = s_cut = smallest s
//   with sqrt_rn(s) >= cutoff, edges[nbins + 1] = +inf.
//   bin k  <=>  edges[k] <= s < edges[k + 1];   "bin" nbins collects everything outside the cutoff.
//
// The bin is found with one approximate step and one exact step: k~ = floor(sqrt.approx(s) *
// inv_step_biased) where inv_step_biased = (nbins / cutoff) * (1 - 2^-20).  The bias is larger
// than every rounding error of the guess (sqrt.approx 2^-22, the constant 2^-24, the reference's
// own sqrt rounding 2^-24) and smaller than 1 / MDS_MAX_BINS, so k~ is the exact bin or one
// below it; a single compare against edges[k~ + 1] settles it.
constexpr float kBinGuessBias = 1.0f - 9.5367431640625e-07f;  // 1 - 2^-20
constexpr uint32_t kMagic4 = 0x4B000000u * 4u;                 // 4 * bits(2^23), mod 2^32
constexpr uint32_t kMagic8 = 0x4B000000u * 8u;                 // 8 * bits(2^23), mod 2^32

// Generic-pointer version for s < s_cut (global-memory histogram fallback).
__device__ __forceinline__ int bin_of_s(float s, float inv_step_biased,
                                        const float* __restrict__ edges) {
  const float f = __fmaf_rz(sqrt_approx(s), inv_step_biased, 8388608.0f);
  const int k = __float_as_int(f) & 0x7fffff;
  return k + ((s >= edges[k + 1]) ? 1 : 0);
}

// Shared-memory paths.  c_edge = smem address of edges[1] minus kMagic4, c_hist = smem address
// of the pair's hist[0] minus kMagic4, four = the constant 4 in a register ptxas cannot see
// through (read back from shared memory): with it the two address computations stay IMADs on the
// FMA pipe instead of LEA / IADD3 on the half-rate ALU pipe, which ncu shows as the busiest one.
//   e = bits(f) * 4 + c_edge = &edges[k~ + 1]      h = bits(f) * 4 + c_hist = &hist[k~]
//
// Dense path: branch-free and unpredicated.  The squared distance is clamped to s_cut (NaN and,
// on diagonal tiles, col <= row pairs clamp to s_cut too), which lands in the overflow slot
// hist[nbins] that is never merged — so the cutoff test costs one FMNMX and every lane runs the
// same instructions.
//
// COND = true adds a second guess with the opposite bias, k^ = floor(sqrt.approx(s) * inv_hi) with
// inv_hi = (nbins / cutoff) * (1 + 2^-20) >= the exact quotient: k~ <= k <= k^, so when the two
// agree (all but ~k * 2^-19 of the lanes) the bin is settled without touching the table and the
// LDS is predicated off.  It trades one FFMA + one ISETP for ~2.5 shared-memory wavefronts per
// warp (32 random table addresses conflict on banks) — worth it where the LSU data pipe is the
// limiter (all-pairs kernel, most pairs inside the cutoff), not where it is idle (cell kernel).
// `hi` is carried by the caller between calls: a lane that skips the load keeps a stale value,
// which the AND with p makes irrelevant (declaring it inside the asm block makes ptxas spill it).
// -DMDS_BIN_MATCHANY builds the variant the north star sketched — explicit warp aggregation with
// __match_any_sync: lanes that hit the same bin elect a leader, which adds the population count
// with one atomic.  Kept for measurement only (scripts/binning_variants.py): the
// default, red.shared.add.u32 -> ATOMS.POPC.INC, lets the hardware merge same-address lanes, costs
// one issue slot instead of MATCH + popc + elect + a predicated atomic, and needs no convergence.
template <bool DIAG, bool COND>
__device__ __forceinline__ void bin_count_dense(float s, float s_cut, float inv_lo, float inv_hi,
                                                uint32_t c_edge, uint32_t c_hist, uint32_t four,
                                                int jg, int ig, float& hi) {
  float t = fminf(s, s_cut);  // min.f32 returns the non-NaN operand
  if (DIAG) t = (jg > ig) ? t : s_cut;
#ifdef MDS_BIN_MATCHANY
  {
    (void)inv_hi; (void)hi;
    const float f = __fmaf_rz(sqrt_approx(t), inv_lo, 8388608.0f);
    const uint32_t a = __float_as_uint(f);
    const uint32_t e = a * four + c_edge;
    uint32_t h = a * four + c_hist;
    float edge;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(edge) : "r"(e));
    h += (t >= edge) ? 4u : 0u;
    const unsigned peers = __match_any_sync(0xffffffffu, h);
    if ((threadIdx.x & 31) == __ffs(peers) - 1)
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(h), "r"((uint32_t)__popc(peers)) : "memory");
    return;
  }
#endif
  if (COND) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .f32 d, f, g;\n"
        ".reg .u32 a, b, e, h;\n"
        "sqrt.approx.ftz.f32 d, %1;\n"
        "fma.rz.f32 f, d, %2, 0f4B000000;\n"
        "fma.rz.f32 g, d, %3, 0f4B000000;\n"
        "mov.b32 a, f;\n"
        "mov.b32 b, g;\n"
        "setp.ne.u32 p, a, b;\n"
        "mad.lo.u32 e, a, %6, %4;\n"
        "mad.lo.u32 h, a, %6, %5;\n"
        "@p ld.shared.f32 %0, [e];\n"
        "setp.ge.and.f32 q, %1, %0, p;\n"
        "@q add.u32 h, h, 4;\n"
        "red.shared.add.u32 [h], 1;\n"
        "}\n"
        : "+f"(hi)
        : "f"(t), "f"(inv_lo), "f"(inv_hi), "r"(c_edge), "r"(c_hist), "r"(four)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .f32 d, f, hi;\n"
        ".reg .u32 a, e, h;\n"
        "sqrt.approx.ftz.f32 d, %0;\n"
        "fma.rz.f32 f, d, %1, 0f4B000000;\n"
        "mov.b32 a, f;\n"
        "mad.lo.u32 e, a, %4, %2;\n"
        "mad.lo.u32 h, a, %4, %3;\n"
        "ld.shared.f32 hi, [e];\n"
        "setp.ge.f32 q, %0, hi;\n"
        "@q add.u32 h, h, 4;\n"
        "red.shared.add.u32 [h], 1;\n"
        "}\n" ::"f"(t),
        "f"(inv_lo), "r"(c_edge), "r"(c_hist), "r"(four)
        : "memory");
  }
}

// Interleaved table (cell kernel, one species pair per launch): slot k of an 8-byte array holds
// edges[k + 1] and, four bytes on, hist[k], so ONE address serves the threshold load and the atomic
// (the increment of the bin becomes "+ 8" and the histogram word is reached through the immediate
// offset of the atomic): one IMAD per pair less than with two arrays.  c_slot = smem address of
// slot 0 minus kMagic8, eight = the constant 8 in a register ptxas cannot see through.
__device__ __forceinline__ void bin_count_dense_interleaved(float s, float s_cut, float inv_lo,
                                                            uint32_t c_slot, uint32_t eight) {
  const float t = fminf(s, s_cut);  // min.f32 returns the non-NaN operand
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .f32 d, f, hi;\n"
      ".reg .u32 a, e;\n"
      "sqrt.approx.ftz.f32 d, %0;\n"
      "fma.rz.f32 f, d, %1, 0f4B000000;\n"
      "mov.b32 a, f;\n"
      "mad.lo.u32 e, a, %3, %2;\n"
      "ld.shared.f32 hi, [e];\n"
      "setp.ge.f32 q, %0, hi;\n"
      "@q add.u32 e, e, 8;\n"
      "red.shared.add.u32 [e+4], 1;\n"
      "}\n" ::"f"(t),
      "f"(inv_lo), "r"(c_slot), "r"(eight)
      : "memory");
}

// Sparse path: the unconditional sequence for a pair already known to be inside the cutoff (under
// a branch).
__device__ __forceinline__ void bin_count(float s, float inv_lo, uint32_t c_edge, uint32_t c_hist,
                                          uint32_t four) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .f32 d, f, hi;\n"
      ".reg .u32 a, e, h;\n"
      "sqrt.approx.ftz.f32 d, %0;\n"
      "fma.rz.f32 f, d, %1, 0f4B000000;\n"
      "mov.b32 a, f;\n"
      "mad.lo.u32 e, a, %4, %2;\n"
      "mad.lo.u32 h, a, %4, %3;\n"
      "ld.shared.f32 hi, [e];\n"
      "setp.ge.f32 q, %0, hi;\n"
      "@q add.u32 h, h, 4;\n"
      "red.shared.add.u32 [h], 1;\n"
      "}\n" ::"f"(s),
      "f"(inv_lo), "r"(c_edge), "r"(c_hist), "r"(four)
      : "memory");
}

// ---- TMA 1-D bulk copy + mbarrier (sm_90+/sm_100a) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes % 16 == 0, 16 B aligned).
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace mds
