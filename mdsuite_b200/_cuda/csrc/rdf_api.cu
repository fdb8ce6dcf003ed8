// C ABI of the RDF hot path (include/mds_rdf.h): handle, threshold table, work schedule,
// stream pipeline.  Host code only; the kernels live in rdf_allpairs.cu / rdf_pack.cu /
// rdf_cells.cu.  There is deliberately no CPU fallback: every entry point that computes
// needs a CUDA device and fails with MDS_ERR_CUDA without one.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>

#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in Nsight Systems, no-ops otherwise

#include "mds_rdf.h"
#include "rdf_common.cuh"
#include "rdf_launch.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(MDS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                  __FILE__, __LINE__);                                                     \
  } while (0)

// ---- the squared-distance threshold table ----------------------------------------------------
// Reference chain for an in-range pair:
//   d = sqrt_rn(s);  keep iff d < float32(cutoff);
//   step = double(float32(cutoff) - 0.0f) / double(B);  bin = int32(min(double(d) / step, B - 1)).
// Every stage is monotone non-decreasing in s, so bin(s) is a step function of s and the whole
// chain collapses to comparisons of s against float32 thresholds computed here, on the host,
// with the same IEEE operations (sqrtf and double division are correctly rounded on x86-64).

inline int ref_bin_of_d(float d, double step, int nbins) {
  double q = (double)d / step;
  const double last = (double)(nbins - 1);
  if (q > last) q = last;
  return (int)q;
}

// smallest float32 s >= 0 with sqrtf(s) >= d  (d >= 0)
float smallest_s_with_sqrt_ge(float d) {
  if (d <= 0.0f) return 0.0f;
  float s = (float)((double)d * (double)d);
  while (s > 0.0f && sqrtf(s) >= d) s = nextafterf(s, -INFINITY);
  while (sqrtf(s) < d) s = nextafterf(s, INFINITY);
  return s;
}

void build_threshold_table(float cutoff, int nbins, std::vector<float>& edges, float& s_cut) {
  const double step = (double)(cutoff - 0.0f) / (double)nbins;
  edges.assign((size_t)nbins + 2, 0.0f);
  edges[0] = -INFINITY;
  for (int k = 1; k < nbins; ++k) {
    float d = (float)((double)k * step);
    while (d > 0.0f && ref_bin_of_d(d, step, nbins) >= k) d = nextafterf(d, -INFINITY);
    while (ref_bin_of_d(d, step, nbins) < k) d = nextafterf(d, INFINITY);
    edges[k] = smallest_s_with_sqrt_ge(d);
  }
  s_cut = smallest_s_with_sqrt_ge(cutoff);
  edges[nbins] = s_cut;  // everything from here on falls into the out-of-cutoff sink
  edges[nbins + 1] = INFINITY;
}

}  // namespace

struct mds_rdf {
  // configuration
  int n_species = 0;
  std::vector<int64_t> counts;
  int nbins = 0;
  float cutoff = 0.f;
  float box[3] = {0, 0, 0};
  uint32_t flags = 0;
  int device = 0;
  int batch_frames = 0;
  int sm_count = 0;
  // derived
  int64_t n_atoms = 0;   // raw atoms per frame
  int32_t n_packed = 0;  // packed slots per frame: Q1 atoms dropped, species padded to multiples of 4
  int32_t n_real = 0;    // packed slots that hold an atom
  int64_t frame_stride = 0;
  int n_pairs = 0;
  double pairs_per_frame = 0;  // the reference's (row, col) count per frame for this mode
  std::vector<float> edges;
  float s_cut = 0.f, inv_step = 0.f, inv_step_hi = 0.f;
  std::vector<int32_t> poff, pcnt;  // packed offsets / counts per species
  int variant = 2;
  std::vector<mds::PairItem> items;
  int pair_group = 1;  // species pairs per launch (shared-memory capacity)
  bool hist_in_smem = true;
  bool dense = true;  // expected in-cutoff fraction is high: predicated binning
  int path = MDS_RDF_PATH_ALLPAIRS;
  mds::CellGrid grid = {0, 0, 0, 0, {0, 0, 0}, 0, 0};
  int kcell[3] = {0, 0, 0};       // stencil half-widths of the cell path (cells per cutoff)
  int cap_atoms = 0;              // positions a CTA of the cell kernel stages in shared memory
  int cell_segments = 1;          // work items per pencil (x segments), see CellsParams.n_seg
  std::vector<int> pair_pieces;   // the same per species pair (launches with one pair use their own)
  int cell_pair_group = 1;        // species pairs per launch of the cell pair kernel
  uint32_t crowded_limit = 0;     // atoms per cell above which a frame goes to the all-pairs kernel
  int64_t off_stride = 0;         // uint32 entries per frame of the cell-offset table
  uint32_t flush_threshold = 1u << 30;
  uint32_t shard_index = 0, shard_count = 1;  // item sharding (mds_rdf_set_item_shard)
  // device state
  int32_t* d_src_index = nullptr;
  int32_t* d_species = nullptr;
  float* d_edges = nullptr;
  mds::PairItem* d_items = nullptr;
  unsigned long long* d_counts = nullptr;
  unsigned long long* d_general = nullptr;
  unsigned int* d_work = nullptr;  // ring of work counters
  int work_slot = 0;
  float* d_packed[2] = {nullptr, nullptr};   // quad-block positions, 12 B per slot
  uint32_t* d_rank[2] = {nullptr, nullptr};  // cell-list path: rank of each atom in its cell
  // Cell lists of a batch: the one-kernel build (a CTA per frame) for batches of at least one wave
  // of its CTAs — below that its CTAs leave most SMs idle and pack / scan / scatter, which spread
  // the atoms of all frames over the device, are faster.  MDS_RDF_CELL_BUILD (tuning / tests):
  // "split" never, "fused" whenever the input is eligible, "eager" = "fused" + the quad-block copy
  // of every frame written by the build instead of only for the frames the all-pairs kernel takes.
  bool build_fused = true;
  bool build_always = false;
  bool lazy_pack = true;
  int build_wave = 0;        // frames per wave of the one-kernel build (0: not eligible)
  int built_fused = 0;       // the last batch went through the one-kernel build
  uint32_t* d_flags[2] = {nullptr, nullptr};
  float* d_raw[2] = {nullptr, nullptr};  // staging for host submits (lazy)
  // cell-list path only
  float4* d_sorted[2] = {nullptr, nullptr};
  uint32_t* d_cell_off[2] = {nullptr, nullptr};
  int4* d_pair_ab = nullptr;
  unsigned long long* d_work64 = nullptr;    // ring of 64-bit work counters
  uint32_t* d_scan_tmp = nullptr;            // scratch of the multi-CTA cell-offset scan
  unsigned long long* d_evaluated = nullptr; // [2] pairs evaluated (candidates on the cell path) / tested
  unsigned long long* d_batch_far = nullptr; // [2] far frames of the batch in each buffer
  size_t raw_bytes = 0;
  // pageable host submits: frames are copied (and, for the atom-major layout, transposed to frame
  // major) into a ring of pinned buffers by the calling thread and sent on when a buffer is full
  float* stage[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_stage_done[3] = {nullptr, nullptr, nullptr};  // H2D out of the buffer finished
  // tickets: one event per host submit, recorded on the copy stream after its last H2D
  cudaEvent_t ev_ticket[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t tickets = 0;  // host submits whose copies have been enqueued so far
  int stage_slot = 0;
  int64_t stage_frames = 0;     // frames waiting in stage[stage_slot]
  int64_t stage_capacity = 0;   // frames per buffer
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_raw_ready[2] = {nullptr, nullptr};  // H2D of staging buffer done
  cudaEvent_t ev_raw_free[2] = {nullptr, nullptr};   // pack consumed staging buffer
  cudaEvent_t ev_ext = nullptr;                      // caller stream -> compute stream
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
  cudaEvent_t ev_join = nullptr;
  std::vector<cudaEvent_t> pk_events;  // start/stop pairs around every pair-kernel launch
  int pk_used = 0;                     // launches of the last submit
  bool timing_valid = false;
  int buf = 0;
  // statistics
  int64_t frames_done = 0;
  int64_t launches = 0;
};

namespace {

constexpr int WORK_RING = 4096;

struct NvtxRange {  // tracing hook the reference lacks (timers at DEBUG level only)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

void build_items(mds_rdf* h, int tile_rows, int64_t frames_hint) {
  h->items.clear();
  // column chunk: keep items big enough to amortise their prologue, small enough that
  // frames_hint * n_items gives every persistent CTA several items.
  int pair = 0;
  std::vector<mds::PairItem> items;
  for (int a = 0; a < h->n_species; ++a) {
    for (int b = a; b < h->n_species; ++b, ++pair) {
      // segments start at multiples of 4 and are NaN-padded to multiples of 4, so column ranges
      // can always be whole quad blocks; the diagonal block starts AT the row tile (col > row
      // is applied per pair inside the kernel)
      const int a0 = h->poff[a], a1 = a0 + h->pcnt[a];
      const int b0 = h->poff[b], b1 = b0 + ((h->pcnt[b] + 3) & ~3);
      if (h->pcnt[a] == 0 || h->pcnt[b] == 0) continue;
      if (a == b && h->pcnt[a] < 2) continue;
      for (int i0 = a0; i0 < a1; i0 += tile_rows) {
        const int i_cnt = std::min(tile_rows, a1 - i0);
        const int j0 = (a == b) ? i0 : b0;
        const int j1 = b1;
        if (j0 >= j1) continue;
        items.push_back({pair, i0, i_cnt, j0, j1, (a == b) ? 1 : 0, 0, 0});
      }
    }
  }
  // split long column ranges until there is enough parallel slack
  const long long want = 8LL * h->sm_count * 3;
  int chunk = 1 << 30;
  auto count_items = [&](int c) {
    long long n = 0;
    for (auto& it : items) n += ((it.j1 - it.j0) + c - 1) / c;
    return n;
  };
  if (!items.empty()) {
    int max_cols = 0;
    for (auto& it : items) max_cols = std::max(max_cols, it.j1 - it.j0);
    chunk = std::max(512, ((max_cols + 511) / 512) * 512);
    while (chunk > 2048 && count_items(chunk) * std::max<int64_t>(frames_hint, 1) < want)
      chunk = std::max(2048, ((chunk / 2 + 511) / 512) * 512);
    if (chunk > 8192) chunk = 8192;
  }
  // species of every pair index, to find where the real (unpadded) columns of an item end
  std::vector<int> pair_b;
  for (int a = 0; a < h->n_species; ++a)
    for (int b = a; b < h->n_species; ++b) pair_b.push_back(b);
  for (auto& it : items) {
    for (int j = it.j0; j < it.j1; j += chunk) {
      mds::PairItem c = it;
      c.j0 = j;
      c.j1 = std::min(it.j1, j + chunk);
      // reference pairs of this item: real rows x real columns, col > row on same-species items
      const int b = pair_b[(size_t)c.pair];
      const long long c1 = std::min<long long>(c.j1, (long long)h->poff[b] + h->pcnt[b]);
      unsigned long long pairs = 0;
      if (c1 > c.j0) {
        if (!c.diag) {
          pairs = (unsigned long long)c.i_cnt * (unsigned long long)(c1 - c.j0);
        } else {
          for (long long r = c.i0; r < (long long)c.i0 + c.i_cnt; ++r) {
            const long long lo = std::max<long long>(c.j0, r + 1);
            if (c1 > lo) pairs += (unsigned long long)(c1 - lo);
          }
        }
      }
      c.pairs_lo = (uint32_t)(pairs & 0xffffffffull);
      c.pairs_hi = (uint32_t)(pairs >> 32);
      h->items.push_back(c);
    }
  }
  std::stable_sort(h->items.begin(), h->items.end(), [](const mds::PairItem& x, const mds::PairItem& y) {
    return (long long)x.i_cnt * (x.j1 - x.j0) > (long long)y.i_cnt * (y.j1 - y.j0);
  });
}

int pick_variant(const mds_rdf* h) {
  if (const char* env = getenv("MDS_RDF_VARIANT")) {
    int v = atoi(env);
    if (v >= 0 && v < mds::ALLPAIRS_N_VARIANTS) return v;
  }
  int max_species = 0;
  for (int c : h->pcnt) max_species = std::max(max_species, c);
  if (max_species >= 4096) return 2;  // 256 threads x 4 rows
  if (max_species >= 1024) return 1;  // 256 x 2
  return 0;                           // 256 x 1
}

int upload_items(mds_rdf* h) {
  if (h->d_items) { cudaFree(h->d_items); h->d_items = nullptr; }
  if (h->items.empty()) return MDS_OK;
  CUDA_TRY(cudaMalloc(&h->d_items, h->items.size() * sizeof(mds::PairItem)));
  CUDA_TRY(cudaMemcpy(h->d_items, h->items.data(), h->items.size() * sizeof(mds::PairItem),
                      cudaMemcpyHostToDevice));
  return MDS_OK;
}

mds::AllPairsParams allpairs_params(mds_rdf* h, int buf, int64_t n_frames, int p0, int pc) {
  mds::AllPairsParams ap;
  ap.pos = h->d_packed[buf];
  ap.frame_stride = h->frame_stride;
  ap.n_frames = (int32_t)n_frames;
  ap.edges = h->d_edges;
  ap.frame_flags = h->d_flags[buf];
  ap.counts = h->d_counts;
  ap.nbins = h->nbins;
  ap.pair_lo = p0;
  ap.pair_cnt = pc;
  ap.hist_in_smem = h->hist_in_smem ? 1 : 0;
  ap.s_cut = h->s_cut;
  ap.inv_step = h->inv_step;
  ap.inv_step_hi = h->inv_step_hi;
  ap.dense = h->dense ? 1 : 0;
#ifdef MDS_CHECKED
  if (const char* env = getenv("MDS_CHECKED_INJECT_FAULT")) ap.dense |= atoi(env) ? 2 : 0;
#endif
  for (int k = 0; k < 3; ++k) ap.box[k] = h->box[k];
  ap.frame_need_bits = 0;
  ap.gate = nullptr;
  ap.evaluated = h->d_evaluated;
  ap.shard_index = h->shard_index;
  ap.shard_count = h->shard_count;
  ap.items = nullptr;
  ap.n_items = 0;
  ap.work_counter = nullptr;
  return ap;
}

// item sub-range of a pair group (items are grouped per pair group at create time)
bool item_range(const mds_rdf* h, int p0, int pc, int* first, int* count) {
  int lo = -1, hi = -1;
  for (int i = 0; i < (int)h->items.size(); ++i)
    if (h->items[i].pair >= p0 && h->items[i].pair < p0 + pc) {
      if (lo < 0) lo = i;
      hi = i;
    }
  if (lo < 0) return false;
  *first = lo;
  *count = hi - lo + 1;
  return true;
}

int next_work_slot(mds_rdf* h) {
  if (h->work_slot >= WORK_RING) {
    cudaError_t e = cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int) * WORK_RING, h->s_compute);
    if (e == cudaSuccess && h->d_work64)
      e = cudaMemsetAsync(h->d_work64, 0, sizeof(unsigned long long) * WORK_RING, h->s_compute);
    if (e != cudaSuccess) return -1;
    h->work_slot = 0;
  }
  return h->work_slot++;
}

// At most PK_MAX launches are timed between two mds_rdf_stats_get calls: a long-lived handle whose
// owner never asks for statistics must not pile up events without bound.
constexpr int PK_MAX = 16384;

int timed_launch_begin(mds_rdf* h) {
  if (h->pk_used >= PK_MAX) return MDS_OK;
  if ((int)h->pk_events.size() < 2 * (h->pk_used + 1)) {
    cudaEvent_t a = nullptr, b = nullptr;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    h->pk_events.push_back(a);
    h->pk_events.push_back(b);
  }
  CUDA_TRY(cudaEventRecord(h->pk_events[2 * h->pk_used], h->s_compute));
  return MDS_OK;
}

int timed_launch_end(mds_rdf* h) {
  h->launches += 1;
  if (h->pk_used >= PK_MAX) return MDS_OK;
  CUDA_TRY(cudaEventRecord(h->pk_events[2 * h->pk_used + 1], h->s_compute));
  h->pk_used += 1;
  return MDS_OK;
}

mds::PackParams pack_params(mds_rdf* h, const float* d_xyz, int layout, int64_t frame0,
                            int64_t n_frames, int64_t n_frames_total, int buf) {
  mds::PackParams pp;
  pp.xyz = d_xyz;
  pp.layout = layout;
  pp.n_atoms = h->n_atoms;
  pp.n_frames = n_frames;
  pp.frame0 = frame0;
  pp.n_frames_total = n_frames_total;
  pp.src_index = h->d_src_index;
  pp.species = h->d_species;
  pp.n_packed = h->n_packed;
  pp.out = h->d_packed[buf];
  pp.frame_stride = h->frame_stride;
  pp.frame_flags = h->d_flags[buf];
  for (int k = 0; k < 3; ++k) pp.box[k] = h->box[k];
  pp.cell_off = nullptr;
  pp.off_stride = 0;
  pp.rank = nullptr;
  pp.gate = nullptr;
  pp.need_bits = 0;
  pp.grid = h->grid;
  if (h->path == MDS_RDF_PATH_CELLLIST && h->shard_count > 1 && (uint32_t)h->grid.nz >= h->shard_count) {
    // only the handle's slab of z layers and the layers its stencil reaches are sorted
    const long long nz = h->grid.nz, z0 = nz * h->shard_index / h->shard_count,
                    z1 = nz * (h->shard_index + 1) / h->shard_count;
    const int kz = h->kcell[2];
    pp.grid.z_len = (z1 > z0) ? (int)std::min<long long>(nz, (z1 - z0) + 2 * kz) : 0;
    pp.grid.z_lo = (int)(((z0 - kz) % nz + nz) % nz);
  }
  if (h->path == MDS_RDF_PATH_CELLLIST) {
    pp.cell_off = h->d_cell_off[buf];
    pp.off_stride = h->off_stride;
    pp.rank = h->d_rank[buf];
  }
  return pp;
}

// all-pairs path: pack one batch that already sits in device memory, then run the pair kernels
int process_batch_allpairs(mds_rdf* h, const float* d_xyz, int layout, int64_t frame0,
                           int64_t n_frames, int64_t n_frames_total, int buf) {
  CUDA_TRY(cudaMemsetAsync(h->d_flags[buf], 0, sizeof(uint32_t) * (size_t)n_frames, h->s_compute));
  const mds::PackParams pp = pack_params(h, d_xyz, layout, frame0, n_frames, n_frames_total, buf);
  CUDA_TRY(mds::launch_pack(pp, h->sm_count, h->s_compute));
  CUDA_TRY(mds::launch_count_general(h->d_flags[buf], (int)n_frames, h->d_general, nullptr,
                                     h->s_compute));
  h->launches += 2;
  if (h->items.empty()) return MDS_OK;
  for (int p0 = 0; p0 < h->n_pairs; p0 += h->pair_group) {
    const int pc = std::min(h->pair_group, h->n_pairs - p0);
    mds::AllPairsParams ap = allpairs_params(h, buf, n_frames, p0, pc);
    int first = 0, count = 0;
    if (!item_range(h, p0, pc, &first, &count)) continue;
    ap.items = h->d_items + first;
    ap.n_items = count;
    const int slot = next_work_slot(h);
    if (slot < 0) return fail(MDS_ERR_CUDA, "work-counter reset failed");
    ap.work_counter = h->d_work + slot;
    int rc = timed_launch_begin(h);
    if (rc != MDS_OK) return rc;
    CUDA_TRY(mds::launch_allpairs(ap, h->variant, h->sm_count, h->s_compute, nullptr));
    rc = timed_launch_end(h);
    if (rc != MDS_OK) return rc;
  }
  return MDS_OK;
}

// cell-list path: pack + counting sort by (species, cell), then the cell pair kernel; frames the
// pack kernel flagged MDS_FRAME_FAR go through the all-pairs kernel (gated: normally a no-op)
int process_batch_cells(mds_rdf* h, const float* d_xyz, int layout, int64_t frame0,
                        int64_t n_frames, int64_t n_frames_total, int buf) {
  CUDA_TRY(cudaMemsetAsync(h->d_flags[buf], 0, sizeof(uint32_t) * (size_t)n_frames, h->s_compute));
  const mds::PackParams pp = pack_params(h, d_xyz, layout, frame0, n_frames, n_frames_total, buf);
  const bool one_kernel = h->build_fused &&
                          mds::cells_build_fused(layout, h->n_species * h->grid.nc) &&
                          (h->build_always || (h->build_wave > 0 && n_frames >= h->build_wave));
  const bool lazy_pack = one_kernel && h->lazy_pack;
  CUDA_TRY(mds::launch_cell_build(pp, h->n_species, h->d_sorted[buf], h->crowded_limit, h->d_scan_tmp,
                                  h->sm_count, one_kernel, !lazy_pack, h->s_compute));
  CUDA_TRY(mds::launch_count_general(h->d_flags[buf], (int)n_frames, h->d_general,
                                     h->d_batch_far + buf, h->s_compute));
  h->built_fused = one_kernel ? 1 : 0;
  h->launches += one_kernel ? 2 : 4;
  if (lazy_pack && !h->items.empty()) {
    // the frames the all-pairs kernel will take (flagged MDS_FRAME_FAR; normally none, and then
    // this launch returns at once) get their quad-block copy now
    mds::PackParams far = pp;
    far.cell_off = nullptr;
    far.rank = nullptr;
    far.gate = h->d_batch_far + buf;
    far.need_bits = mds::MDS_FRAME_FAR;
    CUDA_TRY(mds::launch_pack(far, h->sm_count, h->s_compute));
    h->launches += 1;
  }
  if (h->n_real == 0) return MDS_OK;
  for (int p0 = 0; p0 < h->n_pairs; p0 += h->cell_pair_group) {
    const int pc = std::min(h->cell_pair_group, h->n_pairs - p0);
    mds::CellsParams cp;
    cp.spos = h->d_sorted[buf];
    cp.frame_stride = h->frame_stride;
    cp.n_frames = (int32_t)n_frames;
    cp.cell_off = h->d_cell_off[buf];
    cp.off_stride = h->off_stride;
    cp.frame_flags = h->d_flags[buf];
    cp.nx = h->grid.nx; cp.ny = h->grid.ny; cp.nz = h->grid.nz; cp.nc = h->grid.nc;
    cp.kx = h->kcell[0]; cp.ky = h->kcell[1]; cp.kz = h->kcell[2];
    cp.cap_atoms = h->cap_atoms;
    cp.pair_ab = h->d_pair_ab;
    cp.edges = h->d_edges;
    cp.counts = h->d_counts;
    cp.evaluated = h->d_evaluated;
    cp.nbins = h->nbins;
    cp.pair_lo = p0;
    cp.pair_cnt = pc;
    cp.hist_in_smem = h->hist_in_smem ? 1 : 0;
    cp.s_cut = h->s_cut;
    cp.inv_step = h->inv_step;
    for (int k = 0; k < 3; ++k) cp.box[k] = h->box[k];
    cp.flush_threshold = h->flush_threshold;
    // item shards: a slab of z layers per handle, [nz * index / count, nz * (index + 1) / count),
    // if there are enough layers; else every count-th item
    cp.pen_lo = 0;
    cp.n_pen = h->grid.ny * h->grid.nz;
    cp.n_seg = (pc == 1) ? h->pair_pieces[p0] : h->cell_segments;
    cp.shard_index = h->shard_index;
    cp.shard_count = h->shard_count;
    if (h->shard_count > 1 && (uint32_t)h->grid.nz >= h->shard_count) {
      const long long nz = h->grid.nz, z0 = nz * h->shard_index / h->shard_count,
                      z1 = nz * (h->shard_index + 1) / h->shard_count;
      cp.pen_lo = (int32_t)(z0 * h->grid.ny);
      cp.n_pen = (int32_t)((z1 - z0) * h->grid.ny);
      cp.shard_index = 0;
      cp.shard_count = 1;
    }
    const int slot = next_work_slot(h);
    if (slot < 0) return fail(MDS_ERR_CUDA, "work-counter reset failed");
    cp.work_counter = h->d_work64 + slot;
    int rc = timed_launch_begin(h);
    if (rc != MDS_OK) return rc;
    CUDA_TRY(mds::launch_cells(cp, h->sm_count, h->s_compute, nullptr));
    rc = timed_launch_end(h);
    if (rc != MDS_OK) return rc;
  }
  if (h->items.empty()) return MDS_OK;
  for (int p0 = 0; p0 < h->n_pairs; p0 += h->pair_group) {
    const int pc = std::min(h->pair_group, h->n_pairs - p0);
    mds::AllPairsParams ap = allpairs_params(h, buf, n_frames, p0, pc);
    int first = 0, count = 0;
    if (!item_range(h, p0, pc, &first, &count)) continue;
    ap.items = h->d_items + first;
    ap.n_items = count;
    ap.frame_need_bits = mds::MDS_FRAME_FAR;
    ap.gate = h->d_batch_far + buf;
    const int slot = next_work_slot(h);
    if (slot < 0) return fail(MDS_ERR_CUDA, "work-counter reset failed");
    ap.work_counter = h->d_work + slot;
    CUDA_TRY(mds::launch_allpairs(ap, h->variant, h->sm_count, h->s_compute, nullptr));
    h->launches += 1;
  }
  return MDS_OK;
}

int process_batch(mds_rdf* h, const float* d_xyz, int layout, int64_t frame0, int64_t n_frames,
                  int64_t n_frames_total, int buf) {
  NvtxRange range(h->path == MDS_RDF_PATH_CELLLIST ? "mds_rdf batch (cell list)"
                                                   : "mds_rdf batch (all pairs)");
  return h->path == MDS_RDF_PATH_CELLLIST
             ? process_batch_cells(h, d_xyz, layout, frame0, n_frames, n_frames_total, buf)
             : process_batch_allpairs(h, d_xyz, layout, frame0, n_frames, n_frames_total, buf);
}

}  // namespace

extern "C" {

const char* mds_last_error(void) { return g_last_error.c_str(); }
int mds_abi_version(void) { return MDS_RDF_ABI_VERSION; }

int mds_rdf_create(mds_rdf_t** out, const mds_rdf_config* cfg) {
  if (!out) return fail(MDS_ERR_INVALID, "mds_rdf_create: out is NULL");
  *out = nullptr;
  if (!cfg) return fail(MDS_ERR_INVALID, "mds_rdf_create: cfg is NULL");
  if (cfg->struct_size != sizeof(mds_rdf_config))
    return fail(MDS_ERR_INVALID, "mds_rdf_create: struct_size %u != %zu (ABI mismatch)",
                cfg->struct_size, sizeof(mds_rdf_config));
  if (cfg->n_species < 1 || !cfg->counts)
    return fail(MDS_ERR_INVALID, "mds_rdf_create: need n_species >= 1 and counts");
  if (cfg->nbins < 1 || cfg->nbins > mds::MDS_MAX_BINS)
    return fail(MDS_ERR_INVALID, "mds_rdf_create: nbins %d outside [1, %d]", cfg->nbins,
                mds::MDS_MAX_BINS);
  if (!(cfg->cutoff > 0.f) || !std::isfinite(cfg->cutoff))
    return fail(MDS_ERR_INVALID, "mds_rdf_create: cutoff must be finite and > 0");
  for (int k = 0; k < 3; ++k)
    if (!(cfg->box[k] > 0.f) || !std::isfinite(cfg->box[k]))
      return fail(MDS_ERR_INVALID, "mds_rdf_create: box[%d] must be finite and > 0", k);
  if ((cfg->flags & MDS_RDF_PATH_ALLPAIRS) && (cfg->flags & MDS_RDF_PATH_CELLLIST))
    return fail(MDS_ERR_INVALID, "mds_rdf_create: both PATH flags set");
  int64_t n_atoms = 0;
  for (int s = 0; s < cfg->n_species; ++s) {
    if (cfg->counts[s] < 0) return fail(MDS_ERR_INVALID, "mds_rdf_create: counts[%d] < 0", s);
    n_atoms += cfg->counts[s];
  }
  if (n_atoms > (int64_t)std::numeric_limits<int32_t>::max() / 2)
    return fail(MDS_ERR_UNSUPPORTED, "mds_rdf_create: %lld atoms per frame is too many",
                (long long)n_atoms);

  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || n_dev < 1)
    return fail(MDS_ERR_CUDA, "mds_rdf_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(ce));
  if (cfg->device < 0 || cfg->device >= n_dev)
    return fail(MDS_ERR_INVALID, "mds_rdf_create: device %d not in [0, %d)", cfg->device, n_dev);

  mds_rdf* h = new (std::nothrow) mds_rdf();
  if (!h) return fail(MDS_ERR_NOMEM, "mds_rdf_create: out of host memory");
  h->n_species = cfg->n_species;
  h->counts.assign(cfg->counts, cfg->counts + cfg->n_species);
  h->nbins = cfg->nbins;
  h->cutoff = cfg->cutoff;
  for (int k = 0; k < 3; ++k) h->box[k] = cfg->box[k];
  h->flags = cfg->flags;
  h->device = cfg->device;
  h->n_atoms = n_atoms;
  h->n_pairs = cfg->n_species * (cfg->n_species + 1) / 2;

  DeviceGuard guard(h->device);
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, h->device);
  if (ce != cudaSuccess) {
    delete h;
    return fail(MDS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce));
  }
  if (prop.major < 10) {
    delete h;
    return fail(MDS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                h->device, prop.major, prop.minor);
  }
  h->sm_count = prop.multiProcessorCount;

  // packed order: species concatenated, the Q1 atom of each species dropped in parity mode
  const int first = (h->flags & MDS_RDF_PARITY_SKIP_FIRST) ? 1 : 0;
  std::vector<int32_t> src_index, species;
  int64_t off = 0;
  h->pairs_per_frame = 0;
  for (int s = 0; s < h->n_species; ++s) {
    h->poff.push_back((int32_t)src_index.size());  // a multiple of 4
    const int64_t n = std::max<int64_t>(0, h->counts[s] - first);
    h->pcnt.push_back((int32_t)n);
    for (int64_t i = 0; i < n; ++i) {
      src_index.push_back((int32_t)(off + first + i));
      species.push_back(s);
    }
    while (src_index.size() & 3) {  // NaN padding slots
      src_index.push_back(-1);
      species.push_back(s);
    }
    h->n_real += (int32_t)n;
    off += h->counts[s];
  }
  for (int a = 0; a < h->n_species; ++a)
    for (int b = a; b < h->n_species; ++b)
      h->pairs_per_frame += (a == b) ? 0.5 * (double)h->pcnt[a] * (double)(h->pcnt[a] - 1)
                                     : (double)h->pcnt[a] * (double)h->pcnt[b];
  h->n_packed = (int32_t)src_index.size();
  h->frame_stride = ((int64_t)h->n_packed + 31) & ~31LL;  // 384 B aligned frames
  if (h->frame_stride == 0) h->frame_stride = 32;

  build_threshold_table(h->cutoff, h->nbins, h->edges, h->s_cut);
  h->inv_step = (float)((double)h->nbins / (double)h->cutoff * (double)mds::kBinGuessBias);
  h->inv_step_hi = (float)((double)h->nbins / (double)h->cutoff / (double)mds::kBinGuessBias);

  // ---- path: all-pairs tiles or cell list ------------------------------------------------
  // The cell path needs >= 3 cutoff-sized (1.001 x cutoff) cells per dimension.
  // Its grid is finer: K cells per cutoff length (K = 1 or 2 per dimension, the stencil
  // half-width of rdf_cells.cu), at most 128 K per dimension — the bound the margin rests on.
  {
    int n1[3];
    bool cell_ok = true;
    for (int k = 0; k < 3; ++k) {
      n1[k] = (int)std::min(128.0, std::floor((double)h->box[k] / ((double)h->cutoff * 1.001)));
      if (n1[k] < 3) cell_ok = false;
    }
    const long long nc1 = cell_ok ? (long long)n1[0] * n1[1] * n1[2] : 0;
    const bool forced_cells = (cfg->flags & MDS_RDF_PATH_CELLLIST) != 0;
    const bool forced_all = (cfg->flags & MDS_RDF_PATH_ALLPAIRS) != 0;
    if (forced_cells && !cell_ok) {
      delete h;
      return fail(MDS_ERR_UNSUPPORTED,
                  "mds_rdf_create: cell-list path needs cutoff <= box / 3.003 in every dimension "
                  "(cutoff %g, box %g %g %g)", (double)cfg->cutoff, (double)cfg->box[0],
                  (double)cfg->box[1], (double)cfg->box[2]);
    }
    // the cell kernel tests a small fraction of all pairs at roughly half the all-pairs kernel's rate
    const bool auto_cells = cell_ok && nc1 >= 64 && h->n_real >= 1024;
    h->path = (forced_cells || (!forced_all && auto_cells)) ? MDS_RDF_PATH_CELLLIST
                                                            : MDS_RDF_PATH_ALLPAIRS;
    if (h->path == MDS_RDF_PATH_CELLLIST) {
      // Finer cells along x (the direction the kernel sweeps): with kx cells per cutoff length the
      // partners of an atom lie in 2 + 1/kx cutoff lengths of x instead of 3, at the price of
      // fewer rows per cell.  Measured on the BASELINE shapes: about 8 atoms of the
      // most populous species per cell is the sweet spot; y and z stay cutoff-sized — halving them
      // too cuts the candidates further (15.6 instead of 22.5 cutoff-cell volumes) but quadruples
      // the staging work per row atom, which costs more than it saves unless the cells are full.
      // MDS_RDF_CELL_K = k | "kx,ky,kz" overrides (kx <= 4, ky, kz <= 2).
      int max_pcnt = 0;
      for (int c : h->pcnt) max_pcnt = std::max(max_pcnt, c);
      const double per_cutoff_cell = (double)max_pcnt / (double)nc1;
      int K[3] = {std::min(4, std::max(1, (int)std::floor(per_cutoff_cell / 8.0 + 0.5))), 1, 1};
      // dense systems (>= 6 atoms per half-width cell, e.g. the 50k-atom gas with cutoff 10): there
      // the half-width grid in all three dimensions is ahead
      if (per_cutoff_cell >= 48.0) K[0] = K[1] = K[2] = 2;
      if (const char* env = getenv("MDS_RDF_CELL_K")) {
        int a = 0, b = 0, c = 0;
        const int got = sscanf(env, "%d,%d,%d", &a, &b, &c);
        if (got == 1) b = c = a;
        if ((got == 1 || got == 3) && a >= 1 && a <= 4 && b >= 1 && b <= 2 && c >= 1 && c <= 2) {
          K[0] = a; K[1] = b; K[2] = c;
        }
      }
      int n[3];
      for (int k = 0; k < 3; ++k) {
        const double w_min = (double)h->cutoff * 1.001 / (double)K[k];
        n[k] = (int)std::min(128.0 * K[k], std::floor((double)h->box[k] / w_min));
        // (n1 >= 3 implies n >= 2 K + 1: the cells of a stencil are distinct)
        h->kcell[k] = K[k];
      }
      const long long nc = (long long)n[0] * n[1] * n[2];
      h->grid.nx = n[0]; h->grid.ny = n[1]; h->grid.nz = n[2];
      h->grid.nc = (int)nc;
      h->grid.z_lo = 0;
      h->grid.z_len = n[2];
      for (int k = 0; k < 3; ++k) h->grid.box[k] = h->box[k];
      h->off_stride = (((int64_t)h->n_species * nc + 1) + 3) & ~3LL;
      h->cap_atoms = 2048;
      if (const char* env = getenv("MDS_RDF_CELL_CAP")) {
        const int v = atoi(env);
        if (v >= 256 && v <= mds::cells_max_cap()) h->cap_atoms = v;
      }
      if (const char* env = getenv("MDS_RDF_CELL_BUILD")) {
        h->build_fused = strcmp(env, "split") != 0;
        h->build_always = strcmp(env, "fused") == 0 || strcmp(env, "eager") == 0;
        h->lazy_pack = strcmp(env, "eager") != 0;
      }
      // one row cell (+ its kx own-pencil neighbours) against one neighbour pencil must fit
      h->crowded_limit = (uint32_t)(h->cap_atoms / (3 * K[0] + 2));
    }
    if (const char* env = getenv("MDS_RDF_FLUSH_THRESHOLD")) {
      const long long v = atoll(env);
      if (v >= 64 && v <= (1LL << 30)) h->flush_threshold = (uint32_t)v;
    }
  }
  const bool cells = h->path == MDS_RDF_PATH_CELLLIST;

  // how many species pairs fit in shared memory next to the threshold table
  const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin;
  const size_t target = std::min<size_t>(smem_cap, 100 * 1024);  // keep >= 2 CTAs per SM if possible
  auto smem_need = [&](int pairs) {
    // the all-pairs kernel also serves the cell path's fallback launches: size for the larger
    const size_t a = mds::allpairs_smem_bytes(h->nbins, pairs, true);
    return cells ? std::max(a, mds::cells_smem_bytes(h->nbins, pairs, true, h->cap_atoms)) : a;
  };
  h->hist_in_smem = true;
  h->pair_group = h->n_pairs;
  while (h->pair_group > 1 && smem_need(h->pair_group) > target) --h->pair_group;
  bool group_forced = false;
  if (const char* env = getenv("MDS_RDF_PAIR_GROUP")) {  // tuning aid
    const int v = atoi(env);
    if (v >= 1 && v <= h->n_pairs && smem_need(v) <= smem_cap) {
      h->pair_group = v;
      group_forced = true;
    }
  }
  if (smem_need(h->pair_group) > target) {
    // a single pair does not fit the 2-CTA budget: allow the full opt-in size, then give up on smem
    if (smem_need(1) > smem_cap) {
      h->hist_in_smem = false;
      h->pair_group = h->n_pairs;
    }
  }
  // The cell pair kernel takes ONE species pair per launch when there are few of them: its table
  // and histogram then share one interleaved array (rdf_cells.cu), a CTA needs less shared memory
  // (cfg4, 800 bins x 3 pairs: 4 resident CTAs per SM instead of 3) and a launch has the work items
  // of its own segment length only — 10 % on cfg4.  With many pairs the launches would get too
  // short; those keep the grouping of the all-pairs kernel.
  h->cell_pair_group = (h->hist_in_smem && h->n_pairs <= 6 && !group_forced) ? 1 : h->pair_group;

  {
    // expected fraction of pairs inside the cutoff sphere (ideal gas): 4/3 pi rc^3 / V
    const double vol = (double)h->box[0] * h->box[1] * h->box[2];
    const double phi = std::min(1.0, 4.18879 * (double)h->cutoff * h->cutoff * h->cutoff / vol);
    h->dense = phi > 0.02;  // measured crossover (scripts/dense_probe.py): 1.4 % branchy wins, 3.4 % dense wins
    if (const char* env = getenv("MDS_RDF_DENSE")) h->dense = atoi(env) != 0;
  }
  h->variant = pick_variant(h);
  if (cells && h->build_fused)
    h->build_wave = mds::cells_build_frames_per_wave(MDS_LAYOUT_FRAME_MAJOR_XYZ,
                                                     h->n_species * h->grid.nc, h->sm_count);
  int bf = cfg->max_frames_in_flight;
  if (bf <= 0) {
    const int64_t per_frame = h->frame_stride * (cells ? 12 + 4 + 16 : 12) + h->off_stride * 4;
    // 128 MB of internal buffers per batch; 512 MB on the cell path, whose pair kernel spreads
    // (frame, pencil) items over persistent CTAs and loses ~3 % to ramp and tail on short batches
    bf = (int)std::min<int64_t>(4096, std::max<int64_t>(1, ((cells ? 512LL : 128LL) << 20) / per_frame));
    // the one-kernel build takes a frame per CTA: a batch is a whole number of its waves (1 GB of
    // buffers at most), so that no SM is left with one more frame than the others
    const int64_t most = std::min<int64_t>(4096, (1024LL << 20) / per_frame);
    if (h->build_wave > 0 && most >= h->build_wave) bf = (int)(most / h->build_wave * h->build_wave);
  }
  h->batch_frames = bf;
  build_items(h, mds::allpairs_variant_tile(h->variant), bf);
  {
    double sum = 0;
    for (const auto& it : h->items) sum += (double)(((unsigned long long)it.pairs_hi << 32) | it.pairs_lo);
    if (sum != h->pairs_per_frame) {
      const int rc = fail(MDS_ERR_INVALID, "internal: work items cover %.0f pairs, expected %.0f", sum,
                          h->pairs_per_frame);
      delete h;
      return rc;
    }
  }
  // group items by pair group so a launch sees a contiguous range (stable: keeps cost order)
  std::stable_sort(h->items.begin(), h->items.end(), [h](const mds::PairItem& x, const mds::PairItem& y) {
    return x.pair / h->pair_group < y.pair / h->pair_group;
  });

#define CREATE_TRY(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      fail(MDS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));                  \
      mds_rdf_destroy(h);                                                                  \
      return MDS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

  const size_t np = std::max<size_t>(1, src_index.size());
  CREATE_TRY(cudaMalloc(&h->d_src_index, np * sizeof(int32_t)));
  CREATE_TRY(cudaMalloc(&h->d_species, np * sizeof(int32_t)));
  if (!src_index.empty()) {
    CREATE_TRY(cudaMemcpy(h->d_src_index, src_index.data(), src_index.size() * sizeof(int32_t),
                          cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->d_species, species.data(), species.size() * sizeof(int32_t),
                          cudaMemcpyHostToDevice));
  }
  CREATE_TRY(cudaMalloc(&h->d_edges, h->edges.size() * sizeof(float)));
  CREATE_TRY(cudaMemcpy(h->d_edges, h->edges.data(), h->edges.size() * sizeof(float),
                        cudaMemcpyHostToDevice));
  const size_t n_counts = (size_t)h->n_pairs * (size_t)h->nbins;
  CREATE_TRY(cudaMalloc(&h->d_counts, n_counts * sizeof(unsigned long long)));
  CREATE_TRY(cudaMemset(h->d_counts, 0, n_counts * sizeof(unsigned long long)));
  CREATE_TRY(cudaMalloc(&h->d_evaluated, 2 * sizeof(unsigned long long)));
  CREATE_TRY(cudaMemset(h->d_evaluated, 0, 2 * sizeof(unsigned long long)));
  CREATE_TRY(cudaMalloc(&h->d_general, 2 * sizeof(unsigned long long)));
  CREATE_TRY(cudaMemset(h->d_general, 0, 2 * sizeof(unsigned long long)));
  if (cells) {
    // (row species, column species, x cells per staged segment) of every pair.  For a != b either
    // orientation counts the same pairs (full shell); the more populous species takes the rows,
    // so the less populous one is what a CTA stages 25-fold.  The segment length is the largest
    // whose staged atoms — neighbour pencils incl. ghost cells + the row pencil — fill 3/4 of the
    // CTA's capacity at average density; denser places are split by the kernel itself.
    const int K0 = h->kcell[0], wy = 2 * h->kcell[1] + 1, wz = 2 * h->kcell[2] + 1;
    std::vector<int4> ab;
    for (int a = 0; a < h->n_species; ++a)
      for (int b = a; b < h->n_species; ++b) {
        const int row = (h->pcnt[b] > h->pcnt[a]) ? b : a, col = (row == a) ? b : a;
        const double rho_r = (double)h->pcnt[row] / (double)h->grid.nc;
        const double rho_c = (double)h->pcnt[col] / (double)h->grid.nc;
        const int P = (a == b) ? h->kcell[2] * wy + h->kcell[1] : wy * wz;
        const int seg_max = std::min(h->grid.nx, mds::cells_max_seg());
        int seg = 1;
        for (int sg = 1; sg <= seg_max; ++sg) {
          const double staged = P * (sg + 2 * K0) * rho_c + (sg + K0) * rho_r;
          if (staged <= 0.75 * h->cap_atoms && (sg + 2 * K0) * P <= mds::cells_max_entries()) seg = sg;
        }
        const int pieces = (h->grid.nx + seg - 1) / seg;  // even pieces
        seg = (h->grid.nx + pieces - 1) / pieces;
        ab.push_back(make_int4(row, col, seg, 0));
        h->pair_pieces.push_back(pieces);
        h->cell_segments = std::max(h->cell_segments, pieces);
      }
    CREATE_TRY(cudaMalloc(&h->d_pair_ab, ab.size() * sizeof(int4)));
    CREATE_TRY(cudaMemcpy(h->d_pair_ab, ab.data(), ab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&h->d_work64, sizeof(unsigned long long) * WORK_RING));
    CREATE_TRY(cudaMemset(h->d_work64, 0, sizeof(unsigned long long) * WORK_RING));
    if (bf <= 4096) {
      const size_t words = mds::cells_scan_tmp_words(bf, h->n_species * h->grid.nc);
      CREATE_TRY(cudaMalloc(&h->d_scan_tmp, words * sizeof(uint32_t)));
      CREATE_TRY(cudaMemset(h->d_scan_tmp, 0, words * sizeof(uint32_t)));
    }
    CREATE_TRY(cudaMalloc(&h->d_batch_far, 2 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(h->d_batch_far, 0, 2 * sizeof(unsigned long long)));
    for (int b = 0; b < 2; ++b) {
      CREATE_TRY(cudaMalloc(&h->d_sorted[b], (size_t)h->frame_stride * 16 * (size_t)bf));
      CREATE_TRY(cudaMalloc(&h->d_rank[b], (size_t)h->frame_stride * 4 * (size_t)bf));
      CREATE_TRY(cudaMalloc(&h->d_cell_off[b], (size_t)h->off_stride * 4 * (size_t)bf));
    }
  }
  CREATE_TRY(cudaMalloc(&h->d_work, sizeof(unsigned int) * WORK_RING));
  CREATE_TRY(cudaMemset(h->d_work, 0, sizeof(unsigned int) * WORK_RING));
  for (int b = 0; b < 2; ++b) {
    CREATE_TRY(cudaMalloc(&h->d_packed[b], (size_t)h->frame_stride * 12 * (size_t)bf));
    CREATE_TRY(cudaMalloc(&h->d_flags[b], sizeof(uint32_t) * (size_t)bf));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_raw_ready[b], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_raw_free[b], cudaEventDisableTiming));
  }
  CREATE_TRY(cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreate(&h->ev_t0));
  CREATE_TRY(cudaEventCreate(&h->ev_t1));
  CREATE_TRY(cudaEventCreate(&h->ev_k0));
  CREATE_TRY(cudaEventCreate(&h->ev_k1));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
  {
    int rc = upload_items(h);
    if (rc != MDS_OK) { mds_rdf_destroy(h); return rc; }
  }
#undef CREATE_TRY
  *out = h;
  return MDS_OK;
}

void mds_rdf_destroy(mds_rdf_t* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  if (h->s_compute) cudaStreamSynchronize(h->s_compute);
  if (h->s_copy) cudaStreamSynchronize(h->s_copy);
  cudaFree(h->d_src_index);
  cudaFree(h->d_species);
  cudaFree(h->d_edges);
  cudaFree(h->d_items);
  cudaFree(h->d_counts);
  cudaFree(h->d_general);
  cudaFree(h->d_work);
  cudaFree(h->d_pair_ab);
  cudaFree(h->d_work64);
  cudaFree(h->d_scan_tmp);
  cudaFree(h->d_evaluated);
  cudaFree(h->d_batch_far);
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_packed[b]);
    cudaFree(h->d_flags[b]);
    cudaFree(h->d_raw[b]);
    cudaFree(h->d_sorted[b]);
    cudaFree(h->d_rank[b]);
    cudaFree(h->d_cell_off[b]);
    if (h->ev_raw_ready[b]) cudaEventDestroy(h->ev_raw_ready[b]);
    if (h->ev_raw_free[b]) cudaEventDestroy(h->ev_raw_free[b]);
  }
  for (int k = 0; k < 8; ++k)
    if (h->ev_ticket[k]) cudaEventDestroy(h->ev_ticket[k]);
  for (int k = 0; k < 3; ++k) {
    if (h->stage[k]) cudaFreeHost(h->stage[k]);
    if (h->ev_stage_done[k]) cudaEventDestroy(h->ev_stage_done[k]);
  }
  if (h->ev_ext) cudaEventDestroy(h->ev_ext);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (cudaEvent_t e : h->pk_events) cudaEventDestroy(e);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->ev_k0) cudaEventDestroy(h->ev_k0);
  if (h->ev_k1) cudaEventDestroy(h->ev_k1);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  delete h;
}

static int flush_stage(mds_rdf_t* h);

static int check_submit_args(mds_rdf_t* h, const float* xyz, int64_t n_frames, int layout,
                             const char* who) {
  if (!h) return fail(MDS_ERR_INVALID, "%s: handle is NULL", who);
  if (n_frames < 0) return fail(MDS_ERR_INVALID, "%s: n_frames < 0", who);
  if (n_frames > 0 && h->n_atoms > 0 && !xyz) return fail(MDS_ERR_INVALID, "%s: xyz is NULL", who);
  if (layout != MDS_LAYOUT_FRAME_MAJOR_XYZ && layout != MDS_LAYOUT_ATOM_MAJOR_XYZ)
    return fail(MDS_ERR_INVALID, "%s: unknown layout %d", who, layout);
  return MDS_OK;
}

int mds_rdf_submit_device(mds_rdf_t* h, const float* d_xyz, int64_t n_frames, int layout,
                          void* stream) {
  int rc = check_submit_args(h, d_xyz, n_frames, layout, "mds_rdf_submit_device");
  if (rc != MDS_OK) return rc;
  if (n_frames == 0) return MDS_OK;
  NvtxRange range("mds_rdf_submit_device");
  DeviceGuard guard(h->device);
  rc = flush_stage(h);  // frames submitted earlier go first
  if (rc != MDS_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_ext, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(h->s_compute, h->ev_ext, 0));
  CUDA_TRY(cudaEventRecord(h->ev_t0, h->s_compute));
  CUDA_TRY(cudaEventRecord(h->ev_k0, h->s_compute));
  for (int64_t f0 = 0; f0 < n_frames; f0 += h->batch_frames) {
    const int64_t nf = std::min<int64_t>(h->batch_frames, n_frames - f0);
    rc = process_batch(h, d_xyz, layout, f0, nf, n_frames, h->buf);
    if (rc != MDS_OK) return rc;
    h->buf ^= 1;
  }
  CUDA_TRY(cudaEventRecord(h->ev_k1, h->s_compute));
  CUDA_TRY(cudaEventRecord(h->ev_t1, h->s_compute));
  h->timing_valid = true;
  h->frames_done += n_frames;
  return MDS_OK;
}

// Batch plan of a host submit.  If the compute stream is idle the first host->device copy cannot
// overlap any compute, so the batches start small (1/16 of a full one) and grow no faster than
// the kernels of one batch can hide the copy of the next (cell path: x 1.5 — its compute time per
// frame is close to the copy time; all-pairs path: x 4).  If earlier work is still running (the
// caller streams a trajectory block by block) the copy of the first batch is hidden behind it and
// a ramp would only make the full-size copy that follows wait on a small batch's kernels.  The
// rest of the call is split into equal batches, for the same reason: a short last batch would
// expose most of the next call's first copy.
static std::vector<int64_t> plan_host_batches(int64_t n_frames, int64_t batch_frames, bool idle,
                                              bool cells) {
  std::vector<int64_t> plan;
  int64_t rem = n_frames;
  if (idle && n_frames >= 8) {
    for (int64_t cur = std::max<int64_t>(1, batch_frames / 16); rem > 0 && cur < batch_frames;
         cur = cells ? cur + (cur + 1) / 2 : cur * 4) {
      const int64_t nf = std::min(cur, rem);
      plan.push_back(nf);
      rem -= nf;
    }
  }
  if (rem > 0) {
    const int64_t nb = (rem + batch_frames - 1) / batch_frames;
    const int64_t per = (rem + nb - 1) / nb;
    for (; rem > 0; rem -= std::min(per, rem)) plan.push_back(std::min(per, rem));
  }
  return plan;
}

int64_t mds_rdf_plan_host_batches(int64_t n_frames, int32_t batch_frames, int compute_idle,
                                  int cell_path, int64_t* out, int64_t capacity) {
  if (n_frames < 0 || batch_frames < 1 || capacity < 0 || (capacity > 0 && !out)) return -1;
  const std::vector<int64_t> plan =
      plan_host_batches(n_frames, batch_frames, compute_idle != 0, cell_path != 0);
  for (int64_t k = 0; k < (int64_t)plan.size() && k < capacity; ++k) out[k] = plan[k];
  return (int64_t)plan.size();
}

// Host frames in page-locked memory: H2D on the side stream in batches, double buffered against
// the kernels.
static int submit_pinned(mds_rdf_t* h, const float* xyz, int64_t n_frames, int layout) {
  NvtxRange range("mds_rdf_submit (host frames)");
  const size_t frame_bytes = (size_t)h->n_atoms * 3 * sizeof(float);
  const size_t need = std::max<size_t>(16, frame_bytes * (size_t)h->batch_frames);
  if (h->raw_bytes < need) {
    CUDA_TRY(cudaStreamSynchronize(h->s_compute));
    CUDA_TRY(cudaStreamSynchronize(h->s_copy));
    for (int b = 0; b < 2; ++b) {
      cudaFree(h->d_raw[b]);
      h->d_raw[b] = nullptr;
      CUDA_TRY(cudaMalloc(&h->d_raw[b], need));
    }
    h->raw_bytes = need;
  }
  const bool idle = cudaStreamQuery(h->s_compute) == cudaSuccess;
  (void)cudaGetLastError();  // (cudaErrorNotReady is an answer, not a failure)
  CUDA_TRY(cudaEventRecord(h->ev_t0, h->s_compute));
  bool first_kernel = true;
  // batch plan (plan_host_batches above): ramp only if nothing is running, equal batches after it
  // (an atom-major block that fits one batch stays whole: it travels as ONE plain copy, while
  // pieces of it would need strided copies with rows of a few hundred bytes)
  const bool ramp = idle && (layout == MDS_LAYOUT_FRAME_MAJOR_XYZ || n_frames > h->batch_frames);
  const std::vector<int64_t> plan =
      plan_host_batches(n_frames, h->batch_frames, ramp, h->path == MDS_RDF_PATH_CELLLIST);
  int64_t f0 = 0;
  for (const int64_t nf : plan) {
    const int b = h->buf;
    // the staging buffer must have been consumed by the pack kernel of two batches ago
    CUDA_TRY(cudaStreamWaitEvent(h->s_copy, h->ev_raw_free[b], 0));
    if (layout == MDS_LAYOUT_FRAME_MAJOR_XYZ) {
      CUDA_TRY(cudaMemcpyAsync(h->d_raw[b], xyz + (size_t)f0 * h->n_atoms * 3, frame_bytes * nf,
                               cudaMemcpyHostToDevice, h->s_copy));
    } else if (nf == n_frames) {
      // atom-major and the whole call in one batch (the reference's (N, C, 3) positions_tensor
      // with a modest C): the block is contiguous, one plain copy
      CUDA_TRY(cudaMemcpyAsync(h->d_raw[b], xyz, frame_bytes * nf, cudaMemcpyHostToDevice, h->s_copy));
    } else {
      // atom-major: one strided 2-D copy gathers frames [f0, f0+nf) of every atom
      CUDA_TRY(cudaMemcpy2DAsync(h->d_raw[b], (size_t)nf * 12, xyz + (size_t)f0 * 3,
                                 (size_t)n_frames * 12, (size_t)nf * 12, (size_t)h->n_atoms,
                                 cudaMemcpyHostToDevice, h->s_copy));
    }
    CUDA_TRY(cudaEventRecord(h->ev_raw_ready[b], h->s_copy));
    CUDA_TRY(cudaStreamWaitEvent(h->s_compute, h->ev_raw_ready[b], 0));
    if (first_kernel) {
      CUDA_TRY(cudaEventRecord(h->ev_k0, h->s_compute));
      first_kernel = false;
    }
    // inside the staging buffer the batch starts at frame 0 and holds nf frames
    const int rc = process_batch(h, h->d_raw[b], layout, 0, nf, nf, b);
    if (rc != MDS_OK) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_raw_free[b], h->s_compute));
    h->buf ^= 1;
    f0 += nf;
  }
  CUDA_TRY(cudaEventRecord(h->ev_k1, h->s_compute));
  CUDA_TRY(cudaEventRecord(h->ev_t1, h->s_compute));
  {
    cudaEvent_t& ev = h->ev_ticket[h->tickets % 8];
    if (!ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ev, h->s_copy));  // every H2D of this submit has been enqueued
    h->tickets += 1;
  }
  h->timing_valid = true;
  h->frames_done += n_frames;
  return MDS_OK;
}

// Send the frames waiting in the current pinned ring buffer (pageable submits) on their way.
static int flush_stage(mds_rdf_t* h) {
  if (h->stage_frames == 0) return MDS_OK;
  const int slot = h->stage_slot;
  const int64_t nf = h->stage_frames;
  h->stage_frames = 0;
  h->stage_slot = (slot + 1) % 3;
  const int rc = submit_pinned(h, h->stage[slot], nf, MDS_LAYOUT_FRAME_MAJOR_XYZ);
  if (rc != MDS_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_stage_done[slot], h->s_copy));  // every H2D out of the slot is queued
  return MDS_OK;
}

// Host frames in PAGEABLE memory (the reference-side binding passes plain NumPy arrays, one batch
// of C configurations per call).  A cudaMemcpyAsync from pageable memory is
// staged by the driver call by call and an atom-major batch would go as N rows of C x 12 bytes;
// instead the calling thread copies the frames into a ring of pinned buffers — transposing the
// (N, C, 3) layout to frame-major on the way — and a buffer is sent when it is full or when the
// caller asks for results.  The caller's array may be reused as soon as the call returns.
static int submit_pageable(mds_rdf_t* h, const float* xyz, int64_t n_frames, int layout) {
  NvtxRange range("mds_rdf_submit (pageable host frames)");
  const size_t frame_floats = (size_t)h->n_atoms * 3;
  if (frame_floats == 0) { h->frames_done += n_frames; return MDS_OK; }
  if (!h->stage[0]) {
    const size_t frame_bytes = frame_floats * sizeof(float);
    h->stage_capacity = std::max<int64_t>(1, std::min<int64_t>(h->batch_frames,
                                                               (int64_t)((16u << 20) / frame_bytes)));
    for (int k = 0; k < 3; ++k) {
      CUDA_TRY(cudaHostAlloc((void**)&h->stage[k], frame_bytes * (size_t)h->stage_capacity,
                             cudaHostAllocDefault));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_stage_done[k], cudaEventDisableTiming));
    }
  }
  if (layout == MDS_LAYOUT_ATOM_MAJOR_XYZ && n_frames == 1) layout = MDS_LAYOUT_FRAME_MAJOR_XYZ;  // same bytes
  if (layout == MDS_LAYOUT_ATOM_MAJOR_XYZ && n_frames <= h->stage_capacity) {
    // a (N, C, 3) block that fits one buffer: copy it as it is — the pack kernel reads the
    // atom-major layout directly — and send it as a batch of its own
    int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
    const int slot = h->stage_slot;
    CUDA_TRY(cudaEventSynchronize(h->ev_stage_done[slot]));
    memcpy(h->stage[slot], xyz, (size_t)n_frames * frame_floats * sizeof(float));
    h->stage_slot = (slot + 1) % 3;
    rc = submit_pinned(h, h->stage[slot], n_frames, MDS_LAYOUT_ATOM_MAJOR_XYZ);
    if (rc != MDS_OK) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_stage_done[slot], h->s_copy));
    return MDS_OK;
  }
  for (int64_t f0 = 0; f0 < n_frames;) {
    if (h->stage_frames == 0)  // first write into this buffer: its previous copies must be done
      CUDA_TRY(cudaEventSynchronize(h->ev_stage_done[h->stage_slot]));
    const int64_t nf = std::min<int64_t>(h->stage_capacity - h->stage_frames, n_frames - f0);
    float* dst = h->stage[h->stage_slot] + (size_t)h->stage_frames * frame_floats;
    if (layout == MDS_LAYOUT_FRAME_MAJOR_XYZ) {
      memcpy(dst, xyz + (size_t)f0 * frame_floats, (size_t)nf * frame_floats * sizeof(float));
    } else {
      // (N, C, 3) with more configurations than a buffer holds -> (nf, N, 3), in blocks of atoms so
      // that both sides stay in cache
      const int64_t N = h->n_atoms;
      for (int64_t a0 = 0; a0 < N; a0 += 256) {
        const int64_t a1 = std::min<int64_t>(N, a0 + 256);
        for (int64_t f = 0; f < nf; ++f) {
          float* d = dst + ((size_t)f * N + a0) * 3;
          const float* sp = xyz + ((size_t)a0 * n_frames + f0 + f) * 3;
          for (int64_t a = a0; a < a1; ++a, d += 3, sp += (size_t)n_frames * 3) {
            d[0] = sp[0]; d[1] = sp[1]; d[2] = sp[2];
          }
        }
      }
    }
    h->stage_frames += nf;
    f0 += nf;
    if (h->stage_frames == h->stage_capacity) {
      const int rc = flush_stage(h);
      if (rc != MDS_OK) return rc;
    }
  }
  return MDS_OK;
}

int mds_rdf_submit(mds_rdf_t* h, const float* xyz, int64_t n_frames, int layout,
                   int host_is_pinned) {
  int rc = check_submit_args(h, xyz, n_frames, layout, "mds_rdf_submit");
  if (rc != MDS_OK) return rc;
  if (n_frames == 0) return MDS_OK;
  DeviceGuard guard(h->device);
  if (!host_is_pinned) return submit_pageable(h, xyz, n_frames, layout);
  rc = flush_stage(h);  // frames submitted earlier go first
  if (rc != MDS_OK) return rc;
  return submit_pinned(h, xyz, n_frames, layout);
}

int mds_rdf_join(mds_rdf_t* h, void* stream) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_join: handle is NULL");
  DeviceGuard guard(h->device);
  {
    const int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
  }
  CUDA_TRY(cudaEventRecord(h->ev_join, h->s_compute));
  CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_join, 0));
  return MDS_OK;
}

int mds_rdf_wait(mds_rdf_t* h, void* stream) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_wait: handle is NULL");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaEventRecord(h->ev_ext, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(h->s_compute, h->ev_ext, 0));
  return MDS_OK;
}

// ncclAllReduce, resolved at run time: the library has no link-time dependency on NCCL, a caller
// that shards over GPUs already has it loaded (or libnccl.so.2 is on the loader path).
namespace {
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
nccl_allreduce_fn g_nccl_allreduce = nullptr;
nccl_errstr_fn g_nccl_errstr = nullptr;
bool resolve_nccl() {
  if (g_nccl_allreduce) return true;
  void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
  void* err = dlsym(RTLD_DEFAULT, "ncclGetErrorString");
  if (!sym) {
    if (void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL)) {
      sym = dlsym(lib, "ncclAllReduce");
      err = dlsym(lib, "ncclGetErrorString");
    }
  }
  g_nccl_allreduce = reinterpret_cast<nccl_allreduce_fn>(sym);
  g_nccl_errstr = reinterpret_cast<nccl_errstr_fn>(err);
  return g_nccl_allreduce != nullptr;
}
}  // namespace

int mds_rdf_allreduce(mds_rdf_t* h, struct ncclComm* comm) {
  if (!h || !comm) return fail(MDS_ERR_INVALID, "mds_rdf_allreduce: NULL argument");
  if (!resolve_nccl())
    return fail(MDS_ERR_UNSUPPORTED, "mds_rdf_allreduce: ncclAllReduce not found (load NCCL first or "
                                     "put libnccl.so.2 on the loader path)");
  DeviceGuard guard(h->device);
  {
    const int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
  }
  NvtxRange range("mds_rdf_allreduce");
  const size_t n = (size_t)h->n_pairs * (size_t)h->nbins;
  // in place, on the library's own stream: ordered after every kernel submitted so far and before
  // anything submitted or reset afterwards — no join / wait needed around it
  const int rc = g_nccl_allreduce(h->d_counts, h->d_counts, n, /*ncclUint64*/ 5, /*ncclSum*/ 0, comm,
                                  h->s_compute);
  if (rc != 0)
    return fail(MDS_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl_errstr ? g_nccl_errstr(rc) : "?");
  return MDS_OK;
}

int mds_rdf_wait_host_copies(mds_rdf_t* h) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_wait_host_copies: handle is NULL");
  DeviceGuard guard(h->device);
  {
    const int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  return MDS_OK;
}

int64_t mds_rdf_host_ticket(mds_rdf_t* h) { return h ? h->tickets : -1; }

int mds_rdf_wait_host_ticket(mds_rdf_t* h, int64_t ticket) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_wait_host_ticket: handle is NULL");
  DeviceGuard guard(h->device);
  if (ticket < 1 || ticket > h->tickets)
    return fail(MDS_ERR_INVALID, "mds_rdf_wait_host_ticket: ticket %lld not in [1, %lld]",
                (long long)ticket, (long long)h->tickets);
  if (h->tickets - ticket >= 8) {  // older than the event ring: wait for the stream
    CUDA_TRY(cudaStreamSynchronize(h->s_copy));
    return MDS_OK;
  }
  CUDA_TRY(cudaEventSynchronize(h->ev_ticket[(ticket - 1) % 8]));
  return MDS_OK;
}

int mds_rdf_synchronize(mds_rdf_t* h) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_synchronize: handle is NULL");
  DeviceGuard guard(h->device);
  {
    const int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  CUDA_TRY(cudaStreamSynchronize(h->s_compute));
  return MDS_OK;
}

int mds_rdf_counts(mds_rdf_t* h, uint64_t* out, int sync) {
  if (!h || !out) return fail(MDS_ERR_INVALID, "mds_rdf_counts: NULL argument");
  DeviceGuard guard(h->device);
  {
    const int rc = flush_stage(h);
    if (rc != MDS_OK) return rc;
  }
  const size_t bytes = (size_t)h->n_pairs * (size_t)h->nbins * sizeof(uint64_t);
  CUDA_TRY(cudaMemcpyAsync(out, h->d_counts, bytes, cudaMemcpyDeviceToHost, h->s_compute));
  if (sync) return mds_rdf_synchronize(h);
  return MDS_OK;
}

int mds_rdf_device_counts(mds_rdf_t* h, void** d_counts, void** stream) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_device_counts: handle is NULL");
  if (d_counts) *d_counts = h->d_counts;
  if (stream) *stream = (void*)h->s_compute;
  return MDS_OK;
}

int mds_rdf_reset(mds_rdf_t* h) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_reset: handle is NULL");
  DeviceGuard guard(h->device);
  h->stage_frames = 0;  // frames still waiting in the pinned ring belong to the run being discarded
  const size_t bytes = (size_t)h->n_pairs * (size_t)h->nbins * sizeof(uint64_t);
  CUDA_TRY(cudaMemsetAsync(h->d_counts, 0, bytes, h->s_compute));
  CUDA_TRY(cudaMemsetAsync(h->d_general, 0, 2 * sizeof(unsigned long long), h->s_compute));
  if (h->d_evaluated)
    CUDA_TRY(cudaMemsetAsync(h->d_evaluated, 0, 2 * sizeof(unsigned long long), h->s_compute));
  h->frames_done = 0;
  h->launches = 0;
  h->timing_valid = false;
  return MDS_OK;
}

int mds_rdf_set_item_shard(mds_rdf_t* h, int32_t index, int32_t count) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_set_item_shard: handle is NULL");
  if (count < 1 || index < 0 || index >= count)
    return fail(MDS_ERR_INVALID, "mds_rdf_set_item_shard: shard %d of %d", index, count);
  h->shard_index = (uint32_t)index;
  h->shard_count = (uint32_t)count;
  return MDS_OK;
}

int mds_rdf_stats_get(mds_rdf_t* h, mds_rdf_stats* out) {
  if (!h || !out) return fail(MDS_ERR_INVALID, "mds_rdf_stats_get: NULL argument");
  if (out->struct_size != sizeof(mds_rdf_stats))
    return fail(MDS_ERR_INVALID, "mds_rdf_stats_get: struct_size %u != %zu", out->struct_size,
                sizeof(mds_rdf_stats));
  DeviceGuard guard(h->device);
  int rc = mds_rdf_synchronize(h);
  if (rc != MDS_OK) return rc;
  const size_t n = (size_t)h->n_pairs * (size_t)h->nbins;
  std::vector<unsigned long long> host(n);
  CUDA_TRY(cudaMemcpy(host.data(), h->d_counts, n * sizeof(unsigned long long),
                      cudaMemcpyDeviceToHost));
  double binned = 0;
  for (auto v : host) binned += (double)v;
  unsigned long long general[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(general, h->d_general, sizeof(general), cudaMemcpyDeviceToHost));
  out->path = h->path;
  out->frames = h->frames_done;
  out->pairs_allpairs_equiv = h->pairs_per_frame * (double)h->frames_done;
  {
    // counted on the device: reference pairs of every (item, frame) unit the all-pairs kernel
    // processed (all of them, or this handle's item shard, or the far frames handed over by the
    // cell path) + the candidate pairs the cell kernel tested
    unsigned long long ev[2] = {0, 0};
    CUDA_TRY(cudaMemcpy(ev, h->d_evaluated, sizeof(ev), cudaMemcpyDeviceToHost));
    out->pairs_evaluated = (double)ev[0];
    // the cell kernel tests whole 64-column chunks against contiguous row ranges: a superset of
    // the candidate pairs (rdf_cells.cu)
    out->pairs_tested = h->path == MDS_RDF_PATH_CELLLIST ? (double)ev[1] : (double)ev[0];
  }
  out->pairs_allpairs_equiv /= (double)h->shard_count;
  out->pairs_binned = binned;
  out->frames_general_minimage = (int64_t)general[0];
  out->kernel_launches = h->launches;
  out->kernel_ms = 0.f;
  out->submit_ms = 0.f;
  if (h->timing_valid) {
    cudaEventElapsedTime(&out->kernel_ms, h->ev_k0, h->ev_k1);
    cudaEventElapsedTime(&out->submit_ms, h->ev_t0, h->ev_t1);
  }
  out->n_pairs = h->n_pairs;
  out->nbins = h->nbins;
  out->sm_count = h->sm_count;
  out->variant = h->variant;
  out->pair_kernel_ms = 0.f;
  out->pair_kernel_launches = h->pk_used;
  for (int i = 0; i < h->pk_used; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->pk_events[2 * i], h->pk_events[2 * i + 1]);
    out->pair_kernel_ms += ms;
  }
  h->pk_used = 0;  // "since the previous mds_rdf_stats_get"; survives submit and reset
  out->dense = h->dense ? 1 : 0;
  out->frames_far = (int32_t)general[1];
  out->cells[0] = h->grid.nx;
  out->cells[1] = h->grid.ny;
  out->cells[2] = h->grid.nz;
  for (int k = 0; k < 3; ++k) out->cell_stencil[k] = h->kcell[k];
  out->cell_build = h->built_fused;
  out->item_shards = (int32_t)h->shard_count;
  return MDS_OK;
}

int mds_rdf_threshold_table(mds_rdf_t* h, float* edges, float* s_cut) {
  if (!h) return fail(MDS_ERR_INVALID, "mds_rdf_threshold_table: handle is NULL");
  if (edges) {
    edges[0] = 0.0f;
    for (int k = 1; k <= h->nbins; ++k) edges[k] = h->edges[k];
  }
  if (s_cut) *s_cut = h->s_cut;
  return MDS_OK;
}

int mds_rdf_build_threshold_table(float cutoff, int nbins, float* edges, float* s_cut) {
  if (nbins < 1 || nbins > mds::MDS_MAX_BINS || !(cutoff > 0.f) || !std::isfinite(cutoff))
    return fail(MDS_ERR_INVALID, "mds_rdf_build_threshold_table: bad cutoff / nbins");
  std::vector<float> e;
  float sc = 0.f;
  build_threshold_table(cutoff, nbins, e, sc);
  if (edges) {
    edges[0] = 0.0f;
    for (int k = 1; k <= nbins; ++k) edges[k] = e[k];
  }
  if (s_cut) *s_cut = sc;
  return MDS_OK;
}

int mds_bench_shared_atomics(int device, int nbins, int iters, double* spread, double* random,
                             double* sm_clock_mhz) {
  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || device < 0 || device >= n_dev)
    return fail(MDS_ERR_CUDA, "mds_bench_shared_atomics: no such CUDA device %d", device);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  float ms0 = 0, ms1 = 0;
  double l0 = 0, l1 = 0;
  CUDA_TRY(mds::run_shared_atomic_bench(nbins, iters, 0, prop.multiProcessorCount, &ms0, &l0));
  CUDA_TRY(mds::run_shared_atomic_bench(nbins, iters, 1, prop.multiProcessorCount, &ms1, &l1));
  if (spread) *spread = l0;
  if (random) *random = l1;
  if (sm_clock_mhz) {
    // cycles per block / elapsed time of the same launch
    const double cycles = 1024.0 * (double)iters / l1;
    *sm_clock_mhz = cycles / ((double)ms1 * 1e3);
  }
  return MDS_OK;
}

}  // extern "C"
