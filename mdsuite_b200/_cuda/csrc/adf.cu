// Angular distribution function for sm_100a behind include/mds_adf.h: neighbour search with the
// reference's float16 cutoff, then every (i; j, k) triple of a centre atom binned by its angle.
//
// Arithmetic contract (oracle/adf_oracle.py; reference utils/neighbour_list.py:63-78,104-177,
// utils/linalg.py:31-81), float32 with every operation rounded separately:
//   r = pos[i] - pos[j];  r -= rint(r / L) * L;  s = (x*x + y*y) + z*z;  n = sqrt(s)
//   neighbour iff float16(n) != 0 and float16(n) < float16(r_cut)   <=>  s_lo <= s < s_hi
//   u = r / n (per component);  c = clip((ux1*ux2 + uy1*uy2) + uz1*uz2, -1, 1)
//   theta = float32(acos(c)), bin by np.histogram over [0, 3.15]    <=>  c against cos_edges[]
//   weight = 1 / (n1 * n2)^norm_power
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

constexpr int ADF_WARPS = 8;
constexpr int ADF_THREADS = ADF_WARPS * 32;
constexpr int MAX_NB = MDS_ADF_MAX_NEIGHBOURS;
constexpr int CENTRES_PER_ITEM = 128;
constexpr int MAX_SPECIES = 8;

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
  float lo = 5.9604645e-8f;  // 2^-24, the smallest float16 subnormal
  while (lo > 0.0f && half_round(nextafterf(lo, -INFINITY)) != 0.0f) lo = nextafterf(lo, -INFINITY);
  *n_lo = lo;
  *n_hi = hi;
  return MDS_ADF_OK;
}

}  // namespace

// ---- device -----------------------------------------------------------------------------------

struct AdfParams {
  const float* pos;         // [F][N][3]
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

__global__ void __launch_bounds__(ADF_THREADS, 2) adf_kernel(const AdfParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* nb_vec = reinterpret_cast<float4*>(smem_raw);                    // [WARPS][MAX_NB]
  uint8_t* nb_sp = reinterpret_cast<uint8_t*>(nb_vec + ADF_WARPS * MAX_NB);  // [WARPS][MAX_NB]
  float* hist = reinterpret_cast<float*>(nb_sp + ADF_WARPS * MAX_NB);       // [combos][nbins]
  float* edges = hist + p.n_combos * p.nbins;                               // [nbins + 1]
  __shared__ unsigned long long s_item;
  __shared__ int8_t s_combo[MAX_SPECIES * MAX_SPECIES * MAX_SPECIES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hist_len = p.n_combos * p.nbins;
  for (int k = tid; k < hist_len; k += ADF_THREADS) hist[k] = 0.0f;
  for (int k = tid; k <= p.nbins; k += ADF_THREADS) edges[k] = p.cos_edges[k];
  for (int k = tid; k < MAX_SPECIES * MAX_SPECIES * MAX_SPECIES; k += ADF_THREADS)
    s_combo[k] = p.combo_of[k];
  float4* my_vec = nb_vec + warp * MAX_NB;
  uint8_t* my_sp = nb_sp + warp * MAX_NB;
  const unsigned long long n_items = (unsigned long long)p.n_frames * p.items_per_frame;
  unsigned long long my_triples = 0;
  int my_max_nb = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(p.work_counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int frame = (int)(item / p.items_per_frame);
    const int c0 = (int)(item % p.items_per_frame) * CENTRES_PER_ITEM;
    const int c1 = min(p.n_atoms, c0 + CENTRES_PER_ITEM);
    const float* fpos = p.pos + (size_t)frame * p.n_atoms * 3;

    for (int i = c0 + warp; i < c1; i += ADF_WARPS) {
      const float xi = __ldg(fpos + 3 * i), yi = __ldg(fpos + 3 * i + 1), zi = __ldg(fpos + 3 * i + 2);
      const int sp_i = p.species[i];
      // ---- neighbour search: every atom of the configuration against centre i ----
      int m = 0;
      for (int j0 = 0; j0 < p.n_atoms; j0 += 32) {
        const int j = j0 + lane;
        bool is_nb = false;
        float dx = 0.f, dy = 0.f, dz = 0.f, s = 0.f;
        int sp_j = 0;
        if (j < p.n_atoms) {
          dx = min_image(__fsub_rn(xi, __ldg(fpos + 3 * j)), p.box[0]);
          dy = min_image(__fsub_rn(yi, __ldg(fpos + 3 * j + 1)), p.box[1]);
          dz = min_image(__fsub_rn(zi, __ldg(fpos + 3 * j + 2)), p.box[2]);
          s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          sp_j = p.species[j];
          is_nb = (s >= p.s_lo) && (s < p.s_hi) && (sp_j >= sp_i);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, is_nb);
        if (is_nb) {
          const int slot = m + __popc(ballot & ((1u << lane) - 1u));
          if (slot < MAX_NB) {
            const float n = __fsqrt_rn(s);
            my_vec[slot] = make_float4(__fdiv_rn(dx, n), __fdiv_rn(dy, n), __fdiv_rn(dz, n), n);
            my_sp[slot] = (uint8_t)sp_j;
          }
        }
        m += __popc(ballot);
      }
      if (m > MAX_NB) {
        if (lane == 0) atomicExch(p.status, 1);
        m = MAX_NB;
      }
      my_max_nb = max(my_max_nb, m);
      __syncwarp();
      // ---- triples: every unordered pair (j, k) of the neighbour list ----
      for (int a = 0; a + 1 < m; ++a) {
        const float4 va = my_vec[a];
        const int sa = my_sp[a];
        for (int b = a + 1 + lane; b < m; b += 32) {
          const float4 vb = my_vec[b];
          const int sb = my_sp[b];
          float c = __fadd_rn(__fadd_rn(__fmul_rn(va.x, vb.x), __fmul_rn(va.y, vb.y)),
                              __fmul_rn(va.z, vb.z));
          c = fminf(fmaxf(c, -1.0f), 1.0f);
          int g = (int)(acosf(c) * p.bin_scale);
          g = max(0, min(g, p.nbins - 1));
          while (g < p.nbins - 1 && c <= edges[g + 1]) ++g;
          while (g > 0 && c > edges[g]) --g;
          const int lo = min(sa, sb), hi = max(sa, sb);
          const int combo = s_combo[(sp_i * MAX_SPECIES + lo) * MAX_SPECIES + hi];
          float w = (sa == sb) ? 2.0f : 1.0f;
          my_triples += (sa == sb) ? 2ull : 1ull;
          if (p.norm_power > 0) {
            const float pr = __fmul_rn(va.w, vb.w);
            float pw = pr;
            for (int e = 1; e < p.norm_power; ++e) pw = __fmul_rn(pw, pr);
            w = __fmul_rn(w, __fdiv_rn(1.0f, pw));
          }
          atomicAdd(&hist[combo * p.nbins + g], w);
        }
      }
      __syncwarp();
    }
    // ---- flush this item's sums into the configuration's histogram ----
    __syncthreads();
    double* fout = p.out + (size_t)frame * hist_len;
    for (int k = tid; k < hist_len; k += ADF_THREADS) {
      const float v = hist[k];
      if (v != 0.0f) {
        atomicAdd(fout + k, (double)v);
        hist[k] = 0.0f;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_triples += __shfl_xor_sync(0xffffffffu, my_triples, o);
    my_max_nb = max(my_max_nb, __shfl_xor_sync(0xffffffffu, my_max_nb, o));
  }
  if (lane == 0) {
    if (my_triples) atomicAdd(p.triples, my_triples);
    atomicMax(p.status + 1, my_max_nb);
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
  // device
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
  float* d_pos = nullptr;
  double* d_out = nullptr;
  uint8_t* d_species = nullptr;
  float* d_edges = nullptr;
  unsigned long long* d_counters = nullptr;  // [0] work counter, [1] triples
  int* d_status = nullptr;
  // statistics
  int64_t frames = 0, launches = 0;
  double triples = 0, pair_tests = 0;
  float kernel_ms = 0, compute_ms = 0;
  int max_neighbours = 0;
};

namespace {

int run_batch(mds_adf* h, const float* d_xyz, int n_frames, double* d_out) {
  const size_t out_bytes = (size_t)n_frames * h->n_combos * h->nbins * sizeof(double);
  ADF_CUDA_TRY(cudaMemsetAsync(d_out, 0, out_bytes, h->stream));
  ADF_CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, 2 * sizeof(unsigned long long), h->stream));
  AdfParams p = h->base;
  p.pos = d_xyz;
  p.out = d_out;
  p.n_frames = n_frames;
  const unsigned long long items = (unsigned long long)n_frames * p.items_per_frame;
  const int grid = (int)std::min<unsigned long long>((unsigned long long)h->grid, items);
  adf_kernel<<<std::max(grid, 1), ADF_THREADS, h->smem_bytes, h->stream>>>(p);
  ADF_CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  unsigned long long counters[2];
  int status[2];
  ADF_CUDA_TRY(cudaMemcpyAsync(counters, h->d_counters, sizeof(counters), cudaMemcpyDeviceToHost,
                               h->stream));
  ADF_CUDA_TRY(cudaMemcpyAsync(status, h->d_status, sizeof(status), cudaMemcpyDeviceToHost,
                               h->stream));
  ADF_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->triples += (double)counters[1];
  h->pair_tests += (double)n_frames * (double)h->n_atoms * (double)h->n_atoms;
  h->max_neighbours = std::max(h->max_neighbours, status[1]);
  h->frames += n_frames;
  if (status[0]) {
    cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream);
    return fail(MDS_ADF_ERR_UNSUPPORTED,
                "an atom has more than %d neighbours within the cutoff %g; reduce the cutoff",
                MAX_NB, (double)h->cutoff);
  }
  return MDS_ADF_OK;
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
  h->smem_bytes = (size_t)ADF_WARPS * MAX_NB * (sizeof(float4) + 1) +
                  sizeof(float) * ((size_t)h->n_combos * h->nbins + h->nbins + 1);
  if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
    return cleanup(fail(MDS_ADF_ERR_UNSUPPORTED,
                        "%d combinations x %d bins need %zu bytes of shared memory (limit %zu)",
                        h->n_combos, h->nbins, h->smem_bytes, (size_t)prop.sharedMemPerBlockOptin));
  if (cudaFuncSetAttribute(adf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)h->smem_bytes) != cudaSuccess)
    return cleanup(fail(MDS_ADF_ERR_CUDA, "cudaFuncSetAttribute failed: %s",
                        cudaGetErrorString(cudaGetLastError())));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adf_kernel, ADF_THREADS,
                                                    h->smem_bytes) != cudaSuccess || per_sm < 1)
    return cleanup(fail(MDS_ADF_ERR_CUDA, "the ADF kernel does not fit on an SM"));
  h->grid = per_sm * h->sm_count;

  const size_t frame_bytes = (size_t)n_atoms * 3 * sizeof(float);
  const size_t out_frame_bytes = (size_t)h->n_combos * h->nbins * sizeof(double);
  int64_t bf = cfg->max_frames_in_flight > 0 ? cfg->max_frames_in_flight
                                             : (int64_t)((256u << 20) / (frame_bytes + out_frame_bytes));
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
  ADF_TRY_CREATE(cudaMalloc(&h->d_pos, frame_bytes * (size_t)bf));
  ADF_TRY_CREATE(cudaMalloc(&h->d_out, out_frame_bytes * (size_t)bf));
  ADF_TRY_CREATE(cudaMalloc(&h->d_species, (size_t)n_atoms));
  ADF_TRY_CREATE(cudaMalloc(&h->d_edges, sizeof(float) * ((size_t)h->nbins + 1)));
  ADF_TRY_CREATE(cudaMalloc(&h->d_counters, 2 * sizeof(unsigned long long)));
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
  p.species = h->d_species;
  p.cos_edges = h->d_edges;
  p.work_counter = h->d_counters;
  p.triples = h->d_counters + 1;
  p.status = h->d_status;
  p.n_atoms = (int)n_atoms;
  p.items_per_frame = (int)((n_atoms + CENTRES_PER_ITEM - 1) / CENTRES_PER_ITEM);
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
  if (n_frames > 0) {
    const int rc = run_batch(h, d_xyz, (int)n_frames, d_out);
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
  return MDS_ADF_OK;
}

}  // extern "C"
