// Micro-benchmark behind the shared-atomic roofline: how many shared-memory
// atomicAdd lanes per clock one SM retires, conflict-free (lanes on distinct banks) and with
// pseudo-random bins (what a distance histogram looks like).
#include "rdf_common.cuh"
#include "rdf_launch.h"

namespace mds {

// mode 0: lane-consecutive bins (no bank conflicts)   mode 1: LCG-random bins
template <int MODE>
__global__ void __launch_bounds__(1024, 1) shared_atomic_bench_kernel(int nbins, int iters,
                                                                      long long* cycles,
                                                                      unsigned int* sink) {
  extern __shared__ uint32_t hist[];
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) hist[k] = 0u;
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    uint32_t k;
    if (MODE == 0) {
      const uint32_t base = __shfl_sync(0xffffffffu, x, 0) >> 8;  // warp-uniform
      k = (__umulhi(base << 8, (uint32_t)(nbins - 32)) & ~31u) + lane;
    } else {
      k = __umulhi(x, (uint32_t)nbins);
    }
    atomicAdd(hist + k, 1u);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  unsigned int acc = 0;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) acc += hist[k];
  if (acc == 0xffffffffu) *sink = acc;
}

cudaError_t run_shared_atomic_bench(int nbins, int iters, int mode, int sm_count, float* ms,
                                    double* lanes) {
  if (nbins < 64) nbins = 64;
  long long* d_cycles = nullptr;
  unsigned int* d_sink = nullptr;
  cudaError_t e = cudaMalloc(&d_cycles, sizeof(long long) * sm_count);
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&d_sink, sizeof(unsigned int));
  if (e != cudaSuccess) { cudaFree(d_cycles); return e; }
  const size_t smem = (size_t)nbins * sizeof(uint32_t);
  auto kern = mode == 0 ? shared_atomic_bench_kernel<0> : shared_atomic_bench_kernel<1>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  if (e == cudaSuccess) {
    kern<<<sm_count, 1024, smem>>>(nbins, iters / 8 + 1, d_cycles, d_sink);  // warm-up
    cudaEventRecord(a);
    kern<<<sm_count, 1024, smem>>>(nbins, iters, d_cycles, d_sink);
    cudaEventRecord(b);
    e = cudaEventSynchronize(b);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) {
    cudaEventElapsedTime(ms, a, b);
    long long* h = new long long[sm_count];
    cudaMemcpy(h, d_cycles, sizeof(long long) * sm_count, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < sm_count; ++i) sum += (double)h[i];
    delete[] h;
    const double mean_cycles = sum / sm_count;
    *lanes = 1024.0 * (double)iters / mean_cycles;  // lanes per clock per SM
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_cycles);
  cudaFree(d_sink);
  return e;
}

}  // namespace mds
