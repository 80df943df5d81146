// microbench_peaks.cu — measured FP64 (DFMA) and shared-memory peaks of the GPU it runs on.
// The replay kernels keep each env's state in registers / shared memory, so HBM is not their roofline;
// these two numbers are.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o scripts/microbench_peaks scripts/microbench_peaks.cu ; prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, const double a, const double b, const int iters) {
  double x[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) x[k] = (double)(threadIdx.x + k) * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) x[k] = __fma_rn(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += x[k];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

// every thread streams LDS.128 over a conflict-free window of the CTA's shared memory
__global__ void __launch_bounds__(256) lds_kernel(double* out, const int iters) {
  extern __shared__ double2 sm[];
  const int n = 4096;                                  // 64 KiB
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double2 acc = make_double2(0.0, 0.0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const double2 v = sm[(idx + r * 256) & (n - 1)];
      acc.x += v.x; acc.y += v.y;
    }
    idx = (idx + 37 * 256) & (n - 1);
  }
  if (acc.x == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

// LDS.128 + STS.128 round trip (the access pattern of an in-place sweep over a 64-KiB state)
__global__ void __launch_bounds__(256) ldsts_kernel(double* out, const int iters) {
  extern __shared__ double2 sm[];
  const int n = 4096;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  volatile double2* vs = sm;                           // every access really goes to shared memory
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int j = (threadIdx.x + r * 256) & (n - 1);
      double2 v;
      v.x = vs[j].x; v.y = vs[j].y;
      v.x += 1.0;
      vs[j].x = v.x; vs[j].y = v.y;
    }
  }
  __syncthreads();
  if (sm[threadIdx.x].x == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x].y;
}

template <typename F>
static float best_ms(F&& launch, int reps) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CHECK(cudaEventRecord(e0));
    launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CHECK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
  CHECK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CHECK(cudaFuncSetAttribute(ldsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));

  // ---- DFMA: 8 independent chains per thread, 8 CTAs of 256 threads per SM
  const int iters = 4096, ctas = sms * 8;
  for (int w = 0; w < 3; ++w) dfma_kernel<8><<<ctas, 256>>>(out, 1.0000001, 1e-9, iters);
  CHECK(cudaDeviceSynchronize());
  const float ms_f = best_ms([&] { dfma_kernel<8><<<ctas, 256>>>(out, 1.0000001, 1e-9, iters); }, 10);
  const double fma = (double)ctas * 256 * iters * 8 * 8;
  const double tflops = 2.0 * fma / (ms_f * 1e-3) / 1e12;

  // ---- shared memory: 3 CTAs of 64 KiB per SM
  const int it2 = 2048, ctas2 = sms * 3;
  for (int w = 0; w < 3; ++w) lds_kernel<<<ctas2, 256, 65536>>>(out, it2);
  CHECK(cudaDeviceSynchronize());
  const float ms_l = best_ms([&] { lds_kernel<<<ctas2, 256, 65536>>>(out, it2); }, 10);
  const double lds_bytes = (double)ctas2 * 256 * it2 * 16 * 16;
  for (int w = 0; w < 3; ++w) ldsts_kernel<<<ctas2, 256, 65536>>>(out, it2);
  CHECK(cudaDeviceSynchronize());
  const float ms_s = best_ms([&] { ldsts_kernel<<<ctas2, 256, 65536>>>(out, it2); }, 10);
  const double ldsts_bytes = (double)ctas2 * 256 * it2 * 16 * 32;

  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu_name\": \"%s\", \"sms\": %d, \"sm_clock_khz_attr\": %d, "
         "\"fp64_dfma_tflops\": %.3f, \"fp64_dfma_per_clk_per_sm\": %.2f, "
         "\"smem_lds128_gbs\": %.1f, \"smem_lds128_bytes_per_clk_per_sm\": %.1f, "
         "\"smem_lds_sts128_gbs\": %.1f, "
         "\"how\": \"dfma: 8 independent __fma_rn chains/thread, 8x256 threads/SM, best of 10 (2 flop per FMA); "
         "smem: LDS.128 (and a volatile 64-bit load+store round trip) conflict-free over 64 KiB, 3 CTAs x 256 threads/SM, best of 10; "
         "per-clk figures use the cudaDevAttrClockRate clock\"}\n",
         p.name, sms, clk, tflops, fma / (ms_f * 1e-3) / ((double)clk * 1e3) / sms,
         lds_bytes / (ms_l * 1e-3) / 1e9, lds_bytes / (ms_l * 1e-3) / ((double)clk * 1e3) / sms,
         ldsts_bytes / (ms_s * 1e-3) / 1e9);
  return 0;
}
