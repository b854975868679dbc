// ubench.cu -- micro-benchmarks that bound the fused kernel's design on B200:
//   fp64 FMA throughput, fp64 add/mul throughput, streaming copy bandwidth, shared-memory
//   LDS.64 bandwidth.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_dadd(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
    a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_ffma(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void k_read(const double2* __restrict__ in, double* __restrict__ out, size_t n) {
  double acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double2 v = in[i]; acc += v.x + v.y; }
  if (acc == 1.2345) out[0] = acc;
}
__global__ void k_lds(double* out, int iters) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += sm[(idx + u * 37) & 4095];
    idx = (idx + 1) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  float ms = timeit([&] { k_dfma<<<blocks, threads>>>(d, iters); });
  printf("DFMA: %.2f TFLOP/s (%.1f DFMA/clk/SM at %d kHz nominal)\n", 2.0 * 8 * iters * blocks * threads / ms / 1e9, 8.0 * iters * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate);
  ms = timeit([&] { k_dadd<<<blocks, threads>>>(d, iters); });
  printf("DADD: %.2f Tadd/s\n", 8.0 * iters * blocks * threads / ms / 1e9);
  ms = timeit([&] { k_ffma<<<blocks, threads>>>((float*)d, iters); });
  printf("FFMA: %.2f TFLOP/s\n", 2.0 * 8 * iters * blocks * threads / ms / 1e9);
  ms = timeit([&] { k_lds<<<blocks, threads>>>(d, 4000); });
  printf("LDS.64: %.2f TB/s aggregate (%.1f B/clk/SM nominal)\n", 8.0 * 8 * 4000 * blocks * threads / ms / 1e9, 8.0 * 8 * 4000 * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  size_t n = (size_t)1 << 26;  // 64M double2 = 1 GiB
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
  ms = timeit([&] { k_copy<<<p.multiProcessorCount * 16, 512>>>(a, b, n); });
  printf("copy 1GiB->1GiB: %.1f GB/s (read+write)\n", 2.0 * n * 16 / ms / 1e6);
  ms = timeit([&] { k_read<<<p.multiProcessorCount * 16, 512>>>(a, d, n); });
  printf("read 1GiB: %.1f GB/s\n", 1.0 * n * 16 / ms / 1e6);
  size_t m = (size_t)100663296 / 16;  // ~100 MB, the cfg3 plane size
  ms = timeit([&] { k_copy<<<p.multiProcessorCount * 16, 512>>>(a, b, m); });
  printf("copy 100MB->100MB: %.1f GB/s (read+write), %.1f us\n", 2.0 * m * 16 / ms / 1e6, ms * 1e3);
  return 0;
}
