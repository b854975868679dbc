// tma_probe.cu -- which (box, coordinate) combinations does a FLOAT64 tiled TMA load accept on B200?
// usage: tma_probe W H box_w box_h cx cy
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int cx, int cy, int n, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* dst = (double*)sm;
  unsigned long long* bar = (unsigned long long*)(sm + ((n * 8 + 127) & ~127));
  unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 8));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(d), "l"((unsigned long long)&m), "r"(cx), "r"(cy), "r"(0), "r"(b) : "memory");
  }
  unsigned done = 0;
  while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), cx = atoi(argv[5]), cy = atoi(argv[6]);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  std::vector<double> h((size_t)W * H);
  for (int i = 0; i < W * H; ++i) h[i] = i + 1;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, (size_t)bw * bh * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, 1}, str[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((enc_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 0; }
  int n = bw * bh;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8 + 256);
  k<<<1, 128, n * 8 + 256>>>(m, cx, cy, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL %s\n", cudaGetErrorString(e)); return 0; }
  std::vector<double> res(n); cudaMemcpy(res.data(), o, n * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw; ++c) {
    int gr = cy + r2, gc = cx + c; double exp = (gr >= 0 && gr < H && gc >= 0 && gc < W) ? h[(size_t)gr * W + gc] : 0.0;
    if (res[r2 * bw + c] != exp) ++bad;
  }
  printf("ok, mismatches %d\n", bad);
  return 0;
}
