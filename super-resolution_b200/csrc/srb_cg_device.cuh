// srb_cg_device.cuh -- CUDA vector backend of the CG restatement (srb_cg.h): the solver vectors of
// ALGLIB's mincg (x, g, d, ...; alglib_objective.cpp:47-75 hands ALGLIB host arrays) live in HBM,
// the objective is eval_core() on device pointers, and per line-search step only two scalars
// (f and <g, d>) reach the host, in one 32-byte copy.  All kernels are HBM-bound streaming passes
// (grid = 8 CTAs per SM, grid-stride); reductions are two-stage with a fixed order (one slot per
// CTA, then one CTA sums the slots), hence deterministic run to run.
#pragma once
#include <algorithm>
#include <string>

#include "srb_cg.h"
#include "srb_common.cuh"

namespace srb {

constexpr int CG_NT = 256;
constexpr int CG_MAX_BLOCKS = 2048;

__device__ __forceinline__ void cg_block_store3(double a, double b, double c, double* part, int nblk) {
  __shared__ double sm[3][CG_NT / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if (lane == 0) { sm[0][wid] = a; sm[1][wid] = b; sm[2][wid] = c; }
  __syncthreads();
  if (wid == 0) {
    a = lane < CG_NT / 32 ? sm[0][lane] : 0.0;
    b = lane < CG_NT / 32 ? sm[1][lane] : 0.0;
    c = lane < CG_NT / 32 ? sm[2][lane] : 0.0;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) {
      part[blockIdx.x] = a;
      part[nblk + blockIdx.x] = b;
      part[2 * nblk + blockIdx.x] = c;
    }
  }
}

// MODE 0: <a, b>   1: sum a^2   2: sum (a - b)^2   3: y = a - b: <y, c>, <a, a>, <a, y>
template <int MODE>
__global__ void __launch_bounds__(CG_NT)
k_cg_reduce(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
            long long n, double* __restrict__ part) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double ai = a[i];
    if (MODE == 0) s0 = fma(ai, b[i], s0);
    if (MODE == 1) s0 = fma(ai, ai, s0);
    if (MODE == 2) { const double t = ai - b[i]; s0 = fma(t, t, s0); }
    if (MODE == 3) {
      const double y = ai - b[i];
      s0 = fma(y, c[i], s0);
      s1 = fma(ai, ai, s1);
      s2 = fma(ai, y, s2);
    }
  }
  cg_block_store3(s0, s1, s2, part, gridDim.x);
}

// dk = -g + beta * dk (rounded like ALGLIB's two statements: product, then sum); sum d^2, sum g^2
__global__ void __launch_bounds__(CG_NT)
k_cg_direction(double* __restrict__ dk, const double* __restrict__ g, double beta, const double* __restrict__ d,
               long long n, double* __restrict__ part) {
  double s0 = 0.0, s1 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double gi = g[i], di = d[i];
    dk[i] = __dadd_rn(-gi, __dmul_rn(beta, dk[i]));
    s0 = fma(di, di, s0);
    s1 = fma(gi, gi, s1);
  }
  cg_block_store3(s0, s1, 0.0, part, gridDim.x);
}

__global__ void __launch_bounds__(CG_NT)
k_cg_max_abs(const double* __restrict__ a, long long n, double* __restrict__ part) {
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT)
    m = fmax(m, fabs(a[i]));  // fmax drops NaN: a NaN direction is caught by the isfinite test on |g|
  __shared__ double sm[CG_NT / 32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CG_NT / 32; ++w) m = fmax(m, sm[w]);
    part[blockIdx.x] = m;
  }
}

// out[k] = sum (or max) of part[k * nblk + 0 .. nblk), fixed order; one CTA
__global__ void __launch_bounds__(CG_NT)
k_cg_finish(const double* __restrict__ part, int nblk, int nsums, int is_max, double* __restrict__ out) {
  __shared__ double sm[CG_NT];
  for (int k = 0; k < nsums; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nblk; i += CG_NT) v = is_max ? fmax(v, part[k * nblk + i]) : v + part[k * nblk + i];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = CG_NT / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] = is_max ? fmax(sm[threadIdx.x], sm[threadIdx.x + o]) : sm[threadIdx.x] + sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
    __syncthreads();
  }
}

// MODE 0: dst = -src   1: dst = src * a   2: dst = base(src) + a * dir (product, then sum)
template <int MODE>
__global__ void __launch_bounds__(CG_NT)
k_cg_map(double* dst, const double* src, double a, const double* __restrict__ dir, long long n) {
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    if (MODE == 0) dst[i] = -src[i];
    if (MODE == 1) dst[i] = __dmul_rn(src[i], a);
    if (MODE == 2) dst[i] = __dadd_rn(src[i], __dmul_rn(a, dir[i]));
  }
}

// eval_core(ctx, x, g, tail, data term, regularization term, accumulate) and reweight_dev(ctx, x)
// are srb_api.cu's; this header is included there, after their definitions.

struct DeviceCgBackend {
  using Vec = double*;
  srb_ctx* c;
  long long n;
  double* d_part;   // [3 * CG_MAX_BLOCKS]
  double* d_out;    // [4]: three reduction results + the objective value
  double* h_out;    // pinned mirror
  int nblk;
  srb_status status = SRB_OK;
  long long evals = 0;

  long long size() const { return n; }
  bool ok() const { return status == SRB_OK; }
  void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess && status == SRB_OK)
      status = c->fail(SRB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  // results of the reductions queued so far -> host (one copy, one synchronisation)
  void fetch() {
    check(cudaMemcpyAsync(h_out, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream), "cg fetch");
    check(cudaStreamSynchronize(c->stream), "cg synchronize");
    if (!ok()) h_out[0] = h_out[1] = h_out[2] = h_out[3] = NAN;
  }
  void finish(int nsums, int is_max = 0) {
    k_cg_finish<<<1, CG_NT, 0, c->stream>>>(d_part, nblk, nsums, is_max, d_out);
    c->timing.kernel_launches += 2;
  }

  void eval(Vec x, Vec g, double* f) {
    if (ok()) {
      const srb_status st = eval_core(c, x, g, d_out + 3, true, true, false);
      if (st != SRB_OK) status = st;
    }
    ++evals;
    fetch();
    *f = h_out[3];
  }
  void eval_with_slope(Vec x, Vec g, Vec d, double* f, double* dg) {
    if (ok()) {
      const srb_status st = eval_core(c, x, g, d_out + 3, true, true, false);
      if (st != SRB_OK) status = st;
    }
    ++evals;
    k_cg_reduce<0><<<nblk, CG_NT, 0, c->stream>>>(g, d, nullptr, n, d_part);
    finish(1);
    fetch();
    *f = h_out[3];
    *dg = h_out[0];
  }
  void copy(Vec dst, Vec src) {
    check(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream), "cg copy");
  }
  void neg_copy(Vec dst, Vec src) { k_cg_map<0><<<nblk, CG_NT, 0, c->stream>>>(dst, src, 0.0, nullptr, n); c->timing.kernel_launches += 1; }
  void scale_to(Vec dst, Vec src, double a) { k_cg_map<1><<<nblk, CG_NT, 0, c->stream>>>(dst, src, a, nullptr, n); c->timing.kernel_launches += 1; }
  void scale(Vec v, double a) { scale_to(v, v, a); }
  void step_to(Vec dst, Vec base, double a, Vec dir) { k_cg_map<2><<<nblk, CG_NT, 0, c->stream>>>(dst, base, a, dir, n); c->timing.kernel_launches += 1; }
  void zero(Vec v) { check(cudaMemsetAsync(v, 0, (size_t)n * sizeof(double), c->stream), "cg zero"); }
  double dot(Vec a, Vec b) {
    k_cg_reduce<0><<<nblk, CG_NT, 0, c->stream>>>(a, b, nullptr, n, d_part);
    finish(1);
    fetch();
    return h_out[0];
  }
  double sum_sq(Vec a) {
    k_cg_reduce<1><<<nblk, CG_NT, 0, c->stream>>>(a, nullptr, nullptr, n, d_part);
    finish(1);
    fetch();
    return h_out[0];
  }
  double sum_sq_diff(Vec a, Vec b) {
    k_cg_reduce<2><<<nblk, CG_NT, 0, c->stream>>>(a, b, nullptr, n, d_part);
    finish(1);
    fetch();
    return h_out[0];
  }
  double max_abs(Vec a) {
    k_cg_max_abs<<<nblk, CG_NT, 0, c->stream>>>(a, n, d_part);
    finish(1, 1);
    fetch();
    return h_out[0];
  }
  void beta_terms(Vec gn, Vec go, Vec dk, double* dy, double* gg, double* gy) {
    k_cg_reduce<3><<<nblk, CG_NT, 0, c->stream>>>(gn, go, dk, n, d_part);
    finish(3);
    fetch();
    *dy = h_out[0]; *gg = h_out[1]; *gy = h_out[2];
  }
  void direction(Vec dk, Vec g, double beta, Vec d, double* dd, double* gg) {
    k_cg_direction<<<nblk, CG_NT, 0, c->stream>>>(dk, g, beta, d, n, d_part);
    finish(2);
    fetch();
    *dd = h_out[0]; *gg = h_out[1];
  }
  void reweight(Vec x) {  // w = 1 / max(1e-5, reg(x)), irls_map_solver.cpp:128-143 (stream-ordered)
    if (!ok()) return;
    const srb_status st = reweight_dev(c, x);
    if (st != SRB_OK) status = st;
  }
};

// Scratch vectors, reduction slots and the pinned scalar mirror of one solve.
struct DeviceCgWorkspace {
  double* store = nullptr;
  double* h_out = nullptr;
  double* scratch[kCgScratchVectors];
  ~DeviceCgWorkspace() {
    if (store) cudaFree(store);
    if (h_out) cudaFreeHost(h_out);
  }
  srb_status init(srb_ctx* c, DeviceCgBackend* be) {
    const long long n = (long long)c->n_active();
    const size_t doubles = (size_t)kCgScratchVectors * n + 3 * CG_MAX_BLOCKS + 4;
    if (cudaMalloc((void**)&store, doubles * sizeof(double)) != cudaSuccess) {
      (void)cudaGetLastError();
      store = nullptr;
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (CG vectors)");
    }
    if (cudaMallocHost((void**)&h_out, 4 * sizeof(double)) != cudaSuccess) {
      (void)cudaGetLastError();
      h_out = nullptr;
      return c->fail(SRB_ERR_NOMEM, "cudaMallocHost failed (CG scalars)");
    }
    be->c = c;
    be->n = n;
    be->d_part = store + (size_t)kCgScratchVectors * n;
    be->d_out = be->d_part + 3 * CG_MAX_BLOCKS;
    be->h_out = h_out;
    const long long want = (n + CG_NT - 1) / CG_NT;
    be->nblk = (int)std::max(1LL, std::min<long long>(std::min<long long>(want, (long long)c->num_sms * 8), CG_MAX_BLOCKS));
    for (int i = 0; i < kCgScratchVectors; ++i) scratch[i] = store + (size_t)i * n;
    return SRB_OK;
  }
  srb_status finish(srb_ctx* c, const DeviceCgBackend& be) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (!be.ok()) return be.status;
    if (e != cudaSuccess) return c->fail(SRB_ERR_CUDA, std::string("CG: ") + cudaGetErrorString(e));
    return SRB_OK;
  }
};

// RunCGSolverAnalyticalDiff on a device-resident estimate (active channel range, n = Ca * H * W).
inline srb_status cg_minimize_dev(srb_ctx* c, double* d_x, const CgOptions& opt, CgReport* rep) {
  DeviceCgBackend be;
  DeviceCgWorkspace ws;
  srb_status st = ws.init(c, &be);
  if (st != SRB_OK) return st;
  *rep = cg_minimize(be, d_x, ws.scratch, opt);
  return ws.finish(c, be);
}

// IRLSMapSolver::RunIRLSLoop on a device-resident estimate; the weights must have been reset to 1.
inline srb_status irls_solve_dev(srb_ctx* c, double* d_x, const CgOptions& opt, int max_irls_iterations,
                                 double cost_difference_threshold, bool has_regularizer, IrlsReport* rep) {
  DeviceCgBackend be;
  DeviceCgWorkspace ws;
  srb_status st = ws.init(c, &be);
  if (st != SRB_OK) return st;
  *rep = irls_solve(be, d_x, ws.scratch, opt, max_irls_iterations, cost_difference_threshold, has_regularizer);
  return ws.finish(c, be);
}

}  // namespace srb
