// srb_cg_device.cuh -- CUDA vector backend of the CG restatement (srb_cg.h): the solver vectors of
// ALGLIB's mincg (x, g, d, ...; alglib_objective.cpp:47-75 hands ALGLIB host arrays) live in HBM,
// the objective is eval_core() on device pointers, and per line-search step only three scalars
// (f, <g, d> and |x - x0|^2) reach the host, in one 32-byte copy.  All kernels are HBM-bound streaming passes
// (grid = 8 CTAs per SM, grid-stride); reductions are two-stage with a fixed order (one slot per
// CTA, then one CTA sums the slots), hence deterministic run to run.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

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

// <g, d> of the trial point and the three beta sums of the direction update in one pass over g (trial
// gradient), g0 (gradient at the start of the line search) and dk (the un-normalised direction; the unit
// direction is d = (dk * s1) * s2, rounded as linminnormalized rounds it): with y = g - g0
//   part[0] <g, d>   part[1] <y, dk>   part[2] <g, g>   part[3] <g, y>
__global__ void __launch_bounds__(CG_NT)
k_cg_trial_sums(const double* __restrict__ g, const double* __restrict__ g0, const double* __restrict__ dk,
                double s1, double s2, long long n, double* __restrict__ part) {
  double s0 = 0.0, sy = 0.0, sg = 0.0, sgy = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double gi = g[i], di = dk[i];
    const double y = gi - g0[i];
    s0 = fma(gi, __dmul_rn(__dmul_rn(di, s1), s2), s0);
    sy = fma(y, di, sy);
    sg = fma(gi, gi, sg);
    sgy = fma(gi, y, sgy);
  }
  cg_block_store3(s0, sy, sg, part, gridDim.x);
  __syncthreads();
  cg_block_store3(sgy, 0.0, 0.0, part + 3 * gridDim.x, gridDim.x);
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

// Block maximum of m -> slot[blockIdx.x] (fmax drops NaN: a NaN vector is caught by the isfinite
// test on |g| in cg_minimize).
__device__ __forceinline__ void cg_block_store_max(double m, double* slot) {
  __shared__ double smx[CG_NT / 32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CG_NT / 32; ++w) m = fmax(m, smx[w]);
    slot[blockIdx.x] = m;
  }
}

// dk = -g + beta * dk (rounded like ALGLIB's two statements: product, then sum); sum g^2; max |dk|; and
// for the normalisation that follows (linminnormalized) without another pass: sum dk^2 and <g, dk>
//   part[0] sum g^2   part[1] max |dk|   part[2] sum dk^2   part[3] <g, dk>
__global__ void __launch_bounds__(CG_NT)
k_cg_direction(double* dk, const double* __restrict__ g, double beta, long long n, double* __restrict__ part) {
  double s0 = 0.0, m = 0.0, sd = 0.0, sgd = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double gi = g[i];
    const double v = __dadd_rn(-gi, __dmul_rn(beta, dk[i]));
    dk[i] = v;
    s0 = fma(gi, gi, s0);
    m = fmax(m, fabs(v));
    sd = fma(v, v, sd);
    sgd = fma(gi, v, sgd);
  }
  cg_block_store3(s0, 0.0, sd, part, gridDim.x);
  __syncthreads();
  cg_block_store3(sgd, 0.0, 0.0, part + 3 * gridDim.x, gridDim.x);
  __syncthreads();
  cg_block_store_max(m, part + gridDim.x);  // slot 1 <- max (overwrites the zero sum written above)
}

__global__ void __launch_bounds__(CG_NT)
k_cg_max_abs(const double* __restrict__ a, long long n, double* __restrict__ part) {
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT)
    m = fmax(m, fabs(a[i]));
  cg_block_store_max(m, part);
}

// linminnormalized, pass 1: sum (a_i * s1)^2, the product rounded as ALGLIB's in-place scaling does
__global__ void __launch_bounds__(CG_NT)
k_cg_scaled_sumsq(const double* __restrict__ a, double s1, long long n, double* __restrict__ part) {
  double s0 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double t = __dmul_rn(a[i], s1);
    s0 = fma(t, t, s0);
  }
  cg_block_store3(s0, 0.0, 0.0, part, gridDim.x);
}

// linminnormalized, pass 2: d = (dk * s1) * s2; <g0, d>; sum d^2
__global__ void __launch_bounds__(CG_NT)
k_cg_normalize(double* __restrict__ d, const double* __restrict__ dk, double s1, double s2,
               const double* __restrict__ g0, long long n, double* __restrict__ part) {
  double s0 = 0.0, sq = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double v = __dmul_rn(__dmul_rn(dk[i], s1), s2);
    d[i] = v;
    s0 = fma(g0[i], v, s0);
    sq = fma(v, v, sq);
  }
  cg_block_store3(s0, sq, 0.0, part, gridDim.x);
}

// trial point x = x0 + stp * d (product, then sum) and sum (x0 - x)^2
__global__ void __launch_bounds__(CG_NT)
k_cg_step(double* __restrict__ x, const double* __restrict__ x0, double stp, const double* __restrict__ d,
          long long n, double* __restrict__ part) {
  double s0 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double b = x0[i];
    const double v = __dadd_rn(b, __dmul_rn(stp, d[i]));
    x[i] = v;
    const double t = b - v;
    s0 = fma(t, t, s0);
  }
  cg_block_store3(s0, 0.0, 0.0, part, gridDim.x);
}

// trial point with the unit direction formed on the fly: x = x0 + stp * ((dk * s1) * s2)
__global__ void __launch_bounds__(CG_NT)
k_cg_step_scaled(double* __restrict__ x, const double* __restrict__ x0, double stp, const double* __restrict__ dk,
                 double s1, double s2, long long n, double* __restrict__ part) {
  double s0 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    const double b = x0[i];
    const double v = __dadd_rn(b, __dmul_rn(stp, __dmul_rn(__dmul_rn(dk[i], s1), s2)));
    x[i] = v;
    const double t = b - v;
    s0 = fma(t, t, s0);
  }
  cg_block_store3(s0, 0.0, 0.0, part, gridDim.x);
}

// <a, (dk * s1) * s2>
__global__ void __launch_bounds__(CG_NT)
k_cg_dot_scaled(const double* __restrict__ a, const double* __restrict__ dk, double s1, double s2, long long n,
                double* __restrict__ part) {
  double s0 = 0.0;
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT)
    s0 = fma(a[i], __dmul_rn(__dmul_rn(dk[i], s1), s2), s0);
  cg_block_store3(s0, 0.0, 0.0, part, gridDim.x);
}

// out[k] = sum (or max, where bit k of max_mask is set) of part[k * nblk + 0 .. nblk), fixed order
__global__ void __launch_bounds__(CG_NT)
k_cg_finish(const double* __restrict__ part, int nblk, int nsums, int max_mask, double* __restrict__ out) {
  __shared__ double sm[CG_NT];
  for (int k = 0; k < nsums; ++k) {
    const bool is_max = (max_mask >> k) & 1;
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

// MODE 0: dst += src   1: dst += a * src   2: dst -= a * src   3: dst *= a   (products rounded first)
template <int MODE>
__global__ void __launch_bounds__(CG_NT)
k_cg_update(double* __restrict__ dst, double a, const double* __restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT) {
    if (MODE == 0) dst[i] = __dadd_rn(dst[i], src[i]);
    if (MODE == 1) dst[i] = __dadd_rn(dst[i], __dmul_rn(a, src[i]));
    if (MODE == 2) dst[i] = __dsub_rn(dst[i], __dmul_rn(a, src[i]));
    if (MODE == 3) dst[i] = __dmul_rn(dst[i], a);
  }
}

__global__ void __launch_bounds__(CG_NT)
k_cg_neg_copy(double* __restrict__ dst, const double* __restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < n; i += (long long)gridDim.x * CG_NT)
    dst[i] = -src[i];
}

// eval_core(ctx, x, g, tail, data term, regularization term, accumulate) and reweight_dev(ctx, x)
// are srb_api.cu's; this header is included there, after their definitions.

struct DeviceCgBackend {
  using Vec = double*;
  srb_ctx* c;
  long long n;
  double* d_part;   // [6 * CG_MAX_BLOCKS]
  double* d_out;    // [8]: 0..3 reduction results, 4 = |x - x0|^2 of the trial step, 5 = objective value
  double* h_out;    // pinned mirror
  int nblk;
  srb_status status = SRB_OK;
  long long evals = 0;

  // The unit search direction d = (dk * s1) * s2 of linminnormalized is never stored: normalize_to()
  // records (d, dk, s1, s2) and trial() / dot() form it on the fly from dk -- same rounding, one vector
  // write and one read less per line-search step.
  Vec unit_d = nullptr, unit_src = nullptr, unit_g0 = nullptr;
  double unit_s1 = 1.0, unit_s2 = 1.0;
  // sums that direction() produced for the normalisation that follows it
  Vec dir_dk = nullptr, dir_g = nullptr;
  double dir_sumsq = 0.0, dir_gdk = 0.0;
  // beta sums that the last trial() produced (valid for trial_g against unit_g0 / unit_src)
  Vec trial_g = nullptr;
  double trial_dy = 0.0, trial_gg = 0.0, trial_gy = 0.0;

  long long size() const { return n; }
  bool ok() const { return status == SRB_OK; }
  void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess && status == SRB_OK)
      status = c->fail(SRB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  // results of the reductions queued so far -> host (one copy, one synchronisation)
  void fetch() {
    check(cudaMemcpyAsync(h_out, d_out, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream), "cg fetch");
    check(cudaStreamSynchronize(c->stream), "cg synchronize");
    if (!ok())
      for (int i = 0; i < 8; ++i) h_out[i] = NAN;
  }
  // partial sums of the kernel just launched -> d_out[offset .. offset + nsums)
  void finish(int nsums, int max_mask = 0, int offset = 0) {
    k_cg_finish<<<1, CG_NT, 0, c->stream>>>(d_part, nblk, nsums, max_mask, d_out + offset);
    c->timing.kernel_launches += 2;
  }
  void forget(Vec v) {  // v is about to be overwritten: drop what was cached about it
    if (v == unit_d || v == unit_src) unit_d = unit_src = nullptr;
    if (v == dir_dk || v == dir_g) dir_dk = dir_g = nullptr;
    if (v == trial_g) trial_g = nullptr;
    if (v == unit_g0) trial_g = unit_g0 = nullptr;
  }

  void eval(Vec x, Vec g, double* f) {
    forget(g);
    if (ok()) {
      const srb_status st = eval_core(c, x, g, d_out + 5, true, true, false);
      if (st != SRB_OK) status = st;
    }
    ++evals;
    fetch();
    *f = h_out[5];
  }
  void copy(Vec dst, Vec src) {
    forget(dst);
    check(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream), "cg copy");
  }
  void neg_copy(Vec dst, Vec src) {
    forget(dst);
    k_cg_neg_copy<<<nblk, CG_NT, 0, c->stream>>>(dst, src, n);
    c->timing.kernel_launches += 1;
  }
  void zero(Vec v) {
    forget(v);
    check(cudaMemsetAsync(v, 0, (size_t)n * sizeof(double), c->stream), "cg zero");
  }
  double dot(Vec a, Vec b) {
    if (b == unit_d && unit_src) {
      k_cg_dot_scaled<<<nblk, CG_NT, 0, c->stream>>>(a, unit_src, unit_s1, unit_s2, n, d_part);
    } else if (a == unit_d && unit_src) {
      k_cg_dot_scaled<<<nblk, CG_NT, 0, c->stream>>>(b, unit_src, unit_s1, unit_s2, n, d_part);
    } else {
      k_cg_reduce<0><<<nblk, CG_NT, 0, c->stream>>>(a, b, nullptr, n, d_part);
    }
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
  double max_abs(Vec a) {
    k_cg_max_abs<<<nblk, CG_NT, 0, c->stream>>>(a, n, d_part);
    finish(1, 1);
    fetch();
    return h_out[0];
  }
  void normalize_to(Vec d, Vec dk, double mx, Vec g0, double* stp, double* slope, double* dd) {
    forget(d);
    if (mx == 0.0) {  // zero direction: d = dk, slope and length are zero
      copy(d, dk);
      *slope = 0.0;
      *dd = 0.0;
      return;
    }
    const double s1 = 1 / mx;
    double sumsq_scaled, gdk;
    bool have = dir_dk == dk && dir_g == g0;
    if (have) {
      // sums of the un-normalised direction from the pass that built it; scaling them instead of the
      // elements changes the result by rounding only.  Not when the scaling matters for the range.
      sumsq_scaled = dir_sumsq * s1 * s1;
      gdk = dir_gdk;
      have = std::isfinite(dir_sumsq) && dir_sumsq > 1e-280 && std::isfinite(sumsq_scaled) && sumsq_scaled > 0.0;
    }
    if (!have) {
      k_cg_scaled_sumsq<<<nblk, CG_NT, 0, c->stream>>>(dk, s1, n, d_part);
      finish(1);
      k_cg_reduce<0><<<nblk, CG_NT, 0, c->stream>>>(g0, dk, nullptr, n, d_part);
      finish(1, 0, 1);
      fetch();
      sumsq_scaled = h_out[0];
      gdk = h_out[1];
    }
    const double s2 = 1 / std::sqrt(sumsq_scaled);
    unit_d = d; unit_src = dk; unit_g0 = g0; unit_s1 = s1; unit_s2 = s2;
    trial_g = nullptr;
    *stp = *stp / s1;
    *stp = *stp / s2;
    *slope = gdk * s1 * s2;
    *dd = sumsq_scaled * s2 * s2;
  }
  void trial(Vec x, Vec x0, double stp, Vec d, Vec g, double* f, double* dg, double* moved) {
    forget(x);
    const bool unit = d == unit_d && unit_src != nullptr;
    if (unit) k_cg_step_scaled<<<nblk, CG_NT, 0, c->stream>>>(x, x0, stp, unit_src, unit_s1, unit_s2, n, d_part);
    else k_cg_step<<<nblk, CG_NT, 0, c->stream>>>(x, x0, stp, d, n, d_part);
    finish(1, 0, 4);  // moved -> d_out[4]
    if (g == trial_g) trial_g = nullptr;
    if (ok()) {
      const srb_status st = eval_core(c, x, g, d_out + 5, true, true, false);
      if (st != SRB_OK) status = st;
    }
    ++evals;
    const bool sums = unit && unit_g0 != nullptr && g != unit_g0;  // (L-BFGS searches overwrite g0 itself)
    if (sums) {
      k_cg_trial_sums<<<nblk, CG_NT, 0, c->stream>>>(g, unit_g0, unit_src, unit_s1, unit_s2, n, d_part);
      finish(4);      // <g, d>, <y, dk>, <g, g>, <g, y> -> d_out[0..3]
    } else if (unit) {
      k_cg_dot_scaled<<<nblk, CG_NT, 0, c->stream>>>(g, unit_src, unit_s1, unit_s2, n, d_part);
      finish(1);
    } else {
      k_cg_reduce<0><<<nblk, CG_NT, 0, c->stream>>>(g, d, nullptr, n, d_part);
      finish(1);      // <g, d> -> d_out[0]
    }
    fetch();
    *f = h_out[5];
    *dg = h_out[0];
    *moved = h_out[4];
    if (sums) {
      trial_g = g;
      trial_dy = h_out[1]; trial_gg = h_out[2]; trial_gy = h_out[3];
    }
  }
  void beta_terms(Vec gn, Vec go, Vec dk, double* dy, double* gg, double* gy) {
    if (gn == trial_g && go == unit_g0 && dk == unit_src) {  // the accepted trial already summed them
      *dy = trial_dy; *gg = trial_gg; *gy = trial_gy;
      return;
    }
    k_cg_reduce<3><<<nblk, CG_NT, 0, c->stream>>>(gn, go, dk, n, d_part);
    finish(3);
    fetch();
    *dy = h_out[0]; *gg = h_out[1]; *gy = h_out[2];
  }
  void direction(Vec dk, Vec g, double beta, double* gg, double* mx) {
    forget(dk);
    k_cg_direction<<<nblk, CG_NT, 0, c->stream>>>(dk, g, beta, n, d_part);
    finish(4, 2);     // slot 0: sum g^2, slot 1: max |dk|, slot 2: sum dk^2, slot 3: <g, dk>
    fetch();
    *gg = h_out[0]; *mx = h_out[1];
    dir_dk = dk; dir_g = g; dir_sumsq = h_out[2]; dir_gdk = h_out[3];
  }
  // L-BFGS only (two-loop recursion, pair updates)
  void add(Vec dst, Vec src) { forget(dst); k_cg_update<0><<<nblk, CG_NT, 0, c->stream>>>(dst, 0.0, src, n); c->timing.kernel_launches += 1; }
  void add_scaled(Vec dst, double a, Vec src) { forget(dst); k_cg_update<1><<<nblk, CG_NT, 0, c->stream>>>(dst, a, src, n); c->timing.kernel_launches += 1; }
  void sub_scaled(Vec dst, double a, Vec src) { forget(dst); k_cg_update<2><<<nblk, CG_NT, 0, c->stream>>>(dst, a, src, n); c->timing.kernel_launches += 1; }
  void scale(Vec v, double a) { forget(v); k_cg_update<3><<<nblk, CG_NT, 0, c->stream>>>(v, a, nullptr, n); c->timing.kernel_launches += 1; }
  void reweight(Vec x) {  // w = 1 / max(1e-5, reg(x)), irls_map_solver.cpp:128-143 (stream-ordered)
    if (!ok()) return;
    const srb_status st = reweight_dev(c, x);
    if (st != SRB_OK) status = st;
  }
};

// Scratch vectors, reduction slots and the pinned scalar mirror of a solve.  The storage belongs to the
// context and is kept between solves (cudaMalloc / cudaFree of five 100 MB vectors cost more than twenty
// CG iterations at cfg3); srb_destroy releases it.
struct DeviceCgWorkspace {
  std::vector<double*> scratch;
  srb_status init(srb_ctx* c, DeviceCgBackend* be, int num_vectors = kCgScratchVectors) {
    const long long n = (long long)c->n_active();
    const size_t doubles = (size_t)num_vectors * n + 6 * CG_MAX_BLOCKS + 8;
    if (doubles > c->cg_store_doubles) {
      if (c->cg_store) cudaFree(c->cg_store);
      c->cg_store = nullptr;
      c->cg_store_doubles = 0;
      if (cudaMalloc((void**)&c->cg_store, doubles * sizeof(double)) != cudaSuccess) {
        (void)cudaGetLastError();
        c->cg_store = nullptr;
        return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (CG vectors)");
      }
      c->cg_store_doubles = doubles;
    }
    if (!c->cg_h_out && cudaMallocHost((void**)&c->cg_h_out, 8 * sizeof(double)) != cudaSuccess) {
      (void)cudaGetLastError();
      c->cg_h_out = nullptr;
      return c->fail(SRB_ERR_NOMEM, "cudaMallocHost failed (CG scalars)");
    }
    double* store = c->cg_store;
    be->c = c;
    be->n = n;
    be->d_part = store + (size_t)num_vectors * n;
    be->d_out = be->d_part + 6 * CG_MAX_BLOCKS;
    be->h_out = c->cg_h_out;
    const long long want = (n + CG_NT - 1) / CG_NT;
    be->nblk = (int)std::max(1LL, std::min<long long>(std::min<long long>(want, (long long)c->num_sms * 8), CG_MAX_BLOCKS));
    scratch.resize(num_vectors);
    for (int i = 0; i < num_vectors; ++i) scratch[i] = store + (size_t)i * n;
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
  *rep = cg_minimize(be, d_x, ws.scratch.data(), opt);
  return ws.finish(c, be);
}

// RunLBFGSSolverAnalyticalDiff (alglib_objective.cpp:111-140) on a device-resident estimate.
inline srb_status lbfgs_minimize_dev(srb_ctx* c, double* d_x, int m, const CgOptions& opt, CgReport* rep) {
  DeviceCgBackend be;
  DeviceCgWorkspace ws;
  srb_status st = ws.init(c, &be, lbfgs_scratch_vectors(m));
  if (st != SRB_OK) return st;
  *rep = lbfgs_minimize(be, d_x, ws.scratch.data(), m, opt);
  return ws.finish(c, be);
}

// IRLSMapSolver::RunIRLSLoop on a device-resident estimate; the weights must have been reset to 1.
inline srb_status irls_solve_dev(srb_ctx* c, double* d_x, const CgOptions& opt, int max_irls_iterations,
                                 double cost_difference_threshold, bool has_regularizer, int lbfgs_corrections,
                                 IrlsReport* rep) {
  DeviceCgBackend be;
  DeviceCgWorkspace ws;
  srb_status st = ws.init(c, &be, lbfgs_corrections > 0 ? lbfgs_scratch_vectors(lbfgs_corrections) : kCgScratchVectors);
  if (st != SRB_OK) return st;
  *rep = irls_solve(be, d_x, ws.scratch.data(), opt, max_irls_iterations, cost_difference_threshold,
                    has_regularizer, lbfgs_corrections);
  return ws.finish(c, be);
}

}  // namespace srb
