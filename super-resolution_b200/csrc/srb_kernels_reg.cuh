// srb_kernels_reg.cuh -- TV / 3-D TV / BTV regularizer kernels (values, partial derivatives,
// IRLS term, IRLS re-weighting).
//
// Follows src/optimization/tv_regularizer.cpp:21-227, btv_regularizer.cpp:19-170 and
// objective_irls_regularization_term.cpp:10-58 including their quirks: the 3-D
// TV self term has no z component, BTV values use the inclusive window 0..R while its gradient
// uses 0..R-1, and the BTV neighbour loop skips image pixel (0,0).  Operation order and explicit
// round-to-nearest arithmetic make values and partials bit-identical to the reference.
#pragma once
#include "srb_common.cuh"

namespace srb {

struct RegParams {
  int H, W, C;
  int kind;             // SRB_REG_TV / TV3D / BTV
  int R;                // BTV scale range
  const double* decay;  // [2R+1] pow(spatial_decay, i+j), host-computed
};

__device__ __forceinline__ double sgn_pos(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

#define SRB_IDX(c, r, col) ((size_t)(c) * HW + (size_t)(r) * P.W + (col))

// tv_regularizer.cpp:21-55
__device__ __forceinline__ double tv_gx(const RegParams& P, const double* __restrict__ x, size_t HW,
                                        int c, int r, int col) {
  return (col >= 0 && col + 1 < P.W) ? __dadd_rn(x[SRB_IDX(c, r, col + 1)], -x[SRB_IDX(c, r, col)])
                                     : 0.0;
}
__device__ __forceinline__ double tv_gy(const RegParams& P, const double* __restrict__ x, size_t HW,
                                        int c, int r, int col) {
  return (r >= 0 && r + 1 < P.H) ? __dadd_rn(x[SRB_IDX(c, r + 1, col)], -x[SRB_IDX(c, r, col)])
                                 : 0.0;
}
// tv_regularizer.cpp:57-70
__device__ __forceinline__ double tv_gz(const RegParams& P, const double* __restrict__ x, size_t HW,
                                        int c, int r, int col) {
  return __dadd_rn(x[SRB_IDX(c + 1, r, col)], -x[SRB_IDX(c, r, col)]);
}
// tv_regularizer.cpp:72-107: |gy| + |gx| (+ |gz|)
__device__ __forceinline__ double tv_value(const RegParams& P, const double* __restrict__ x,
                                           size_t HW, int c, int r, int col) {
  double tv = __dadd_rn(fabs(tv_gy(P, x, HW, c, r, col)), fabs(tv_gx(P, x, HW, c, r, col)));
  if (P.kind == SRB_REG_TV3D && c + 1 < P.C) tv = __dadd_rn(tv, fabs(tv_gz(P, x, HW, c, r, col)));
  return tv;
}
// btv_regularizer.cpp:19-46
__device__ __forceinline__ double btv_value(const RegParams& P, const double* __restrict__ x,
                                            size_t HW, int c, int r, int col) {
  const double xp = x[SRB_IDX(c, r, col)];
  double tv = 0.0;
  for (int i = 0; i <= P.R; ++i) {
    const int orow = r + i;
    if (orow >= P.H) break;
    for (int j = 0; j <= P.R; ++j) {
      const int ocol = col + j;
      if (ocol >= P.W) break;
      tv = __dadd_rn(tv, __dmul_rn(P.decay[i + j], fabs(__dadd_rn(xp, -x[SRB_IDX(c, orow, ocol)]))));
    }
  }
  return tv;
}
__device__ __forceinline__ double reg_value(const RegParams& P, const double* __restrict__ x,
                                            size_t HW, int c, int r, int col) {
  return P.kind == SRB_REG_BTV ? btv_value(P, x, HW, c, r, col) : tv_value(P, x, HW, c, r, col);
}

// Regularizer::ApplyToImage.  mode 0: out = r(x);  mode 1 (IRLS re-weighting,
// irls_map_solver.cpp:128-143): out = 1 / max(1e-5, r(x)).
// grid: (ceil(W/32), ceil(H/8), C)
template <int kMode>
__global__ void __launch_bounds__(256)
k_reg_values(RegParams P, const double* __restrict__ x, double* __restrict__ out) {
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r = blockIdx.y * 8 + threadIdx.y;
  const int c = blockIdx.z;
  if (col >= P.W || r >= P.H) return;
  const size_t HW = (size_t)P.H * P.W;
  const double v = reg_value(P, x, HW, c, r, col);
  out[SRB_IDX(c, r, col)] = (kMode == 0) ? v : __ddiv_rn(1.0, v > 0.00001 ? v : 0.00001);
}

// Partial derivative of sum_j cst_j * r_j^2 at one pixel, given the value plane `vals`.
__device__ __forceinline__ double reg_partial(const RegParams& P, const double* __restrict__ x,
                                              const double* __restrict__ vals,
                                              const double* __restrict__ cst, double lambda,
                                              size_t HW, int c, int r, int col) {
  // gradient constant of pixel i: the caller's vector, or lambda * w_i
  // (objective_irls_regularization_term.cpp:27-32)
#define SRB_CST(i) (lambda == 0.0 ? cst[i] : __dmul_rn(lambda, cst[i]))
#define SRB_TERM(i, sgn) __dmul_rn(__dmul_rn(__dmul_rn(2.0, SRB_CST(i)), vals[i]), (sgn))
  const size_t index = SRB_IDX(c, r, col);
  double g = 0.0;
  if (P.kind != SRB_REG_BTV) {
    // tv_regularizer.cpp:152-170 (self term; no z component even with 3-D TV)
    double didi = 0.0;
    const double gx = tv_gx(P, x, HW, c, r, col);
    if (gx < 0.0) didi += 1.0; else if (gx > 0.0) didi -= 1.0;
    const double gy = tv_gy(P, x, HW, c, r, col);
    if (gy < 0.0) didi += 1.0; else if (gy > 0.0) didi -= 1.0;
    g = __dadd_rn(g, SRB_TERM(index, didi));
    if (col - 1 >= 0) {  // :171-184
      const size_t li = SRB_IDX(c, r, col - 1);
      g = __dadd_rn(g, SRB_TERM(li, sgn_pos(tv_gx(P, x, HW, c, r, col - 1))));
    }
    if (r - 1 >= 0) {  // :185-201
      const size_t ai = SRB_IDX(c, r - 1, col);
      g = __dadd_rn(g, SRB_TERM(ai, sgn_pos(tv_gy(P, x, HW, c, r - 1, col))));
    }
    if (P.kind == SRB_REG_TV3D && c > 0) {  // :202-220
      const size_t bi = SRB_IDX(c - 1, r, col);
      g = __dadd_rn(g, SRB_TERM(bi, sgn_pos(tv_gz(P, x, HW, c - 1, r, col))));
    }
  } else {
    const double xp = x[index];
    // btv_regularizer.cpp:113-136 (self term, exclusive window)
    double didi = 0.0;
    for (int i = 0; i < P.R; ++i) {
      const int orow = r + i;
      if (orow >= P.H) break;
      for (int j = 0; j < P.R; ++j) {
        const int ocol = col + j;
        if (ocol >= P.W) break;
        const double diff = __dadd_rn(xp, -x[SRB_IDX(c, orow, ocol)]);
        didi = __dadd_rn(didi, __dmul_rn(P.decay[i + j], sgn_pos(diff)));
      }
    }
    g = __dadd_rn(g, SRB_TERM(index, didi));
    // :137-165 (pixels whose window covers this one; image pixel (0,0) is skipped)
    for (int i = 0; i < P.R; ++i) {
      const int orow = r - i;
      if (orow < 0) break;
      for (int j = 0; j < P.R; ++j) {
        const int ocol = col - j;
        if (ocol < 0) break;
        if (orow == 0 && ocol == 0) continue;
        const size_t oi = SRB_IDX(c, orow, ocol);
        const double diff = __dadd_rn(x[oi], -xp);
        double didj = 0.0;
        if (diff < 0.0) didj = 1.0; else if (diff > 0.0) didj = -1.0;
        didj = __dmul_rn(didj, P.decay[i + j]);
        g = __dadd_rn(g, SRB_TERM(oi, didj));
      }
    }
  }
#undef SRB_TERM
#undef SRB_CST
  return g;
}

// Regularizer::ApplyToImageWithDifferentiation, second half: partials from the value plane.
//   mode 0: partials_out = d/dx sum cst*r^2                 (caller-supplied constants, lambda = 0)
//   mode 1: IRLS term (objective_irls_regularization_term.cpp:46-55) restricted to HR rows
//           [row0,row1): grad += partial (cst = IRLS weights, lambda != 0) and per-block partial
//           sums of lambda*w*r*r into cost_partial.  grad may be NULL (cost only).
template <int kMode>
__global__ void __launch_bounds__(256)
k_reg_partials(RegParams P, const double* __restrict__ x, const double* __restrict__ vals,
               const double* __restrict__ cst, double lambda, int row0, int row1,
               double* __restrict__ out, double* __restrict__ cost_partial) {
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r = row0 + blockIdx.y * 8 + threadIdx.y;
  const int c = blockIdx.z;
  double cost = 0.0;
  if (col < P.W && r < row1) {
    const size_t HW = (size_t)P.H * P.W;
    const size_t index = SRB_IDX(c, r, col);
    if (kMode == 0) {
      out[index] = reg_partial(P, x, vals, cst, lambda, HW, c, r, col);
    } else {
      if (out) out[index] = __dadd_rn(out[index], reg_partial(P, x, vals, cst, lambda, HW, c, r, col));
      const double v = vals[index];
      cost = __dmul_rn(__dmul_rn(__dmul_rn(lambda, cst[index]), v), v);
    }
  }
  if (kMode == 1) {
    const double bs = block_sum(cost);
    if (threadIdx.x == 0 && threadIdx.y == 0)
      cost_partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = bs;
  }
}

// The IRLS regularization cost summed in the reference's order (objective_irls_regularization_term.cpp:
// 44-55): residual_sum += lambda * w_i * r_i * r_i, i ascending.  One thread (srb_set_strict_cost).
__global__ void k_strict_reg_cost(const double* __restrict__ vals, const double* __restrict__ w, double lambda, size_t n,
                                  double* __restrict__ cost_out) {
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double v = vals[i];
    sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(__dmul_rn(lambda, w[i]), v), v));
  }
  *cost_out = sum;
}

__global__ void k_fill(double* __restrict__ p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}


// cost[2] = cost[0] + cost[1]; optionally also written behind the gradient (multi-GPU form).
__global__ void k_finish_cost(double* __restrict__ cost, double* __restrict__ tail) {
  const double t = __dadd_rn(__dadd_rn(0.0, cost[0]), cost[1]);
  cost[2] = t;
  if (tail) *tail = t;
}

#undef SRB_IDX
}  // namespace srb
