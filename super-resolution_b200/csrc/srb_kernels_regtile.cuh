// srb_kernels_regtile.cuh -- tiled IRLS regularization kernels of the fused path for the two regularizers
// that are not evaluated inside the tile kernel itself: bilateral total variation and 3-D total variation.
//
// One launch per evaluation (or per range of gradient units), on the same (channel, 32-row, 64-column) tile
// grid as k_tile / k_tile_zt and right behind it on the stream: the kernel ADDS the partial derivatives of
//     sum_q lambda w_q r_q(x)^2                  (objective_irls_regularization_term.cpp:10-58)
// into the gradient rows the tile kernel has just written and leaves the cost of its tile in the tile's
// regularization slot.  They replace, on the fused path, the three reference-order launches
// (k_reg_values -> k_reg_partials -> k_reduce_partials, srb_kernels_reg.cuh), which stay what
// SRB_PATH_REFERENCE_ORDER, srb_irls_term and srb_reg_apply* run: same mathematics including the reference's
// quirks, but the sums are formed in this kernel's own order (fused-path bar: 1e-12 relative, not bit identity).
//
//   BTV  (btv_regularizer.cpp:19-170):  r_q = sum_{i,j=0..R} a^(i+j) |x_q - x_{q+(i,j)}|   (inclusive window)
//        d/dx_p = t_p sum_{i,j=0..R-1} a^(i+j) sgn(x_p - x_{p+(i,j)})                      (exclusive window)
//               + sum_{i,j=0..R-1, q=p-(i,j)} t_q a^(i+j) sgn(x_p - x_q),   t_q = 2 lambda w_q r_q,
//        taps outside the image dropped, q = image pixel (0,0) skipped in the second sum (:143-146).
//        x tile + halo in shared memory; t over the tile + the R-1 rows / columns above / left of it is formed
//        once per CTA in shared memory (16 taps each), then every pixel reads its 3x3 (R = 3) neighbourhoods.
//   3-D TV (tv_regularizer.cpp:72-107, 134-227): r_q = |gy| + |gx| + [c+1 < C] |gz|,
//        d/dx_p = -t_p (sgn gx_p + sgn gy_p) + t_l sgn gx_l + t_a sgn gy_a + t_b sgn gz_b   (l / a / b = left,
//        above, previous channel; the self term has no z part: the reference's quirk).  A thread walks 8 rows of
//        one column; everything comes from coalesced global loads (the three channel planes involved stay in
//        the 126 MB L2 because the grid runs channel by channel).
#pragma once
#include "srb_common.cuh"
#include "srb_kernels_tile.cuh"

namespace srb {

struct RegTileParams {
  int H, W, Ca;
  int row0, row1;              // HR row band of the regularization term on this rank
  int unit_begin, tile_rows;   // first (channel, tile row) unit of this launch; tile rows per channel
  const double* x;             // [Ca][H][W]
  const double* w;             // IRLS weights, same shape
  double* g;                   // gradient, same shape (may be NULL: cost only)
  double two_lambda;
  double decay[9];             // BTV: pow(spatial_decay, k), k = 0 .. 2R (host-computed, btv_regularizer.cpp:39)
  double* part_reg;            // per-tile cost slots (same indexing as the tile kernel's)
};

template <int R>
struct BtvDims {
  static constexpr int TH = 32, TW = FT_W, NT = 256;
  static constexpr int A = R - 1;                       // rows / columns above / left whose windows reach the tile
  static constexpr int XR = TH + A + R, XC = TW + A + R, XP = XC | 1;
  static constexpr int TR = TH + A, TC = TW + A, TP = TC | 1;
};

template <int R, bool BORDER>
__device__ __forceinline__ void btv_tile_body(const RegTileParams& P, double* __restrict__ xs, double* __restrict__ ts,
                                              int ch, int ty0, int tx0, double& cost_out) {
  using D = BtvDims<R>;
  const int tid = threadIdx.x;
  const size_t HW = (size_t)P.H * P.W;
  const double* __restrict__ xg = P.x + (size_t)ch * HW;
  const double* __restrict__ wg = P.w + (size_t)ch * HW;
  // ---- x tile + halo (zero outside the image: such taps are masked or multiply a zero t) ----------------
#pragma unroll 5
  for (int id = tid; id < D::XR * D::XC; id += D::NT) {
    const int r = id / D::XC, c = id - r * D::XC;
    const int gr = ty0 - D::A + r, gc = tx0 - D::A + c;
    double v = 0.0;
    if (!BORDER || (gr >= 0 && gc >= 0 && gr < P.H && gc < P.W)) v = xg[(size_t)gr * P.W + gc];
    xs[r * D::XP + c] = v;
  }
  __syncthreads();
  // ---- t_q = 2 lambda w_q r_q over the tile and the A rows / columns above / left of it; cost of the tile ----
  // thread = (column qc, segment of VL rows): the (R+1) x (R+1) window of x slides down the column in registers, so
  // a value costs R+1 shared-memory loads instead of (R+1)^2 - 1
  double cost = 0.0;
  constexpr int VSEG = D::NT / D::TC;                    // row segments that fit the CTA (3 for 66 columns)
  constexpr int VL = (D::TR + VSEG - 1) / VSEG;          // rows per segment
  static_assert(VSEG >= 1 && VSEG * VL >= D::TR, "value pass covers the region");
  if (tid < D::TC * VSEG) {
    const int qc = tid % D::TC, r0 = (tid / D::TC) * VL;
    const int gc = tx0 - D::A + qc;
    const int jmax = BORDER ? min(R, P.W - 1 - gc) : R;
    const bool col_in = !BORDER || (gc >= 0 && gc < P.W);
    double win[R + 1][R + 1];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j <= R; ++j) win[i][j] = (r0 + i < D::XR) ? xs[(r0 + i) * D::XP + qc + j] : 0.0;
#pragma unroll
    for (int l = 0; l < VL; ++l) {
      const int qr = r0 + l;
      if (qr < D::TR) {   // (qr + R < XR always: XR = TR + R)
#pragma unroll
        for (int j = 0; j <= R; ++j) win[R][j] = xs[(qr + R) * D::XP + qc + j];
        const int gr = ty0 - D::A + qr;
        double t = 0.0;
        if (col_in && (!BORDER || (gr >= 0 && gr < P.H))) {
          const int imax = BORDER ? min(R, P.H - 1 - gr) : R;
          const double x0 = win[0][0];
          double r = 0.0;
#pragma unroll
          for (int i = 0; i <= R; ++i)
#pragma unroll
            for (int j = 0; j <= R; ++j)
              if ((i | j) != 0 && (!BORDER || (i <= imax && j <= jmax)))
                r = fma(P.decay[i + j], fabs(x0 - win[i][j]), r);
          t = (P.two_lambda * wg[(size_t)gr * P.W + gc]) * r;
          if (qr >= D::A && qc >= D::A && (!BORDER || (gr >= P.row0 && gr < P.row1))) cost = fma(t, r, cost);
        }
        ts[qr * D::TP + qc] = t;
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j <= R; ++j) win[i][j] = win[i + 1][j];
    }
  }
  __syncthreads();
  cost_out = 0.5 * cost;  // lambda w r^2
  if (P.g == nullptr) return;
  // ---- partial derivatives: thread = (column, 8 rows) ------------------------------------------------------
  constexpr int EL = D::TH / (D::NT / D::TW);
  const int ec = tid % D::TW, er0 = (tid / D::TW) * EL;
  const int gc = tx0 + ec;
  double* __restrict__ gp = P.g + (size_t)ch * HW + (size_t)(ty0 + er0) * P.W + gc;
  double g_old[EL];  // the data-term gradient the tile kernel left: all loads in flight before the stencil work
#pragma unroll
  for (int l = 0; l < EL; ++l) {
    const int gr = ty0 + er0 + l;
    g_old[l] = (!BORDER || (gr < P.H && gc < P.W && gr >= P.row0 && gr < P.row1)) ? gp[(size_t)l * P.W] : 0.0;
  }
#pragma unroll
  for (int l = 0; l < EL; ++l) {
    const int gr = ty0 + er0 + l;
    if (BORDER && !(gr < P.H && gc < P.W && gr >= P.row0 && gr < P.row1)) continue;
    const double* __restrict__ xp = xs + (er0 + l + D::A) * D::XP + (ec + D::A);
    const double* __restrict__ tp = ts + (er0 + l + D::A) * D::TP + (ec + D::A);
    const double x0 = xp[0];
    double self = 0.0, nb = 0.0;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if ((i | j) == 0) continue;
        if (!BORDER || (gr + i < P.H && gc + j < P.W)) self += signed_by(x0 - xp[i * D::XP + j], P.decay[i + j]);
        if (!BORDER || !(gr - i == 0 && gc - j == 0))  // the reference skips image pixel (0, 0) here
          nb += signed_by(x0 - xp[-i * D::XP - j], tp[-i * D::TP - j] * P.decay[i + j]);
      }
    gp[(size_t)l * P.W] = g_old[l] + fma(tp[0], self, nb);
  }
}

// grid: (ceil(W / 64), units, 1), 256 threads
template <int R>
__global__ void __launch_bounds__(256)
k_btv_tile(const RegTileParams P) {
  using D = BtvDims<R>;
  __shared__ double xs[D::XR * D::XP];
  __shared__ double ts[D::TR * D::TP];
  const int unit = P.unit_begin + blockIdx.y;
  const int ch = unit / P.tile_rows;
  const int tx0 = blockIdx.x * D::TW, ty0 = (unit - ch * P.tile_rows) * D::TH;
  const bool interior = tx0 - D::A >= 0 && ty0 - D::A >= 0 && tx0 + D::TW + R <= P.W && ty0 + D::TH + R <= P.H &&
                        ty0 >= P.row0 && ty0 + D::TH <= P.row1;
  double cost = 0.0;
  if (interior) btv_tile_body<R, false>(P, xs, ts, ch, ty0, tx0, cost);
  else btv_tile_body<R, true>(P, xs, ts, ch, ty0, tx0, cost);
  cost = block_sum(cost);
  if (threadIdx.x == 0) P.part_reg[(size_t)unit * gridDim.x + blockIdx.x] = cost;
}

// ---- 3-D TV ---------------------------------------------------------------------------------------------
template <bool BORDER>
__device__ __forceinline__ void tv3d_tile_body(const RegTileParams& P, int ch, int ty0, int tx0, double& cost_out) {
  constexpr int TH = 32, TW = FT_W, NT = 256, EL = TH / (NT / TW);
  const int tid = threadIdx.x;
  const int ec = tid % TW, er0 = (tid / TW) * EL;
  const int gc = tx0 + ec, gr0 = ty0 + er0;
  cost_out = 0.0;
  if (BORDER && (gc >= P.W || gr0 >= P.H)) return;
  const size_t HW = (size_t)P.H * P.W;
  const size_t W = (size_t)P.W;
  const bool has_next = ch + 1 < P.Ca, has_prev = ch > 0;
  const bool has_r = !BORDER || gc + 1 < P.W, has_l = !BORDER || gc > 0;
  const double* __restrict__ xc = P.x + (size_t)ch * HW + (size_t)gr0 * W + gc;
  const double* __restrict__ xn = xc + HW;   // channel c + 1 (dereferenced only when has_next)
  const double* __restrict__ xv = xc - HW;   // channel c - 1 (only when has_prev)
  const double* __restrict__ wc = P.w + (size_t)ch * HW + (size_t)gr0 * W + gc;
  const double* __restrict__ wv = wc - HW;
  const double tl2 = P.two_lambda;
  // the pixel above the first row of this thread's segment: its d/dy term
  double b_above = 0.0;
  if (gr0 > 0) {
    const double xa = xc[-(long long)W], x0 = xc[0];
    const double gya = x0 - xa;
    const double gxa = has_r ? xc[-(long long)W + 1] - xa : 0.0;
    double ra = fabs(gya) + fabs(gxa);
    if (has_next) ra += fabs(xn[-(long long)W] - xa);
    b_above = signed_by(gya, (tl2 * wc[-(long long)W]) * ra);
  }
  double x0 = xc[0], xl = has_l ? xc[-1] : 0.0;
  double xv0 = has_prev ? xv[0] : 0.0;
  double cost = 0.0;
  double* __restrict__ gp = P.g ? P.g + (size_t)ch * HW + (size_t)gr0 * W + gc : nullptr;
#pragma unroll 2
  for (int l = 0; l < EL; ++l) {
    const int gr = gr0 + l;
    if (BORDER && gr >= P.H) break;
    const bool has_b = !BORDER || gr + 1 < P.H;
    const size_t o = (size_t)l * W;
    const double xr = has_r ? xc[o + 1] : 0.0;
    const double xb = has_b ? xc[o + W] : 0.0;
    const double xbl = (has_b && has_l) ? xc[o + W - 1] : 0.0;
    // own value and the self term (no z part in the derivative: tv_regularizer.cpp:154-170)
    const double gx = has_r ? xr - x0 : 0.0;
    const double gy = has_b ? xb - x0 : 0.0;
    double r = fabs(gy) + fabs(gx);
    if (has_next) r += fabs(xn[o] - x0);
    const double t = (tl2 * wc[o]) * r;
    const double a_own = signed_by(gx, t), b_own = signed_by(gy, t);
    // left neighbour (tv_regularizer.cpp:171-184)
    double a_left = 0.0;
    if (has_l) {
      const double gxl = x0 - xl;
      const double gyl = has_b ? xbl - xl : 0.0;
      double rl = fabs(gyl) + fabs(gxl);
      if (has_next) rl += fabs(xn[o - 1] - xl);
      a_left = signed_by(gxl, (tl2 * wc[o - 1]) * rl);
    }
    // previous channel (tv_regularizer.cpp:202-220)
    double z_prev = 0.0;
    double xvb = 0.0;
    if (has_prev) {
      const double gzb = x0 - xv0;
      const double gxb = has_r ? xv[o + 1] - xv0 : 0.0;
      xvb = has_b ? xv[o + W] : 0.0;
      const double gyb = has_b ? xvb - xv0 : 0.0;
      const double rb = (fabs(gyb) + fabs(gxb)) + fabs(gzb);
      z_prev = signed_by(gzb, (tl2 * wv[o]) * rb);
    }
    if (!BORDER || (gr >= P.row0 && gr < P.row1)) {
      if (gp) gp[o] += ((a_left + b_above) + z_prev) - (a_own + b_own);
      cost = fma(t, r, cost);
    }
    b_above = b_own;
    x0 = xb;
    xl = xbl;
    xv0 = xvb;
  }
  cost_out = 0.5 * cost;
}

__global__ void __launch_bounds__(256, 4)
k_tv3d_tile(const RegTileParams P) {
  const int unit = P.unit_begin + blockIdx.y;
  const int ch = unit / P.tile_rows;
  const int tx0 = blockIdx.x * FT_W, ty0 = (unit - ch * P.tile_rows) * 32;
  const bool interior = tx0 > 0 && tx0 + FT_W < P.W && ty0 + 32 < P.H && ty0 >= P.row0 && ty0 + 32 <= P.row1;
  double cost = 0.0;
  if (interior) tv3d_tile_body<false>(P, ch, ty0, tx0, cost);
  else tv3d_tile_body<true>(P, ch, ty0, tx0, cost);
  cost = block_sum(cost);
  if (threadIdx.x == 0) P.part_reg[(size_t)unit * gridDim.x + blockIdx.x] = cost;
}

}  // namespace srb
