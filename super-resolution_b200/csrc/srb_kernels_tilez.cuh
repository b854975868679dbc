// srb_kernels_tilez.cuh -- k_tile_zt: the fused tile kernel for models with integer shifts and at most
// one DISTINCT shift per sub-pixel phase (cfg1 .. cfg5 and every frame shard of them), observations in the
// TRANSPOSED Z layout.  Same mathematics as k_tile (srb_kernels_tile.cuh):
//     r_k = D M_k (B x) - y_k,     g = 2 s^2 B^T ( sum_k M_k^T D^T r_k ) + TV part,
// (objective_data_term.cpp:15-75, tv_regularizer.cpp:134-227), rearranged around what bounds the tile
// on a B200 -- the shared-memory pipe, not HBM and not the fp64 pipe (profiles/r02_k_tile_z_full.txt):
//
//  * Frames with the SAME shift (cfg4: 2 per phase, cfg5: 4 per phase) have the same operator A, so
//        sum_e ||A x - y_e||^2 = n ||A x - m||^2 + sum_e ||y_e - m||^2,      m = mean_e y_e,
//    exactly, with no cancellation (both terms are sums of squares).  The observations are constant during
//    a solve: m and the constant second term are formed ONCE, when the observations are uploaded
//    (k_build_yzt), and the kernel below runs on m with the factor n folded into its two scale factors.
//    The n frames then cost one observation load per HR pixel instead of n.
//
//  * The observations are gathered once, at upload, onto the HR grid ("Z layout": yzt(c, p) = the one
//    regular LR sample that lands on HR pixel p), stored COLUMN-MAJOR per channel, padded by the PSF half
//    width on every side and rounded up to whole tiles, with NaN wherever no regular sample lands (sub-
//    pixel phases without a frame: a frame shard; samples of the border band; positions outside the LR
//    image).  Every tile -- interior or on the image border -- then runs the same code: Z = Bx - yzt
//    where yzt is a number, 0 elsewhere; there is no per-sample table walk left in this kernel.
//  * Column-major because the residual is consumed by the ROW pass: thread (row r, segment) walks along
//    the row, and with the box stored [column][row] the 16 threads of a half-warp (consecutive rows) read
//    consecutive shared-memory words.  (A TMA box is dense and its pitch therefore even; a row-major box
//    would put a half-warp's 16 rows on 8 or fewer distinct banks.)
//  * Z = Bx - yzt is never written: the adjoint horizontal pass forms it on the fly from the Bx row
//    (in place, in registers) and the yzt column box, and counts the cost of the positions its thread
//    owns.  Per Z-region row that is 252 shared-memory accesses instead of 376 (K = 7).
//  * Row passes split a row into segments of L columns with L chosen so that (row pitch) * (rows) == L
//    (mod 16): stepping from the last row of one segment to the first row of the next then looks like one
//    more row step to the bank mapping, and half-warps that straddle two segments stay conflict-free
//    (15 % of all shared-memory wavefronts of k_tile_z were such replays).
//
// Phases (A, B = the two shared-memory buffers; 4 CTAs / SM for K <= 7):
//     0  TMA: x tile + halo -> A, IRLS weights -> B                      (zero fill outside the image)
//     1  2-D TV gradient + cost -> registers                              (tile_tv, as k_tile)
//     2a vertical PSF pass   A -> B
//        TMA: yzt box -> A   (lands while 2b runs)
//     2b horizontal PSF pass B -> B in place
//     3  adjoint horizontal pass on Z = Bx - yzt, B -> B in place; cost
//     4  adjoint vertical pass + TV part -> g                             (one coalesced store)
#pragma once
#include "srb_kernels_tile.cuh"

namespace srb {

template <int KH>
struct ZtDims {
  using D = TileDims<KH, false, 32>;
  static constexpr int TH = 32;
  static constexpr int NT = 256;
  static constexpr int K = 2 * KH + 1;
  static constexpr int HYR = (KH + 1) & ~1;   // even row halo of the yzt box (FLOAT64 TMA: even inner start)
  static constexpr int YR = TH + 2 * HYR;     // yzt box: rows (inner, contiguous)
  static constexpr int YC = FT_W + 2 * KH;    // yzt box: columns
  static constexpr int TR = TH + 2 * KH;      // rows of the Z region = rows of the vertical-pass output
  static constexpr int TW = FT_W + 4 * KH;    // columns of the vertical-pass output
  static constexpr int BW = FT_W + 2 * KH;    // columns of Bx = columns of the Z region
  // row pitch and segment length of the row passes: TP odd, TP * TR == L (mod 16) where possible
  static constexpr int TP = KH == 1 ? 69 : KH == 2 ? 75 : KH == 3 ? 77 : 81;
  static constexpr int L = KH == 1 ? 10 : KH == 2 ? 12 : KH == 3 ? 14 : 12;
  static constexpr int NSEG_F = (BW + L - 1) / L;    // forward horizontal pass
  static constexpr int NSEG_A = (FT_W + L - 1) / L;  // adjoint horizontal pass
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int A_DOUBLES = (cmax(D::XH * D::XW, YC * YR) + 15) & ~15;  // x tile | yzt box
  static constexpr int B_DOUBLES = (cmax(TR * TP, D::WH * D::WW) + 15) & ~15;  // weights | tmp -> Bx -> t2
  static constexpr size_t SMEM_BYTES = (size_t)(A_DOUBLES + B_DOUBLES) * sizeof(double) + 16;
  static constexpr unsigned Y_BYTES = YR * YC * sizeof(double);
  static_assert(TP >= TW && (TP & 1) == 1, "odd row pitch");
  static_assert(TR * NSEG_F <= NT && TR * NSEG_A <= NT, "one row segment per thread");
  static_assert(L >= 2 * KH && NSEG_F * L >= BW && NSEG_A * L >= FT_W, "segment geometry");
};

// Padded extent of the transposed Z layout along one axis: [-pad, tiles * tile) + pad.
__host__ __device__ inline int zt_rows_padded(int H, int KH) { return (H + 31) / 32 * 32 + 2 * ((KH + 1) & ~1); }
__host__ __device__ inline int zt_cols_padded(int W, int KH) { return (W + FT_W - 1) / FT_W * FT_W + 2 * KH; }

template <int KH>
__global__ void __launch_bounds__(256, KH <= 3 ? 4 : 3)
k_tile_zt(const TileParams P, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
          const __grid_constant__ CUtensorMap map_y) {
  using Z = ZtDims<KH>;
  using D = typename Z::D;
  constexpr int K = Z::K, NT = Z::NT, TH = Z::TH, TP = Z::TP, L = Z::L;
  static_assert(KH >= 1 && KH <= 4, "PSF 3x3 .. 9x9");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* bufA = reinterpret_cast<double*>(smem_raw);  // xs [XH][XW] -> yzt [YC][YR]
  double* bufB = bufA + Z::A_DOUBLES;                  // ws [WH][WW] -> tmp -> bx -> t2, all [TR][TP]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(bufB + Z::B_DOUBLES);
  const int tid = threadIdx.x;
  const int unit = P.unit_begin + blockIdx.y;
  const int ch = unit / P.tile_rows;
  const int tx0 = blockIdx.x * FT_W, ty0 = (unit - ch * P.tile_rows) * TH;
  const size_t HW = (size_t)P.H * P.W;

  // ---- 0. x tile + halo and IRLS weights by TMA -------------------------------------------------
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, D::X_BYTES + (P.reg_fused ? D::W_BYTES : 0u));
    tma_load_3d(bufA, &map_x, tx0 - D::HXC, ty0 - D::HX, ch, bar);
    if (P.reg_fused) tma_load_3d(bufB, &map_w, tx0 - 2, ty0 - 1, ch, bar);
  }
  __syncthreads();  // mbarrier initialised before anyone waits on it
  mbar_wait_bounded(bar, 0);

  // ---- 1. 2-D TV (tv_regularizer.cpp:134-227; see tile_tv) -----------------------------------------
  constexpr int ESEG = NT / FT_W;  // 4
  constexpr int EL = TH / ESEG;    // 8
  const int ec = tid % FT_W, er0 = (tid / FT_W) * EL;
  const int gc = tx0 + ec;
  double tvg[EL];
  double cost_reg = 0.0;
#pragma unroll
  for (int l = 0; l < EL; ++l) tvg[l] = 0.0;
  if (P.reg_fused) {
    const bool tile_inside = tx0 + FT_W < P.W && ty0 + TH < P.H;  // strictly: right / bottom neighbours exist
    if (tile_inside && ty0 >= P.row0 && ty0 + TH <= P.row1)
      tile_tv<KH, false, TH, EL, false>(P, bufA, bufB, ty0, gc, ec, er0, tvg, cost_reg);
    else
      tile_tv<KH, false, TH, EL, true>(P, bufA, bufB, ty0, gc, ec, er0, tvg, cost_reg);
    __syncthreads();  // ws (B) is overwritten by the vertical pass
  }

  // ---- 2a. vertical PSF pass: tmp[r][c] = sum_i u[i] * xs[r+i][c]   (A -> B) ------------------------
  {
    constexpr int NSEG = (NT / Z::TW) > 0 ? (NT / Z::TW) : 1;
    constexpr int LV = (Z::TR + NSEG - 1) / NSEG;
    const double* __restrict__ xo = bufA + D::XOR * D::XW + D::XOC;
    for (int id = tid; id < Z::TW * NSEG; id += NT) {
      const int c = id % Z::TW, seg = id / Z::TW;
      const int r0 = seg * LV;
      const double* __restrict__ src = xo + r0 * D::XW + c;
      double* __restrict__ dst = bufB + r0 * TP + c;
      slide_correlate<K, LV, SRB_SLIDE_B>(P.u, Z::TR - r0,
                                [&](int i) { return src[i * D::XW]; },
                                [&](int l, double v) { dst[l * TP] = v; });
    }
  }
  __syncthreads();

  // x is consumed: the observations of the Z region land in A while the horizontal pass runs
  if (tid == 0) {
    fence_proxy_async();
    mbar_expect_tx(bar, Z::Y_BYTES);
    tma_load_3d(bufA, &map_y, ty0, tx0, P.c0 + ch, bar);  // padded coordinates: (row + HYR, column + KH)
  }

  // ---- 2b. horizontal PSF pass, in place: bx[r][c] = sum_j v[j] * tmp[r][c+j]   (B -> B) ----------
  // thread = (row, segment of L columns); the K-1 inputs beyond its segment belong to the next thread's
  // outputs and are read before the barrier
  const int rr = tid % Z::TR, seg = tid / Z::TR;
  const int c0 = seg * L;
  double* __restrict__ row = bufB + rr * TP + c0;
  {
    const bool active = seg < Z::NSEG_F;
    double tail[K - 1];
#pragma unroll
    for (int i = 0; i < K - 1; ++i) tail[i] = (active && c0 + L + i < Z::TW) ? row[L + i] : 0.0;
    __syncthreads();
    if (active)
      slide_correlate<K, L, SRB_SLIDE_B>(P.v, Z::BW - c0,
                                [&](int j) { return j < L ? row[j < L ? j : 0] : tail[j < L ? 0 : j - L]; },
                                [&](int l, double v) { row[l] = v; });
  }
  __syncthreads();

  // ---- 3. adjoint horizontal pass on the residuals, in place:
  //         Z[r][c] = bx[r][c] - yzt[c][r] (0 where yzt is NaN);  t2[r][c] = sum_j u[j] * Z[r][c+j];
  //         cost += Z^2 over the positions this tile owns (its own pixels, plus the part of its halo that
  //         lies outside the image) ------------------------------------------------------------------
  mbar_wait_bounded(bar, 1);
  double cost_data = 0.0;
  {
    const bool active = seg < Z::NSEG_A;
    double tail[2 * KH];  // bx [c0+L, c0+L+2KH): overwritten by the next segment's outputs
#pragma unroll
    for (int i = 0; i < 2 * KH; ++i) tail[i] = (active && c0 + L + i < Z::BW) ? row[L + i] : 0.0;
    __syncthreads();
    if (active) {
      // positions (rr, c0 + j) whose cost this thread counts: j in [jlo, jhi)
      const bool first_col = tx0 == 0, last_col = tx0 + FT_W >= P.W;
      const bool first_row = ty0 == 0, last_row = ty0 + TH >= P.H;
      const bool own_r = (rr >= KH && rr < KH + TH) || (rr < KH && first_row) || (rr >= KH + TH && last_row);
      int jlo = seg == 0 ? (first_col ? 0 : KH) : KH;               // thread owns j in [KH, KH+L) (+ the left halo in segment 0)
      int jhi = KH + L;
      const int c_end = last_col ? Z::BW : KH + FT_W;                // tile owns columns [first_col ? 0 : KH, c_end)
      if (c0 + jhi > c_end) jhi = c_end - c0;
      if (!own_r) jhi = jlo;
      const double* __restrict__ yp = bufA + c0 * Z::YR + (rr + Z::HYR - KH);
      double cost = 0.0;
      slide_correlate<K, L, SRB_SLIDE_B>(P.u, FT_W - c0,
                                [&](int j) {
                                  const double b = j < L ? row[j < L ? j : 0] : tail[j < L ? 0 : j - L];
                                  const double yv = yp[j * Z::YR];
                                  const double res = (yv == yv) ? b - yv : 0.0;
                                  if (j >= jlo && j < jhi) cost = fma(res, res, cost);
                                  return res;
                                },
                                [&](int l, double v) { row[l] = v; });
      cost_data = cost;
    }
  }
  __syncthreads();

  // ---- 4. adjoint vertical pass + regularizer part + store -----------------------------------------
  if (P.g != nullptr) {
    const double* __restrict__ t2 = bufB;
    double win[K];
#pragma unroll
    for (int i = 0; i < K - 1; ++i) win[i] = t2[(er0 + i) * TP + ec];
    double* __restrict__ gp = P.g + (size_t)ch * HW + (size_t)(ty0 + er0) * P.W + gc;
    const size_t gstep = (size_t)P.W;
    const bool all_in = tx0 + FT_W <= P.W && ty0 + TH <= P.H;
#pragma unroll
    for (int l = 0; l < EL; ++l) {
      win[K - 1] = t2[(er0 + l + K - 1) * TP + ec];
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < K; ++i) acc = fma(P.v[i], win[i], acc);
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
      if (all_in || (ty0 + er0 + l < P.H && gc < P.W)) *gp = fma(P.two_s2, acc, tvg[l]);
      gp += gstep;
    }
  }

  // ---- cost partial sums (deterministic: fixed per-CTA slot, fixed-order final reduction) --------
  const size_t cta = (size_t)unit * gridDim.x + blockIdx.x;
  block_sum2<NT>(cost_data, cost_reg);
  if (tid == 0) {
    // P.s2 = n s^2 with n frames merged per phase; the constant part of the merged data cost is added once per
    // channel, by the channel's first tile
    const double merged_const = (P.yvar != nullptr && blockIdx.x == 0 && ty0 == 0) ? P.yvar[P.c0 + ch] : 0.0;
    P.part_data[cta] = P.s2 * cost_data + merged_const;
    P.part_reg[cta] = cost_reg;
  }
}

// Builds the transposed, padded Z layout from the LR observations (once per srb_set_observations):
//   yzt[(c * cols_p + pc + KH) * rows_p + pr + HYR] = mean over the n frames (k, q) of the sub-pixel phase whose
//   regular sample lands on HR position (pr, pc) -- positions up to the PSF half width outside the image
//   included -- else NaN.  All frames of a phase have the same shift (planner: zt), hence the same q; n = 1 is
//   the plain copy.  var_part[c * var_stride + block] = sum over the block of sum_e (y_e - mean)^2 (0 for n = 1).
// grid: (ceil(rows_p / 256), cols_p, Ct)
__global__ void __launch_bounds__(256)
k_build_yzt(int h, int w, int s, int rows_p, int cols_p, int pad_r, int pad_c, int lo_r, int hi_r, int lo_c,
            int hi_c, const TEntry* __restrict__ entries, const int* __restrict__ phase_begin,
            const double* __restrict__ y, double* __restrict__ yzt, double* __restrict__ var_part, size_t var_stride) {
  const int rp = blockIdx.x * 256 + threadIdx.x, cp = blockIdx.y, c = blockIdx.z;
  double var = 0.0;
  if (rp < rows_p) {
    const int pr = rp - pad_r, pc = cp - pad_c;
    const int mr = floordiv(pr, s), mc = floordiv(pc, s);
    const int ph = (pr - mr * s) * s + (pc - mc * s);
    double v = __longlong_as_double(0x7ff8000000000000LL);
    const int e0 = phase_begin[ph], cnt = phase_begin[ph + 1] - e0;
    if (cnt > 0) {
      const TEntry e = entries[e0];
      const int qr = mr + (int)(short)(e.qoff & 0xffff), qc = mc + (e.qoff >> 16);
      if (qr >= lo_r && qr < hi_r && qc >= lo_c && qc < hi_c) {
        const double* __restrict__ yc = y + (size_t)c * ((size_t)h * w) + ((long long)mr * w + mc);
        if (cnt == 1) {
          v = yc[e.yoff];
        } else {
          double sum = 0.0;
          for (int i = 0; i < cnt; ++i) sum += yc[entries[e0 + i].yoff];
          v = sum / (double)cnt;
          for (int i = 0; i < cnt; ++i) {
            const double d = yc[entries[e0 + i].yoff] - v;
            var = fma(d, d, var);
          }
        }
      }
    }
    yzt[((size_t)c * cols_p + cp) * rows_p + rp] = v;
  }
  if (var_part != nullptr) {
    var = block_sum(var);
    if (threadIdx.x == 0) var_part[(size_t)c * var_stride + (size_t)blockIdx.y * gridDim.x + blockIdx.x] = var;
  }
}

// yvar[c] = scale * sum(var_part[c][0 .. per_channel)), fixed order (deterministic).  grid: Ct
__global__ void __launch_bounds__(1024)
k_reduce_yvar(const double* __restrict__ var_part, size_t per_channel, double scale, double* __restrict__ yvar) {
  const double* __restrict__ p = var_part + (size_t)blockIdx.x * per_channel;
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < per_channel; i += blockDim.x) acc += p[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) yvar[blockIdx.x] = scale * acc;
}

}  // namespace srb
