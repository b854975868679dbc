// srb_kernels_tile.cuh -- the fused tile kernel of the MAP objective (the fast path) and its
// border-band companions.
//
// Algorithm.  The reference evaluates, frame by frame,
//     r_k = D B M_k x - y_k,      g += 2 s^2 M_k^T B^T D^T r_k          (objective_data_term.cpp:15-75)
// with B (PSF correlation) applied at full HR resolution twice per frame.  B and the translation
// M_k are both convolutions, so for every LR sample whose PSF window does not interact with the
// image border ("regular" samples) they commute:
//     r_k = D M_k (B x) - y_k,    g = 2 s^2 B^T ( sum_k M_k^T D^T r_k ).
// One CTA owns one 32 x 64 HR tile of one channel and does, entirely in shared memory:
//     0. TMA (cp.async.bulk.tensor, zero fill outside the image) of the x tile + halo and of the
//        IRLS-weight tile; one mbarrier                                  HBM -> smem, x read ONCE
//     1. IRLS-weighted 2-D TV gradient + cost of the tile from x and w   -> registers
//     2. Bx = separable PSF correlation of the tile                      (2 passes, sliding windows)
//     3. Z(p) = sum over the regular LR samples (k,q) landing on HR pixel p of (Bx-sample - y_k(q));
//        every LR observation is read exactly once; cost += r^2
//     4. g = 2 s^2 B^T Z + TV part, single coalesced store of g          (2 passes)
// The PSF work no longer scales with the number of frames, and per frame an HR pixel costs one LR
// load.  The few samples near the image border whose window is clipped differently before and
// after the shift ("special" samples: a frame-independent band of LR rows / columns) are left out
// here and evaluated in the reference's operation order by k_band_forward / k_band_adjoint, so the
// sum is exact for every integer or fractional shift.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "srb_common.cuh"
#include "srb_kernels_generic.cuh"

namespace srb {

constexpr int FT_H = 32;    // tile rows
constexpr int FT_W = 64;    // tile columns
constexpr int FT_NT = 256;  // threads per CTA
constexpr int FT_MAX_ENTRIES = 4096;  // 128 KB of shared memory
constexpr int FT_MAX_SCALE = 8;
constexpr int SRB_MAX_PEERS = 8;
#ifndef SRB_SLIDE_B
#define SRB_SLIDE_B 2  // inputs fetched ahead per batch in the 1-D PSF passes (1: 0.1403, 2: 0.1374, 4: 0.1457 ms at cfg3)
#endif

// One way a regular LR sample lands on an HR pixel of a given sub-pixel phase.
struct __align__(16) TEntry {
  long long yoff;  // k*Ct*h*w + qoff_r*w + qoff_c : observation offset relative to cell (mr, mc)
  int bxoff;       // Bx element sampled, relative to the Bx element under the Z position
  int qoff;        // (qoff_r & 0xffff) | (qoff_c << 16): LR index = floor(p / s) + qoff
  double wT;       // transpose-warp bilinear weight of this tap          (FRAC only)
  short fy, fx;    // forward-warp bilinear fractions in 1/32 px          (FRAC only)
  int owner;       // 1 for the (0,0) transpose tap: counts the sample's cost
};
static_assert(sizeof(TEntry) == 32, "TEntry layout");

// Precomputed work item of the residual pass for interior tiles of models where every sub-pixel
// phase holds exactly one (frame, tap) entry and the scale divides the tile size: everything that
// does not depend on the tile (observation offset relative to the tile's first LR cell, Bx and Z
// element) is resolved on the host.  Items [0, FT_W*s*(TH/32)) are the pass-A columns (first row
// of the item), the rest the halo-ring pixels.
struct __align__(16) TFast {
  long long yrel;  // + (ty0/s)*w + tx0/s = observation index within the channel
  int bxo;         // Bx element
  int zo;          // Z element
};

struct TileParams {
  int H, W, h, w, s, sshift, Ct, c0, Ca;  // sshift = log2(s) when s is a power of two, else -1
  const double* x;
  const double* y;
  double* g;           // may be NULL (cost only)
  const double* wts;   // IRLS weights
  const TEntry* entries;
  const int* phase_begin;  // [s*s + 1]
  int num_entries;
  int qoff_min_r, qoff_max_r, qoff_min_c, qoff_max_c;
  int lo_r, hi_r, lo_c, hi_c;  // regular LR samples: lo_r <= qr < hi_r, lo_c <= qc < hi_c
  double u[9], v[9];   // psf[i][j] = u[i] * v[j]
  double two_s2;       // 2 * s^2
  double s2;           // s^2
  double two_lambda;   // 2 * lambda
  int reg_fused;       // 1: 2-D TV term evaluated here
  int row0, row1;      // HR row band of the regularizer term on this rank
  int use_tma;
  const TFast* fast;   // NULL: generic residual pass only
  int unit_begin, tile_rows;  // first (channel, tile row) unit of this launch; tile rows per channel
  double* part_data;   // per-CTA partial sums of the data cost
  double* part_reg;    // per-CTA partial sums of the regularization cost
};

template <int KH, bool FRAC, int TH>
struct TileDims {
  static constexpr int NT = TH * (FT_W / 8);      // threads per CTA: 8 pixels per thread
  static constexpr int K = 2 * KH + 1;
  static constexpr int HB = KH + (FRAC ? 1 : 0);  // halo of Bx around the tile
  static constexpr int HX = (KH + HB) > 0 ? (KH + HB) : 1;  // row halo of x around the tile (TV needs 1)
  // column halo: a FLOAT64 TMA box must start on an even column (16-byte aligned global address;
  // measured on B200 with tools/tma_probe.cu: odd start coordinates raise "illegal instruction")
  static constexpr int HXC = (HX + 1) & ~1;
  static constexpr int XOR = HX - (KH + HB);      // xs rows the PSF passes skip
  static constexpr int XOC = HXC - (KH + HB);     // xs columns the PSF passes skip
  static constexpr int XH = TH + 2 * HX, XW = FT_W + 2 * HXC;  // x tile, dense (TMA box)
  static constexpr int TW = FT_W + 2 * (KH + HB);                // vertical-pass output width
  static constexpr int TR = TH + 2 * HB, TP = TW | 1;          // vertical-pass output
  static constexpr int BW = FT_W + 2 * HB, BP = BW | 1;          // Bx
  static constexpr int ZH = TH + 2 * KH, ZW = FT_W + 2 * KH, ZP = ZW | 1;  // Z
  static constexpr int T2P = FT_W | 1;                           // adjoint horizontal pass
  static constexpr int WH = TH + 1, WW = FT_W + 2;             // IRLS weights, dense (TMA box)
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int A_DOUBLES = (cmax(cmax(XH * XW, TR * BP), ZH * T2P) + 15) & ~15;  // xs | bx | t2
  static constexpr int B_DOUBLES = (cmax(cmax(TR * TP, ZH * ZP), WH * WW) + 15) & ~15;   // ws | tmp | z
  static constexpr size_t PB_BYTES = (FT_MAX_SCALE * FT_MAX_SCALE + 1 + 3) / 4 * 16;
  // [A][B][phase table][mbarrier] then the entry list (32 B per entry, sized at launch)
  static constexpr size_t FIXED_BYTES = (size_t)(A_DOUBLES + B_DOUBLES) * sizeof(double) + PB_BYTES + 16;
  static constexpr size_t smem_bytes(int num_entries) { return FIXED_BYTES + (size_t)num_entries * sizeof(TEntry); }
  static constexpr unsigned X_BYTES = XH * XW * sizeof(double);
  static constexpr unsigned W_BYTES = WH * WW * sizeof(double);
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS) ------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done = 0;
  const unsigned a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(smem_u32(bar))
      : "memory");
}

// sgn(g) * t with sgn(0) = 0 (tv_regularizer.cpp:152-201: the three-way branches on the sign of a
// forward difference), on the integer pipes: the fp64 pipe is the scarce one in this kernel.
__device__ __forceinline__ double signed_by(double g, double t) {
  const int ghi = __double2hiint(g), glo = __double2loint(g);
  const bool nz = ((ghi & 0x7fffffff) | glo) != 0;
  const int rhi = __double2hiint(t) ^ (ghi & 0x80000000);
  return nz ? __hiloint2double(rhi, __double2loint(t)) : 0.0;
}

// Sums two values over the thread block; results valid in thread 0.
template <int NT>
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double part[2][32];
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    part[0][wid] = a;
    part[1][wid] = b;
  }
  __syncthreads();
  if (wid == 0) {
    a = (lane < NT / 32) ? part[0][lane] : 0.0;
    b = (lane < NT / 32) ? part[1][lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// floor(a / s) for the downsampling scale: a shift when s is a power of two (sshift >= 0).
__device__ __forceinline__ int floordiv_scale(int a, int s, int sshift) {
  return sshift >= 0 ? (a >> sshift) : floordiv(a, s);
}

// Residuals of the regular LR samples that land on ONE HR pixel of the Z region (row r, column c):
// z = sum over the pixel's (frame, tap) entries [e0, e1) of wT * (Bx sample - observation).
//   EDGE: the tile touches the border band -- samples are range / band checked.
template <int KH, bool FRAC, int TH, bool EDGE>
__device__ __forceinline__ double pixel_residuals(const TileParams& P, const double* __restrict__ bx,
                                                  const TEntry* __restrict__ ents, int e0, int e1,
                                                  const double* __restrict__ ycell, int mr, int mc, int r,
                                                  int c, bool own, double& cost) {
  using D = TileDims<KH, FRAC, TH>;
  double z = 0.0;
#pragma unroll 1
  for (int e = e0; e < e1; ++e) {
    const int4 head = *reinterpret_cast<const int4*>(ents + e);  // yoff, bxoff, qoff
    const long long yoff = ((long long)head.y << 32) | (unsigned)head.x;
    if (EDGE) {
      const int qr = mr + (int)(short)(head.w & 0xffff), qc = mc + (head.w >> 16);
      if (qr < P.lo_r || qr >= P.hi_r || qc < P.lo_c || qc >= P.hi_c) continue;
    }
    const double obs = __ldg(ycell + yoff);
    const double* b = bx + r * D::BP + c + head.z;
    if (!FRAC) {
      const double res = b[0] - obs;
      z += res;
      if (own) cost = fma(res, res, cost);
    } else {
      const TEntry en = ents[e];
      const double wy1 = en.fy * (1.0 / 32.0), wy0 = 1.0 - wy1;
      const double wx1 = en.fx * (1.0 / 32.0), wx0 = 1.0 - wx1;
      const double pred = b[0] * (wy0 * wx0) + b[1] * (wy0 * wx1) + b[D::BP] * (wy1 * wx0) +
                          b[D::BP + 1] * (wy1 * wx1);
      const double res = pred - obs;
      z = fma(en.wT, res, z);
      if (own && en.owner) cost = fma(res, res, cost);
    }
  }
  return z;
}

// 1-D sliding correlation of one thread's segment: out(l) = sum_i coef[i] * in(l + i) for
// l in [0, L), only the first `nvalid` outputs being stored.  Inputs are fetched B at a time ahead of
// the FMAs that use them, so that one shared-memory latency covers B outputs instead of one.
template <int K, int L, int B, class In, class Out>
__device__ __forceinline__ void slide_correlate(const double* __restrict__ coef, int nvalid, In in, Out out) {
  double win[K - 1 + B];
#pragma unroll
  for (int i = 0; i < K - 1; ++i) win[i] = in(i);
#pragma unroll
  for (int l0 = 0; l0 < L; l0 += B) {
#pragma unroll
    for (int b = 0; b < B; ++b) win[K - 1 + b] = (l0 + b < L && l0 + b < nvalid) ? in(l0 + b + K - 1) : 0.0;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      if (l0 + b < L && l0 + b < nvalid) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) acc = fma(coef[i], win[b + i], acc);
        out(l0 + b, acc);
      }
    }
#pragma unroll
    for (int i = 0; i < K - 1; ++i) win[i] = win[i + B];
  }
}

// IRLS-weighted 2-D TV gradient + cost of this thread's EL pixels (column ec, rows er0..).
//   BORDER: the tile touches the right / bottom image border or the edge of the regularizer row
//   band, so neighbours and outputs are checked per pixel.
template <int KH, bool FRAC, int TH, int EL, bool BORDER>
__device__ __forceinline__ void tile_tv(const TileParams& P, const double* __restrict__ xs,
                                        const double* __restrict__ ws, int ty0, int gc, int ec, int er0,
                                        double (&tvg)[EL], double& cost_reg) {
  using D = TileDims<KH, FRAC, TH>;
  const bool has_r = !BORDER || gc + 1 < P.W;
  const double* __restrict__ xp = xs + (er0 + D::HX) * D::XW + (ec + D::HXC);
  const double* __restrict__ wp = ws + (er0 + 1) * D::WW + (ec + 2);
  const double tl2 = P.two_lambda;
  double b_above;
  {  // B of the pixel above the first row of this thread's segment
    const double xa = xp[-D::XW], x0 = xp[0];
    const double gya = x0 - xa;
    const double gxa = has_r ? xp[-D::XW + 1] - xa : 0.0;
    const double ta = (tl2 * wp[-D::WW]) * (fabs(gya) + fabs(gxa));
    b_above = signed_by(gya, ta);
  }
  double x0 = xp[0], xl = xp[-1];
  double cost = 0.0;
#pragma unroll
  for (int l = 0; l < EL; ++l) {
    const int gr = ty0 + er0 + l;
    const bool has_b = !BORDER || gr + 1 < P.H;
    const double xr = xp[l * D::XW + 1];
    const double xb = xp[(l + 1) * D::XW], xbl = xp[(l + 1) * D::XW - 1];
    const double gx = has_r ? xr - x0 : 0.0;
    const double gy = has_b ? xb - x0 : 0.0;
    const double r = fabs(gy) + fabs(gx);
    const double t = (tl2 * wp[l * D::WW]) * r;
    const double a_own = signed_by(gx, t), b_own = signed_by(gy, t);
    const double gxl = x0 - xl;
    const double gyl = has_b ? xbl - xl : 0.0;
    const double tl = (tl2 * wp[l * D::WW - 1]) * (fabs(gyl) + fabs(gxl));
    const double a_left = signed_by(gxl, tl);
    if (!BORDER || (gr >= P.row0 && gr < P.row1 && gr < P.H && gc < P.W)) {
      tvg[l] = (a_left + b_above) - (a_own + b_own);
      cost = fma(t, r, cost);  // 2 lambda w r^2; halved below
    }
    b_above = b_own;
    x0 = xb;
    xl = xbl;
  }
  cost_reg = 0.5 * cost;
}

template <int KH, bool FRAC, int TH>
__global__ void __launch_bounds__(TH * (FT_W / 8), TH == 32 ? (KH <= 3 ? 4 : 3) : 2)
k_tile(const TileParams P, const __grid_constant__ CUtensorMap map_x,
       const __grid_constant__ CUtensorMap map_w) {
  using D = TileDims<KH, FRAC, TH>;
  constexpr int K = D::K;
  constexpr int FT_H = TH, FT_NT = D::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* bufA = reinterpret_cast<double*>(smem_raw);  // xs [XH][XW] -> bx [TR][BP] -> t2 [ZH][T2P]
  double* bufB = bufA + D::A_DOUBLES;                  // ws [WH][WW] -> tmp [TR][TP] -> z [ZH][ZP]
  int* s_pb = reinterpret_cast<int*>(bufB + D::B_DOUBLES);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + D::FIXED_BYTES - 16);
  TEntry* s_ents = reinterpret_cast<TEntry*>(smem_raw + D::FIXED_BYTES);
  double* xs = bufA;
  double* ws = bufB;
  double* tmp = bufB;
  double* bx = bufA;
  double* zs = bufB;
  double* t2 = bufA;

  const int tid = threadIdx.x;
  // blockIdx.y walks "units": (channel, tile row) pairs in memory order, from P.unit_begin
  const int unit = P.unit_begin + blockIdx.y;
  const int ch = unit / P.tile_rows;
  const int tx0 = blockIdx.x * FT_W, ty0 = (unit - ch * P.tile_rows) * FT_H;
  const size_t HW = (size_t)P.H * P.W;
  const int s = P.s, sh = P.sshift;

  // LR cells under the Z region of this tile, and whether all their samples are regular
  const int mr_lo = floordiv_scale(ty0 - KH, s, sh), mr_hi = floordiv_scale(ty0 + FT_H + KH - 1, s, sh);
  const int mc_lo = floordiv_scale(tx0 - KH, s, sh), mc_hi = floordiv_scale(tx0 + FT_W + KH - 1, s, sh);
  const bool interior = mr_lo + P.qoff_min_r >= P.lo_r && mr_hi + P.qoff_max_r < P.hi_r &&
                        mc_lo + P.qoff_min_c >= P.lo_c && mc_hi + P.qoff_max_c < P.hi_c &&
                        ty0 - KH >= 0 && tx0 - KH >= 0 && ty0 + FT_H + KH <= P.H && tx0 + FT_W + KH <= P.W;
  const double* __restrict__ ych = P.y + (size_t)(P.c0 + ch) * ((size_t)P.h * P.w);
  const bool fastpath = !FRAC && interior && P.fast != nullptr;

  // ---- 0. stage the x tile + halo and the IRLS weights (zero outside the image) ------------------
  if (P.use_tma) {
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, D::X_BYTES + (P.reg_fused ? D::W_BYTES : 0u));
      tma_load_3d(xs, &map_x, tx0 - D::HXC, ty0 - D::HX, ch, bar);
      if (P.reg_fused) tma_load_3d(ws, &map_w, tx0 - 2, ty0 - 1, ch, bar);
    }
  }
  // phase table and entry list -> smem while the bulk copies are in flight
  const int nph = s * s + 1;
  for (int i = tid; i < nph; i += FT_NT) s_pb[i] = P.phase_begin[i];
  {
    const int4* src = reinterpret_cast<const int4*>(P.entries);
    int4* dst = reinterpret_cast<int4*>(s_ents);
    for (int i = tid; i < P.num_entries * 2; i += FT_NT) dst[i] = src[i];
  }
  __syncthreads();  // tables visible; mbarrier initialised before anyone waits on it
  if (P.use_tma) {
    mbar_wait(bar, 0);
  } else {
    const double* __restrict__ xc = P.x + (size_t)ch * HW;
    for (int id = tid; id < D::XH * D::XW; id += FT_NT) {
      const int r = id / D::XW, c = id - r * D::XW;
      const int gr = ty0 - D::HX + r, gc = tx0 - D::HXC + c;
      double v = 0.0;
      if (gr >= 0 && gr < P.H && gc >= 0 && gc < P.W) v = xc[(size_t)gr * P.W + gc];
      xs[id] = v;
    }
    if (P.reg_fused) {
      const double* __restrict__ wc = P.wts + (size_t)ch * HW;
      for (int id = tid; id < D::WH * D::WW; id += FT_NT) {
        const int r = id / D::WW, c = id - r * D::WW;
        const int gr = ty0 - 1 + r, gc = tx0 - 2 + c;
        double v = 0.0;
        if (gr >= 0 && gr < P.H && gc >= 0 && gc < P.W) v = wc[(size_t)gr * P.W + gc];
        ws[id] = v;
      }
    }
    __syncthreads();
  }

  // ---- 1. 2-D TV, IRLS weighted (tv_regularizer.cpp:134-227, objective_irls_regularization_term
  //         .cpp:27-55): d/dx_p sum_j lambda w_j r_j^2 with r = |gx| + |gy|.  With
  //         t_q = 2 lambda w_q r_q, A_q = sgn(gx_q) t_q, B_q = sgn(gy_q) t_q:
  //             dp = A_left + B_above - (A_p + B_p)
  //         (weights are zero outside the image, which realises the col > 0 / row > 0 guards). ------
  constexpr int ESEG = FT_NT / FT_W;  // 4
  constexpr int EL = FT_H / ESEG;     // 8
  static_assert(ESEG * FT_W == FT_NT && EL * ESEG == FT_H, "tile/thread shape");
  const int ec = tid % FT_W, er0 = (tid / FT_W) * EL;
  const int gc = tx0 + ec;
  const bool tile_inside = tx0 + FT_W < P.W && ty0 + FT_H < P.H;  // strictly: right/bottom neighbours exist
  double tvg[EL];
  double cost_reg = 0.0;
#pragma unroll
  for (int l = 0; l < EL; ++l) tvg[l] = 0.0;
  if (P.reg_fused) {
    if (tile_inside && ty0 >= P.row0 && ty0 + FT_H <= P.row1)
      tile_tv<KH, FRAC, TH, EL, false>(P, xs, ws, ty0, gc, ec, er0, tvg, cost_reg);
    else
      tile_tv<KH, FRAC, TH, EL, true>(P, xs, ws, ty0, gc, ec, er0, tvg, cost_reg);
    __syncthreads();  // ws (bufB) is overwritten by the vertical pass
  }

  // ---- 2a. vertical PSF pass: tmp[r][c] = sum_i u[i] * xs[r+i][c] ------------------------------
  {
    constexpr int NSEG = (FT_NT / D::TW) > 0 ? (FT_NT / D::TW) : 1;
    constexpr int L = (D::TR + NSEG - 1) / NSEG;
    const double* __restrict__ xo = xs + D::XOR * D::XW + D::XOC;
    for (int id = tid; id < D::TW * NSEG; id += FT_NT) {
      const int c = id % D::TW, seg = id / D::TW;
      const int r0 = seg * L;
      const double* __restrict__ src = xo + r0 * D::XW + c;
      double* __restrict__ dst = tmp + r0 * D::TP + c;
      slide_correlate<K, L, SRB_SLIDE_B>(P.u, D::TR - r0,
                                [&](int i) { return src[i * D::XW]; },
                                [&](int l, double v) { dst[l * D::TP] = v; });
    }
  }
  __syncthreads();


  // ---- 2b. horizontal PSF pass: bx[r][c] = sum_j v[j] * tmp[r][c+j]   (bx overwrites xs) --------
  {
    constexpr int NSEG = (FT_NT / D::TR) > 0 ? (FT_NT / D::TR) : 1;
    constexpr int L = (D::BW + NSEG - 1) / NSEG;
    for (int id = tid; id < D::TR * NSEG; id += FT_NT) {
      const int r = id % D::TR, seg = id / D::TR;
      const int c0 = seg * L;
      const double* __restrict__ src = tmp + r * D::TP + c0;
      double* __restrict__ dst = bx + r * D::BP + c0;
      slide_correlate<K, L, SRB_SLIDE_B>(P.v, D::BW - c0, [&](int j) { return src[j]; },
                                [&](int l, double v) { dst[l] = v; });
    }
  }
  __syncthreads();

  // ---- 3. residuals of the regular LR samples landing in the tile (+ halo), per HR pixel --------
  //   pass A: the FT_H x FT_W pixels the tile owns; work item = (column, row residue mod s) = one
  //           sub-pixel phase, i.e. one entry list; walking down its rows walks down LR rows
  //   pass B: the halo ring of the Z region (not owned unless outside the image), pixel by pixel
  double cost_data = 0.0;
  if (fastpath) {
    // interior tile of a one-entry-per-phase model: table-driven work items (see TFast)
    const double* __restrict__ ytile = ych + ((long long)(ty0 >> sh) * P.w + (tx0 >> sh));
    constexpr int NITEM_A_PER_S = FT_W * (FT_H / 32);
    constexpr int NRING = 2 * KH * D::ZW + FT_H * 2 * KH;
    const int nitem_a = NITEM_A_PER_S * s;
    const int nj = 32 >> sh;  // rows of one pass-A item (s divides 32)
    const int bstep = s * D::BP, zstep = s * D::ZP;
    const size_t ystep = (size_t)P.w;
    const int4* __restrict__ tab = reinterpret_cast<const int4*>(P.fast);
    {
      for (int id = tid; id < nitem_a; id += FT_NT) {
        const int4 f = __ldg(tab + id);
        const double* __restrict__ yp = ytile + (((long long)f.y << 32) | (unsigned)f.x);
        const double* __restrict__ bp = bx + f.z;
        double* __restrict__ zp = zs + f.w;
        int j = 0;
        for (; j + 8 <= nj; j += 8) {
          double o[8], r[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] = __ldg(yp + t * ystep);
#pragma unroll
          for (int t = 0; t < 8; ++t) r[t] = bp[t * bstep] - o[t];
#pragma unroll
          for (int t = 0; t < 8; ++t) zp[t * zstep] = r[t];
#pragma unroll
          for (int t = 0; t < 8; ++t) cost_data = fma(r[t], r[t], cost_data);
          yp += 8 * ystep; bp += 8 * bstep; zp += 8 * zstep;
        }
        for (; j < nj; ++j) {
          const double r0 = bp[0] - __ldg(yp);
          zp[0] = r0;
          cost_data = fma(r0, r0, cost_data);
          yp += ystep; bp += bstep; zp += zstep;
        }
      }
    }
    // halo ring: two pixels per thread and iteration, both loads in flight before the first use
    for (int id = tid; id < NRING; id += 2 * FT_NT) {
      const int id2 = id + FT_NT;
      const int4 f0 = __ldg(tab + nitem_a + id);
      const int4 f1 = id2 < NRING ? __ldg(tab + nitem_a + id2) : f0;
      const double o0 = __ldg(ytile + (((long long)f0.y << 32) | (unsigned)f0.x));
      const double o1 = __ldg(ytile + (((long long)f1.y << 32) | (unsigned)f1.x));
      zs[f0.w] = bx[f0.z] - o0;
      if (id2 < NRING) zs[f1.w] = bx[f1.z] - o1;
    }
  } else {
    constexpr int RB = 32;  // rows per work item block
    for (int id = tid; id < FT_W * s * (FT_H / RB); id += FT_NT) {
      const int blk = id / (FT_W * s), id2 = id - blk * (FT_W * s);
      const int rho = blk * RB + id2 / FT_W, cm = id2 % FT_W;  // first tile row of the item, tile column
      const int c = KH + cm;
      const int pc = tx0 + cm, pr = ty0 + rho;
      const int mc = floordiv_scale(pc, s, sh);
      int mr = floordiv_scale(pr, s, sh);
      const int phase = (pr - mr * s) * s + (pc - mc * s);
      const int e0 = s_pb[phase], e1 = s_pb[phase + 1];
      const double* __restrict__ ycell = ych + ((long long)mr * P.w + mc);
      const int nj = floordiv_scale((blk + 1) * RB - rho + s - 1, s, sh);  // rows rho, rho + s, ... of the block
      if (!FRAC && interior && e1 - e0 == 1) {
        // the common case (one frame per sub-pixel phase): entry in registers, LR loads batched
        const int4 head = *reinterpret_cast<const int4*>(s_ents + e0);
        const double* __restrict__ yp = ycell + (((long long)head.y << 32) | (unsigned)head.x);
        const double* __restrict__ bp = bx + (KH + rho) * D::BP + c + head.z;
        double* __restrict__ zp = zs + (KH + rho) * D::ZP + c;
        const int bstep = s * D::BP, zstep = s * D::ZP;
        const size_t ystep = (size_t)P.w;
        int j = 0;
#ifndef SRB_EXP_BATCH4
        for (; j + 8 <= nj; j += 8) {
          double o[8], r[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] = __ldg(yp + t * ystep);
#pragma unroll
          for (int t = 0; t < 8; ++t) r[t] = bp[t * bstep] - o[t];
#pragma unroll
          for (int t = 0; t < 8; ++t) zp[t * zstep] = r[t];
#pragma unroll
          for (int t = 0; t < 8; ++t) cost_data = fma(r[t], r[t], cost_data);
          yp += 8 * ystep; bp += 8 * bstep; zp += 8 * zstep;
        }
#endif
        for (; j + 4 <= nj; j += 4) {
          const double o0 = __ldg(yp), o1 = __ldg(yp + ystep), o2 = __ldg(yp + 2 * ystep), o3 = __ldg(yp + 3 * ystep);
          const double r0 = bp[0] - o0, r1 = bp[bstep] - o1, r2 = bp[2 * bstep] - o2, r3 = bp[3 * bstep] - o3;
          zp[0] = r0; zp[zstep] = r1; zp[2 * zstep] = r2; zp[3 * zstep] = r3;
          cost_data = fma(r0, r0, cost_data);
          cost_data = fma(r1, r1, cost_data);
          cost_data = fma(r2, r2, cost_data);
          cost_data = fma(r3, r3, cost_data);
          yp += 4 * ystep; bp += 4 * bstep; zp += 4 * zstep;
        }
        for (; j < nj; ++j) {
          const double r0 = bp[0] - __ldg(yp);
          zp[0] = r0;
          cost_data = fma(r0, r0, cost_data);
          yp += ystep; bp += bstep; zp += zstep;
        }
      } else {
        for (int j = 0; j < nj; ++j) {
          const int r = KH + rho + j * s;
          double z;
          if (interior) z = pixel_residuals<KH, FRAC, TH, false>(P, bx, s_ents, e0, e1, ycell, mr, mc, r, c, true, cost_data);
          else z = pixel_residuals<KH, FRAC, TH, true>(P, bx, s_ents, e0, e1, ycell, mr, mc, r, c, true, cost_data);
          zs[r * D::ZP + c] = z;
          ++mr;
          ycell += P.w;
        }
      }
    }
    if (KH > 0) {
      constexpr int NROWRING = 2 * KH * D::ZW;           // top + bottom halo rows, full width
      constexpr int NRING = NROWRING + FT_H * 2 * KH;    // + left / right halo columns of the tile rows
      const bool first_col = tx0 == 0, last_col = tx0 + FT_W >= P.W;
      const bool first_row = ty0 == 0, last_row = ty0 + FT_H >= P.H;
      for (int id = tid; id < NRING; id += FT_NT) {
        int r, c;
        if (id < NROWRING) {
          const int rr = id / D::ZW;
          c = id - rr * D::ZW;
          r = rr < KH ? rr : rr + FT_H;
        } else {
          const int t = id - NROWRING;
          constexpr int HC = KH > 0 ? 2 * KH : 1;  // (the ring is empty when KH == 0)
          const int rm = t / HC, hc = t - rm * HC;
          r = KH + rm;
          c = hc < KH ? hc : hc + FT_W;
        }
        const int pc = tx0 - KH + c, pr = ty0 - KH + r;
        const int mc = floordiv_scale(pc, s, sh), mr = floordiv_scale(pr, s, sh);
        const int phase = (pr - mr * s) * s + (pc - mc * s);
        const int e0 = s_pb[phase], e1 = s_pb[phase + 1];
        const double* ycell = ych + ((long long)mr * P.w + mc);
        double z, dummy = 0.0;
        if (!FRAC && interior && e1 - e0 == 1) {
          const int4 head = *reinterpret_cast<const int4*>(s_ents + e0);
          z = bx[r * D::BP + c + head.z] - __ldg(ycell + (((long long)head.y << 32) | (unsigned)head.x));
        } else if (interior) {
          z = pixel_residuals<KH, FRAC, TH, false>(P, bx, s_ents, e0, e1, ycell, mr, mc, r, c, false, dummy);
        } else {
          // halo positions outside the image belong to the first / last tile row / column
          const bool in_r = (r >= KH && r < KH + FT_H) || (pr < 0 && first_row) || (pr >= P.H && last_row);
          const bool in_c = (c >= KH && c < KH + FT_W) || (pc < 0 && first_col) || (pc >= P.W && last_col);
          z = pixel_residuals<KH, FRAC, TH, true>(P, bx, s_ents, e0, e1, ycell, mr, mc, r, c, in_r && in_c, cost_data);
        }
        zs[r * D::ZP + c] = z;
      }
    }
  }
  __syncthreads();

  if (P.g != nullptr) {
    // ---- 4a. adjoint horizontal pass: t2[r][c] = sum_j u[j] * Z[r][c+j]  (t2 overwrites bx) -----
    {
      constexpr int NSEG = (FT_NT / D::ZH) > 0 ? (FT_NT / D::ZH) : 1;
      constexpr int L = (FT_W + NSEG - 1) / NSEG;
      for (int id = tid; id < D::ZH * NSEG; id += FT_NT) {
        const int r = id % D::ZH, seg = id / D::ZH;
        const int c0 = seg * L;
        const double* __restrict__ src = zs + r * D::ZP + c0;
        double* __restrict__ dst = t2 + r * D::T2P + c0;
        slide_correlate<K, L, SRB_SLIDE_B>(P.u, FT_W - c0, [&](int j) { return src[j]; },
                                  [&](int l, double v) { dst[l] = v; });
      }
    }
    __syncthreads();

    // ---- 4b. adjoint vertical pass + regularizer part + store --------------------------------
    {
      double win[K];
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = t2[(er0 + i) * D::T2P + ec];
      double* __restrict__ gp = P.g + (size_t)ch * HW + (size_t)(ty0 + er0) * P.W + gc;
      const size_t gstep = (size_t)P.W;
      const bool all_in = tx0 + FT_W <= P.W && ty0 + FT_H <= P.H;
      if (all_in) {
#pragma unroll
        for (int l = 0; l < EL; ++l) {
          win[K - 1] = t2[(er0 + l + K - 1) * D::T2P + ec];
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < K; ++i) acc = fma(P.v[i], win[i], acc);
#pragma unroll
          for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
          *gp = fma(P.two_s2, acc, tvg[l]);
          gp += gstep;
        }
      } else {
#pragma unroll
        for (int l = 0; l < EL; ++l) {
          const int r = er0 + l;
          win[K - 1] = t2[(r + K - 1) * D::T2P + ec];
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < K; ++i) acc = fma(P.v[i], win[i], acc);
#pragma unroll
          for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
          if (ty0 + r < P.H && gc < P.W) *gp = fma(P.two_s2, acc, tvg[l]);
          gp += gstep;
        }
      }
    }
  }

  // ---- cost partial sums (deterministic: fixed per-CTA slot, fixed-order final reduction) --------
  const size_t cta = (size_t)unit * gridDim.x + blockIdx.x;
  block_sum2<FT_NT>(cost_data, cost_reg);
  if (tid == 0) {
    P.part_data[cta] = P.s2 * cost_data;
    P.part_reg[cta] = cost_reg;
  }
}

// ---- border band ("special" samples) -----------------------------------------------------------
// The special LR samples form a frame-independent band: LR rows [0, lo_r) and [hi_r, h), and in
// the rows between, LR columns [0, lo_c) and [hi_c, w).  They are stored compactly:
//   index = rr * w + qc                      for the row bands (rr counts band rows top to bottom)
//         = n_rows_part + m * nc + cc        for the column bands (m = qr - lo_r, cc counts band columns)
struct BandGeom {
  int h, w;
  int lo_r, hi_r, lo_c, hi_c;
  __host__ __device__ int band_rows() const { return lo_r + (h - hi_r); }
  __host__ __device__ int band_cols() const { return lo_c + (w - hi_c); }
  __host__ __device__ long long rows_part() const { return (long long)band_rows() * w; }
  __host__ __device__ long long count() const {
    return rows_part() + (long long)(hi_r - lo_r) * band_cols();
  }
  // compact index of LR sample (qr, qc), or -1 when it is a regular sample
  __host__ __device__ long long index_of(int qr, int qc) const {
    if (qr < lo_r) return (long long)qr * w + qc;
    if (qr >= hi_r) return (long long)(lo_r + qr - hi_r) * w + qc;
    if (qc < lo_c) return rows_part() + (long long)(qr - lo_r) * band_cols() + qc;
    if (qc >= hi_c) return rows_part() + (long long)(qr - lo_r) * band_cols() + lo_c + (qc - hi_c);
    return -1;
  }
  __host__ __device__ void sample_of(long long i, int* qr, int* qc) const {
    if (i < rows_part()) {
      const int rr = (int)(i / w);
      *qc = (int)(i - (long long)rr * w);
      *qr = rr < lo_r ? rr : hi_r + (rr - lo_r);
    } else {
      const long long j = i - rows_part();
      const int nc = band_cols();
      const int m = (int)(j / nc), cc = (int)(j - (long long)m * nc);
      *qr = lo_r + m;
      *qc = cc < lo_c ? cc : hi_c + (cc - lo_c);
    }
  }
};

// Forward model + residual of the special samples in the reference's operation order
// (forward_pixel): pooled[(k*Ca + c) * count + i] = s^2-fold sum of r, cost partials s^2 r^2.
// grid: (ceil(count/256), N*Ca)
__global__ void __launch_bounds__(256)
k_band_forward(GenericParams P, BandGeom B, const double* __restrict__ x, const double* __restrict__ y,
               double* __restrict__ pooled, double* __restrict__ cost_partial) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long cnt = B.count();
  const int kc = blockIdx.y;
  const int k = kc / P.Ca, c = kc % P.Ca;
  double cost = 0.0;
  if (i < cnt) {
    int qr, qc;
    B.sample_of(i, &qr, &qc);
    const size_t HW = (size_t)P.H * P.W, hw = (size_t)P.h * P.w;
    const double pred = forward_pixel(P, x + (size_t)c * HW, k, qr, qc);
    const double obs = y[((size_t)k * P.Ct + P.c0 + c) * hw + (size_t)qr * P.w + qc];
    const double r = __dadd_rn(pred, -obs);
    double acc = 0.0;
    const int reps = P.s * P.s;
    for (int t = 0; t < reps; ++t) acc = __dadd_rn(acc, r);
    pooled[(size_t)kc * cnt + i] = acc;
    cost = (double)reps * (r * r);
  }
  const double bs = block_sum(cost);
  if (threadIdx.x == 0) cost_partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = bs;
}

// B^T D^T restricted to the special samples of one frame, at HR pixel (pr, pc) (cf.
// backproject_pixel).
__device__ __forceinline__ double band_backproject(const GenericParams& P, const BandGeom& B,
                                                   const double* __restrict__ pooled_kc, int pr, int pc) {
  if (pr < 0 || pr >= P.H || pc < 0 || pc >= P.W) return 0.0;
  const int s = P.s, K = P.K, hk = P.hk;
  int i0 = (hk - pr) % s;
  if (i0 < 0) i0 += s;
  int j0 = (hk - pc) % s;
  if (j0 < 0) j0 += s;
  double acc = 0.0;
  for (int i = i0; i < K; i += s) {
    const int zr = pr + i - hk;
    if (zr < 0 || zr >= P.H) continue;
    const int qr = zr / s;
    for (int j = j0; j < K; j += s) {
      const double kv = P.psf[j * K + i];  // transposed kernel
      const int zc = pc + j - hk;
      if (kv == 0.0 || zc < 0 || zc >= P.W) continue;
      const long long idx = B.index_of(qr, zc / s);
      if (idx < 0) continue;
      acc = __dadd_rn(acc, __dmul_rn(kv, pooled_kc[idx]));
    }
  }
  return acc;
}

// g[c][p] += 2 * sum_k warp_{-shift_k}( B^T D^T pooled_k )(p) over the HR pixels the special
// samples can reach: HR rows [0, R.lo_r) and [R.hi_r, H), and between them HR columns [0, R.lo_c)
// and [R.hi_c, W)  (R is the BandGeom of the HR-pixel band, B the one of the LR samples).
// grid: (ceil(R.count()/256), Ca)
__global__ void __launch_bounds__(256)
k_band_adjoint(GenericParams P, BandGeom B, BandGeom R, const double* __restrict__ pooled,
               double* __restrict__ g) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= R.count()) return;
  int pr, pc;
  R.sample_of(i, &pr, &pc);
  const int c = blockIdx.y;
  const long long cnt = B.count();
  double acc = 0.0;
  for (int k = 0; k < P.N; ++k) {
    const double* __restrict__ pk = pooled + ((size_t)k * P.Ca + c) * cnt;
    const int Y = P.rowY[(size_t)k * P.H + pr];
    const int X = 32 * pc + P.nX[k];
    const int sy = Y >> 5, fy = Y & 31, sx = X >> 5, fx = X & 31;
    double back;
    if ((fy | fx) == 0) {
      back = band_backproject(P, B, pk, sy, sx);
    } else if (sx >= P.W || sx + 1 < 0 || sy >= P.H || sy + 1 < 0) {
      back = 0.0;
    } else {
      const double wy1 = fy * (1.0 / 32.0), wy0 = (32 - fy) * (1.0 / 32.0);
      const double wx1 = fx * (1.0 / 32.0), wx0 = (32 - fx) * (1.0 / 32.0);
      back = __dmul_rn(band_backproject(P, B, pk, sy, sx), wy0 * wx0);
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, pk, sy, sx + 1), wy0 * wx1));
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, pk, sy + 1, sx), wy1 * wx0));
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, pk, sy + 1, sx + 1), wy1 * wx1));
    }
    acc = __dadd_rn(acc, __dmul_rn(2.0, back));
  }
  const size_t o = (size_t)c * P.H * P.W + (size_t)pr * P.W + pc;
  g[o] += acc;
}

// Multi-GPU reduce + all-gather of one rank's band (after every rank has scattered its partial rows
// into this rank's slots): out_r[band] = sum_s slots[s][band] in fixed slot order (deterministic),
// written to the gradient buffer of EVERY rank (peer stores over NVLink).
struct GatherParams {
  int world, rank;
  long long band_begin, band_len, band_cap;
  const double* slots;           // this rank's slot array [world][band_cap]; slot [rank] is unused:
  const double* own;             // ... this rank's own contribution is its local partial gradient band
  double* out[SRB_MAX_PEERS];    // gradient buffers of all ranks (peer mappings)
};
// Multi-GPU reduce + all-gather of this rank's band: out_r[band] = sum_s partial_s[band] in fixed
// rank order (deterministic), stored into the gradient buffer of EVERY rank (peer stores over
// NVLink).  Every block first waits (bounded spin on the local phase-0 flags) until all ranks'
// contributions have arrived; the last block to finish publishes the total cost locally and raises
// this rank's phase-1 flag on every rank after a system fence.
__global__ void __launch_bounds__(256)
k_sum_gather(GatherParams G, long long n, long long flag_base, unsigned long long epoch, unsigned int* done_counter,
             int* err) {
  if (threadIdx.x < G.world) {
    const volatile unsigned long long* f =
        reinterpret_cast<const volatile unsigned long long*>(G.out[G.rank] + flag_base) + threadIdx.x;
    unsigned long long spins = 0;
    while (*f < epoch)
      if (++spins > (1ull << 24)) {  // seconds: a rank is missing -- give up instead of hanging the GPU
        *err = 1;
        break;
      }
    __threadfence_system();
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((G.band_len | G.band_begin | G.band_cap) & 1) == 0) {
    // 16-byte accesses: 512 B per warp and store instruction on the link
    const long long len2 = G.band_len >> 1, cap2 = G.band_cap >> 1, first2 = G.band_begin >> 1;
    const double2* __restrict__ slots2 = reinterpret_cast<const double2*>(G.slots);
    const double2* __restrict__ own2 = reinterpret_cast<const double2*>(G.own);
    for (long long i = i0; i < len2; i += stride) {
      double2 acc = make_double2(0.0, 0.0);
      for (int s = 0; s < G.world; ++s) {
        const double2 v = s == G.rank ? own2[i] : slots2[(long long)s * cap2 + i];
        acc.x += v.x;
        acc.y += v.y;
      }
      for (int r = 0; r < G.world; ++r) reinterpret_cast<double2*>(G.out[r])[first2 + i] = acc;
    }
  } else {
    for (long long i = i0; i < G.band_len; i += stride) {
      double acc = 0.0;
      for (int s = 0; s < G.world; ++s) acc += s == G.rank ? G.own[i] : G.slots[(long long)s * G.band_cap + i];
      for (int r = 0; r < G.world; ++r) G.out[r][G.band_begin + i] = acc;
    }
  }
  // last block done: total cost + phase-1 flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done_counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int r = 0; r < G.world; ++r) acc += G.out[G.rank][n + 1 + r];
      G.out[G.rank][n] = acc;
      *done_counter = 0;
    }
    if (threadIdx.x < G.world) {
      __threadfence_system();
      volatile unsigned long long* f =
          reinterpret_cast<volatile unsigned long long*>(G.out[threadIdx.x] + flag_base) + G.world + G.rank;
      *f = epoch;
    }
  }
}

// End of a rank's scatter phase, one block: cost = fixed-order sum of the per-CTA partials (as
// k_finish_partials), posted into slot [rank] of every rank's cost array, then -- after a system
// fence -- the phase-0 flag of this rank is raised on every rank.  Runs after k_tile on the same
// stream, i.e. after all of this rank's gradient rows have been stored to their owners.
__global__ void __launch_bounds__(1024)
k_peer_finish_scatter(const double* __restrict__ pd, size_t nd, const double* __restrict__ pr, size_t nr,
                      double* __restrict__ cost, GatherParams G, long long cost_slot_base, long long flag_base,
                      unsigned long long epoch) {
  double a = 0.0, b = 0.0;
  for (size_t i = threadIdx.x; i < nd; i += blockDim.x) a += pd[i];
  for (size_t i = threadIdx.x; i < nr; i += blockDim.x) b += pr[i];
  a = block_sum(a);
  b = block_sum(b);
  __shared__ double total;
  if (threadIdx.x == 0) {
    cost[0] = a;
    cost[1] = b;
    cost[2] = total = a + b;
  }
  __syncthreads();
  if (threadIdx.x < G.world) {
    G.out[threadIdx.x][cost_slot_base + G.rank] = total;
    __threadfence_system();
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(G.out[threadIdx.x] + flag_base) + G.rank;
    *f = epoch;
  }
}

// Device-side barrier between the ranks, on flags that live behind every rank's gradient buffer:
// k_peer_signal (after this rank's kernels of the phase, same stream) publishes `epoch` into slot
// [phase][rank] of every rank; k_peer_wait spins until all ranks have published it.  The spin is
// bounded: on timeout it records the failure in *err and returns instead of hanging the GPU.
__global__ void k_peer_wait(const double* out_local, long long flag_base, int phase, int world,
                            unsigned long long epoch, int* err) {
  if (threadIdx.x < world) {
    const volatile unsigned long long* f =
        reinterpret_cast<const volatile unsigned long long*>(out_local + flag_base) + phase * world + threadIdx.x;
    unsigned long long spins = 0;
    while (*f < epoch) {
      if (++spins > (1ull << 24)) {  // seconds: a rank is missing
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
}

// cost[0] = sum(data partials), cost[1] = sum(reg partials), cost[2] = their sum (also written to
// *tail when given): fixed-order, deterministic.
__global__ void __launch_bounds__(1024)
k_finish_partials(const double* __restrict__ pd, size_t nd, const double* __restrict__ pr, size_t nr,
                  double* __restrict__ cost, double* __restrict__ tail) {
  double a = 0.0, b = 0.0;
  for (size_t i = threadIdx.x; i < nd; i += blockDim.x) a += pd[i];
  for (size_t i = threadIdx.x; i < nr; i += blockDim.x) b += pr[i];
  a = block_sum(a);
  b = block_sum(b);
  if (threadIdx.x == 0) {
    cost[0] = a;
    cost[1] = b;
    const double t = a + b;
    cost[2] = t;
    if (tail) *tail = t;
  }
}


// ================================================================================================
// host side
// ================================================================================================
struct TileState {
  bool supported = false;
  bool frac = false;
  int KH = 0;
  double u[9], v[9];
  TEntry* d_entries = nullptr;
  int* d_phase_begin = nullptr;
  int num_entries = 0;
  int qoff_min_r = 0, qoff_max_r = 0, qoff_min_c = 0, qoff_max_c = 0;
  BandGeom band{};       // special LR samples
  BandGeom reach{};      // HR pixels the special samples can reach
  bool has_band = false;
  double* d_pooled = nullptr;  // [N][Ct][band.count()]
  TFast* d_fast[2] = {nullptr, nullptr};  // tile height 32 / 64; NULL when the model does not qualify
  bool tma_ok = false;
  int tile_h = 32;         // SRB_TILE_H=32|64 overrides (tuning knob; 32 measured faster at cfg3)
  void* encode = nullptr;  // cuTensorMapEncodeTiled
  std::string why;         // why the tile kernel does not cover this model
};

inline TileState*& tile_state(srb_ctx* c) { return reinterpret_cast<TileState*&>(c->fused); }
inline const TileState* tile_state(const srb_ctx* c) { return reinterpret_cast<const TileState*>(c->fused); }
inline bool fused_supported(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  return st && st->supported;
}

inline void fused_teardown(srb_ctx* c) {
  TileState*& st = tile_state(c);
  if (!st) return;
  if (st->d_entries) cudaFree(st->d_entries);
  if (st->d_phase_begin) cudaFree(st->d_phase_begin);
  if (st->d_pooled) cudaFree(st->d_pooled);
  for (TFast* f : st->d_fast)
    if (f) cudaFree(f);
  delete st;
  st = nullptr;
}

// Rank-1 factorisation psf = u v^T (true for blur_module.cpp:20-22's outer-product Gaussian).
inline bool factor_separable(const std::vector<double>& psf, int K, double* u, double* v) {
  int bi = 0, bj = 0;
  double best = 0.0;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      if (std::fabs(psf[i * K + j]) > best) best = std::fabs(psf[i * K + j]), bi = i, bj = j;
  if (!(best > 0.0)) return false;
  const double pivot = psf[bi * K + bj];
  for (int i = 0; i < K; ++i) u[i] = psf[i * K + bj];
  for (int j = 0; j < K; ++j) v[j] = psf[bi * K + j] / pivot;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      if (std::fabs(psf[i * K + j] - u[i] * v[j]) > 8.0 * 2.220446049250313e-16 * best) return false;
  return true;
}

inline int pymod(int a, int b) {
  int m = a % b;
  return m < 0 ? m + b : m;
}
inline int pydiv(int a, int b) { return (a - pymod(a, b)) / b; }

// Is LR sample q (one dimension; HR size L, half PSF width hk, scale s) "special" for a frame whose
// quantised forward / transpose warps are n32 / t32 (1/32 px)?  Regular means: commuting the PSF
// with the shift changes neither the LR prediction nor the back-projected gradient, and every
// transpose tap lands inside the image or in the Z halo (PSF half width) of a border tile.
inline bool sample_is_special(int q, int L, int hk, int s, int n32, int t32) {
  const int n = n32 >> 5, fa = (n32 & 31) ? 1 : 0;
  const int nt = t32 >> 5, fb = (t32 & 31) ? 1 : 0;
  const int p0 = s * q;
  auto in = [L](int p) { return p >= 0 && p < L; };
  for (int b = 0; b <= fb; ++b)
    if (p0 - nt - b < -hk || p0 - nt - b >= L + hk) return true;  // beyond the Z halo of the border tiles
  for (int i = -hk; i <= hk; ++i) {
    if (in(p0 + i)) continue;
    for (int a = 0; a <= fa; ++a)
      if (in(p0 + i + n + a)) return true;   // forward: window tap clipped before, not after, the shift
    for (int b = 0; b <= fb; ++b)
      if (in(p0 + i - nt - b)) return true;  // transpose: G outside the image reaches a pixel inside
  }
  return false;
}

inline srb_status fused_setup(srb_ctx* c) {
  TileState* st = new TileState();
  tile_state(c) = st;
  const Geometry& G = c->g;
  if (G.K > 9) { st->why = "PSF larger than 9x9"; return SRB_OK; }
  if (G.s > FT_MAX_SCALE) { st->why = "downsampling scale larger than 8"; return SRB_OK; }
  if (!c->warps_uniform) { st->why = "a shift sits on a fixed-point rounding boundary"; return SRB_OK; }
  if (!factor_separable(c->psf_h, G.K, st->u, st->v)) { st->why = "PSF is not separable (rank 1)"; return SRB_OK; }
  st->KH = G.hk;
  st->frac = !c->warps_integer;
  const int s = G.s, hk = G.hk;
  const int FR = st->frac ? 1 : 0;

  // ---- band of special samples (frame independent: the union over frames) -----------------------
  int lo[2] = {0, 0}, hi[2] = {G.h, G.w};
  int max_shift = 0;
  for (int dim = 0; dim < 2; ++dim) {
    const int L = dim == 0 ? G.H : G.W, l = dim == 0 ? G.h : G.w;
    for (int k = 0; k < G.N; ++k) {
      const int n32 = dim == 0 ? c->warp_fwd[k].nY : c->warp_fwd[k].nX;
      const int t32 = dim == 0 ? c->warp_tr[k].nY : c->warp_tr[k].nX;
      max_shift = std::max(max_shift, std::max(std::abs(n32 >> 5), std::abs(t32 >> 5)) + 1);
      for (int q = 0; q < l; ++q) {
        if (!sample_is_special(q, L, hk, s, n32, t32)) continue;
        if (2 * q < l) lo[dim] = std::max(lo[dim], q + 1);
        else hi[dim] = std::min(hi[dim], q);
      }
    }
    if (lo[dim] >= hi[dim]) { st->why = "image too small for the shifts (every sample is a border sample)"; return SRB_OK; }
  }
  st->band = BandGeom{G.h, G.w, lo[0], hi[0], lo[1], hi[1]};
  st->has_band = st->band.count() > 0;
  {
    // HR pixels reachable from the band: PSF half width + the largest shift + 1 around its samples
    const int m = hk + max_shift + 1;
    BandGeom R{G.H, G.W, 0, G.H, 0, G.W};
    if (lo[0] > 0) R.lo_r = std::min(G.H, s * (lo[0] - 1) + m + 1);
    if (hi[0] < G.h) R.hi_r = std::max(0, s * hi[0] - m);
    if (lo[1] > 0) R.lo_c = std::min(G.W, s * (lo[1] - 1) + m + 1);
    if (hi[1] < G.w) R.hi_c = std::max(0, s * hi[1] - m);
    if (R.lo_r >= R.hi_r || R.lo_c >= R.hi_c) { st->why = "image too small for the shifts"; return SRB_OK; }
    st->reach = R;
  }

  // ---- phase lists of the regular samples -------------------------------------------------------
  const int HB = hk + FR;
  const int BP = (FT_W + 2 * HB) | 1;
  std::vector<std::vector<TEntry>> lists((size_t)s * s);
  const long long hw = (long long)G.h * G.w;
  bool first = true;
  for (int k = 0; k < G.N; ++k) {
    const int nY = c->warp_fwd[k].nY, nX = c->warp_fwd[k].nX;
    const int tY = c->warp_tr[k].nY, tX = c->warp_tr[k].nX;
    const int n_r = nY >> 5, n_c = nX >> 5, fy = nY & 31, fx = nX & 31;
    const int t_r = tY >> 5, t_c = tX >> 5, ty = tY & 31, tx = tX & 31;
    for (int a = 0; a <= (ty ? 1 : 0); ++a)
      for (int b = 0; b <= (tx ? 1 : 0); ++b)
        for (int pr = 0; pr < s; ++pr)
          for (int pc = 0; pc < s; ++pc) {
            if (pymod(pr + t_r + a, s) != 0 || pymod(pc + t_c + b, s) != 0) continue;
            const int qoff_r = pydiv(pr + t_r + a, s), qoff_c = pydiv(pc + t_c + b, s);
            const int dr = t_r + n_r + a, dc = t_c + n_c + b;
            // the Bx samples (and their bilinear partners) must stay inside the Bx halo
            if (dr < -FR || dr > 0 || dc < -FR || dc > 0) {
              st->why = "forward and transpose warps of a frame quantise too far apart";
              return SRB_OK;
            }
            if (std::abs(qoff_r) > 30000 || std::abs(qoff_c) > 30000) { st->why = "shift too large"; return SRB_OK; }
            TEntry e;
            e.yoff = (long long)k * G.Ct * hw + (long long)qoff_r * G.w + qoff_c;
            e.bxoff = (HB - hk + dr) * BP + (HB - hk + dc);
            e.qoff = (qoff_r & 0xffff) | (qoff_c << 16);
            const double wy = a ? ty / 32.0 : (32 - ty) / 32.0, wx = b ? tx / 32.0 : (32 - tx) / 32.0;
            e.wT = wy * wx;
            e.fy = (short)fy;
            e.fx = (short)fx;
            e.owner = (a == 0 && b == 0) ? 1 : 0;
            lists[(size_t)pr * s + pc].push_back(e);
            if (first) {
              st->qoff_min_r = st->qoff_max_r = qoff_r;
              st->qoff_min_c = st->qoff_max_c = qoff_c;
              first = false;
            }
            st->qoff_min_r = std::min(st->qoff_min_r, qoff_r); st->qoff_max_r = std::max(st->qoff_max_r, qoff_r);
            st->qoff_min_c = std::min(st->qoff_min_c, qoff_c); st->qoff_max_c = std::max(st->qoff_max_c, qoff_c);
          }
  }
  std::vector<TEntry> flat;
  std::vector<int> begin((size_t)s * s + 1, 0);
  for (size_t ph = 0; ph < lists.size(); ++ph) {
    begin[ph] = (int)flat.size();
    flat.insert(flat.end(), lists[ph].begin(), lists[ph].end());
  }
  begin[(size_t)s * s] = (int)flat.size();
  st->num_entries = (int)flat.size();
  if (st->num_entries > FT_MAX_ENTRIES) { st->why = "too many (frame, tap) entries for shared memory"; return SRB_OK; }
  if (cudaMalloc((void**)&st->d_entries, (flat.size() + 1) * sizeof(TEntry)) != cudaSuccess ||
      cudaMalloc((void**)&st->d_phase_begin, begin.size() * sizeof(int)) != cudaSuccess)
    return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (tile kernel tables)");
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_entries, flat.data(), flat.size() * sizeof(TEntry), cudaMemcpyHostToDevice));
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_phase_begin, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (st->has_band) {
    const size_t n = (size_t)G.N * G.Ct * (size_t)st->band.count();
    if (cudaMalloc((void**)&st->d_pooled, n * sizeof(double)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (border band residuals)");
  }

  // ---- table-driven residual pass (TFast): integer shifts, one entry per phase, s | tile size ----
  {
    bool one_each = !st->frac && (32 % s == 0);
    for (size_t ph = 0; one_each && ph < lists.size(); ++ph) one_each = lists[ph].size() == 1;
    for (int v = 0; one_each && v < 2; ++v) {
      const int TH = v == 0 ? 32 : 64;
      const int ZW = FT_W + 2 * hk, ZP = ZW | 1;
      std::vector<TFast> tab;
      auto item = [&](int r, int c) {  // Z-region pixel (r, c): tile-relative HR position (r-hk, c-hk)
        const int pr = r - hk, pc = c - hk;
        const int dmr = pydiv(pr, s), dmc = pydiv(pc, s);
        const TEntry& e = lists[(size_t)(pr - dmr * s) * s + (pc - dmc * s)][0];
        TFast f;
        f.yrel = e.yoff + (long long)dmr * G.w + dmc;
        f.bxo = r * BP + c + e.bxoff;
        f.zo = r * ZP + c;
        tab.push_back(f);
      };
      for (int blk = 0; blk < TH / 32; ++blk)          // pass A: id = blk*(FT_W*s) + rho*FT_W + cm
        for (int rho = 0; rho < s; ++rho)
          for (int cm = 0; cm < FT_W; ++cm) item(hk + blk * 32 + rho, hk + cm);
      for (int rr = 0; rr < 2 * hk; ++rr)             // ring: top + bottom halo rows, full width
        for (int c = 0; c < ZW; ++c) item(rr < hk ? rr : rr + TH, c);
      for (int rm = 0; rm < TH; ++rm)                 // ring: left / right halo columns
        for (int hc = 0; hc < 2 * hk; ++hc) item(hk + rm, hc < hk ? hc : hc + FT_W);
      if (cudaMalloc((void**)&st->d_fast[v], (tab.size() + 1) * sizeof(TFast)) != cudaSuccess)
        return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (tile kernel tables)");
      SRB_CUDA_CHECK(c, cudaMemcpy(st->d_fast[v], tab.data(), tab.size() * sizeof(TFast), cudaMemcpyHostToDevice));
    }
  }

  // ---- TMA: the tensor-map encoder comes from the driver through the runtime ---------------------
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess && fn != nullptr)
      st->encode = fn;
    else
      (void)cudaGetLastError();
    st->tma_ok = st->encode != nullptr && (G.W % 2 == 0);  // global strides must be multiples of 16 B
  }
  if (const char* e = getenv("SRB_TILE_H")) st->tile_h = atoi(e) == 64 ? 64 : 32;
  st->supported = true;
  return SRB_OK;
}

typedef CUresult (*srb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

// 3-D tensor map over [planes][H][W] doubles with a (box_w x box_h x 1) box, zero fill outside.
inline bool make_plane_map(const TileState* st, CUtensorMap* map, const double* base, int W, int H, int planes,
                           int box_w, int box_h) {
  if (((size_t)base & 15) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((srb_encode_tiled_fn)st->encode)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int KH, bool FRAC, int TH>
inline srb_status tile_launch(srb_ctx* c, TileParams& P, int unit_end) {
  using D = TileDims<KH, FRAC, TH>;
  const dim3 grid((P.W + FT_W - 1) / FT_W, unit_end - P.unit_begin, 1);
  const TileState* st = tile_state(c);
  CUtensorMap mx, mw;
  memset(&mx, 0, sizeof mx);
  memset(&mw, 0, sizeof mw);
  P.use_tma = 0;
  if (st->tma_ok) {
    bool ok = make_plane_map(st, &mx, P.x, P.W, P.H, P.Ca, D::XW, D::XH);
    if (ok && P.reg_fused) ok = make_plane_map(st, &mw, P.wts, P.W, P.H, P.Ca, D::WW, D::WH);
    P.use_tma = ok ? 1 : 0;
  }
  static const size_t smem_pad = getenv("SRB_SMEM_PAD") ? (size_t)atoi(getenv("SRB_SMEM_PAD")) : 0;  // occupancy experiments
  const size_t smem = D::smem_bytes(P.num_entries) + smem_pad;
  static size_t attr_set[64] = {};
  if (c->device >= 64 || attr_set[c->device] < smem) {
    SRB_CUDA_CHECK(c, cudaFuncSetAttribute(k_tile<KH, FRAC, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (c->device < 64) attr_set[c->device] = smem;
  }
  if (c->profiling) cudaEventRecord(c->ev[4], c->stream);
  k_tile<KH, FRAC, TH><<<grid, D::NT, smem, c->stream>>>(P, mx, mw);
  if (c->profiling) cudaEventRecord(c->ev[5], c->stream);
  return SRB_OK;
}

// Tile height the current model runs with, and the number of (channel, tile row) units.
inline int tile_height(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  return (c->g.H >= 256 && c->g.W >= 256) ? st->tile_h : 32;
}
inline int tile_rows_per_channel(const srb_ctx* c) {
  const int TH = tile_height(c);
  return (c->g.H + TH - 1) / TH;
}

struct TileLayout {  // cost-partial slots of one evaluation
  size_t nblocks, nband;
  dim3 bgrid;
};
inline TileLayout tile_layout(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  const Geometry& G = c->g;
  TileLayout L;
  L.nblocks = (size_t)((G.W + FT_W - 1) / FT_W) * tile_rows_per_channel(c) * c->Ca();
  const long long bcount = st->has_band ? st->band.count() : 0;
  L.bgrid = dim3((unsigned)((bcount + 255) / 256), (unsigned)(G.N * c->Ca()));
  L.nband = st->has_band ? (size_t)L.bgrid.x * L.bgrid.y : 0;
  return L;
}

// Data term (+ 2-D TV term when fused) of the (channel, tile row) units [unit_begin, unit_end) of
// the active channel range: writes their gradient rows and their cost partial sums.
inline srb_status fused_eval_units(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, int unit_begin,
                                   int unit_end, bool* reg_done) {
  const TileState* st = tile_state(c);
  const Geometry& G = c->g;
  const int Ca = c->Ca();
  TileParams P;
  P.H = G.H; P.W = G.W; P.h = G.h; P.w = G.w; P.s = G.s; P.Ct = G.Ct; P.c0 = c->c0; P.Ca = Ca;
  P.sshift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == G.s) P.sshift = b;
  P.x = d_x; P.y = c->d_y; P.g = d_g; P.wts = c->d_w;
  P.entries = st->d_entries; P.phase_begin = st->d_phase_begin; P.num_entries = st->num_entries;
  P.qoff_min_r = st->qoff_min_r; P.qoff_max_r = st->qoff_max_r;
  P.qoff_min_c = st->qoff_min_c; P.qoff_max_c = st->qoff_max_c;
  P.lo_r = st->band.lo_r; P.hi_r = st->band.hi_r; P.lo_c = st->band.lo_c; P.hi_c = st->band.hi_c;
  for (int i = 0; i < 9; ++i) P.u[i] = i < G.K ? st->u[i] : 0.0, P.v[i] = i < G.K ? st->v[i] : 0.0;
  P.s2 = (double)G.s * G.s;
  P.two_s2 = 2.0 * P.s2;
  P.two_lambda = 2.0 * c->lambda;
  P.reg_fused = (do_reg && c->reg_kind == SRB_REG_TV) ? 1 : 0;
  P.row0 = c->reg_row0; P.row1 = c->reg_row1;
  *reg_done = P.reg_fused != 0;
  const int TH = tile_height(c);
  P.tile_rows = tile_rows_per_channel(c);
  P.fast = st->d_fast[TH == 64 ? 1 : 0];
  P.unit_begin = unit_begin;
  const TileLayout L = tile_layout(c);
  const size_t need = 2 * L.nblocks + L.nband;
  if (need > c->partial_capacity) {
    if (c->d_partial) cudaFree(c->d_partial);
    c->d_partial = nullptr;
    c->partial_capacity = 0;
    if (cudaMalloc((void**)&c->d_partial, need * sizeof(double)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (cost partials)");
    c->partial_capacity = need;
  }
  // layout: [data partials of the tiles][data partials of the band][reg partials of the tiles]
  P.part_data = c->d_partial;
  P.part_reg = c->d_partial + L.nblocks + L.nband;
  srb_status rc = SRB_OK;
  const int key = (st->KH * 2 + (st->frac ? 1 : 0)) * 2 + (TH == 64 ? 1 : 0);
  switch (key) {
#define SRB_TILE_CASE(KH_, FR_)                                                                    \
    case ((KH_) * 2 + (FR_)) * 2: rc = tile_launch<KH_, (FR_) != 0, 32>(c, P, unit_end); break;  \
    case ((KH_) * 2 + (FR_)) * 2 + 1: rc = tile_launch<KH_, (FR_) != 0, 64>(c, P, unit_end); break;
    SRB_TILE_CASE(0, 0) SRB_TILE_CASE(0, 1) SRB_TILE_CASE(1, 0) SRB_TILE_CASE(1, 1) SRB_TILE_CASE(2, 0)
    SRB_TILE_CASE(2, 1) SRB_TILE_CASE(3, 0) SRB_TILE_CASE(3, 1) SRB_TILE_CASE(4, 0) SRB_TILE_CASE(4, 1)
#undef SRB_TILE_CASE
    default: return c->fail(SRB_ERR_STATE, "tile kernel: unsupported PSF size");
  }
  if (rc != SRB_OK) return rc;
  c->timing.kernel_launches += 1;
  return SRB_OK;
}

// After every unit has been evaluated: the border band (exact, reference order) and the cost.
// Leaves the data cost in d_cost[0], the fused regularization cost in d_cost[1] and their sum in
// d_cost[2] (and *tail).
inline srb_status fused_eval_finish(srb_ctx* c, const double* d_x, double* d_g, double* tail) {
  const TileState* st = tile_state(c);
  const Geometry& G = c->g;
  const int Ca = c->Ca();
  const TileLayout L = tile_layout(c);
  if (st->has_band) {
    GenericParams GP;
    GP.H = G.H; GP.W = G.W; GP.h = G.h; GP.w = G.w; GP.s = G.s; GP.K = G.K; GP.hk = G.hk;
    GP.N = G.N; GP.Ca = Ca; GP.Ct = G.Ct; GP.c0 = c->c0;
    GP.src_r = c->d_src_r; GP.src_c = c->d_src_c; GP.psf = c->d_psf;
    GP.rowY = c->d_rowY_fwd; GP.nX = c->d_nX_fwd;
    k_band_forward<<<L.bgrid, 256, 0, c->stream>>>(GP, st->band, d_x, c->d_y, st->d_pooled,
                                                   c->d_partial + L.nblocks);
    c->timing.kernel_launches += 1;
    if (d_g) {
      GP.rowY = c->d_rowY_tr; GP.nX = c->d_nX_tr;
      const dim3 rgrid((unsigned)((st->reach.count() + 255) / 256), (unsigned)Ca);
      k_band_adjoint<<<rgrid, 256, 0, c->stream>>>(GP, st->band, st->reach, st->d_pooled, d_g);
      c->timing.kernel_launches += 1;
    }
  }
  k_finish_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, L.nblocks + L.nband,
                                               c->d_partial + L.nblocks + L.nband, L.nblocks, c->d_cost, tail);
  c->timing.kernel_launches += 1;
  return SRB_OK;
}

inline srb_status fused_eval(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, double* tail,
                             bool* reg_done) {
  const int units = tile_rows_per_channel(c) * c->Ca();
  srb_status st = fused_eval_units(c, d_x, d_g, do_reg, 0, units, reg_done);
  if (st != SRB_OK) return st;
  return fused_eval_finish(c, d_x, d_g, tail);
}

}  // namespace srb
