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
  // table-driven residual pass with 2 or 4 frames per phase (k_tile<..., FE > 1>); kept at the end so
  // that the one-frame-per-phase kernels see the parameter layout they were tuned with
  const long long* fast_y;  // observation offsets of entries 1 .. fast_E-1: [e-1][fast_items]
  int fast_E;          // (frame, tap) entries per sub-pixel phase in the table-driven pass (1, 2 or 4)
  int fast_items;      // items in the table (pass A + ring)
  // k_tile_z only (kept last for the same reason): the observations re-laid out on the HR grid
  const double* yz;    // [Ct][H][W]: yz(c, p) = the one regular LR sample that lands on HR pixel p
  // k_tile_zt with merged frames (several frames with the SAME shift on a sub-pixel phase, averaged at upload):
  // [Ct] per-channel constant s^2 * sum_p sum_e (y_e(p) - mean(p))^2 of the data cost; NULL when nothing is merged
  const double* yvar;
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
// mbar_wait that traps instead of spinning forever (a lost TMA must not hang the device).
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned phase) {
  unsigned done = 0;
  const unsigned a = smem_u32(bar);
  for (unsigned tries = 0; !done; ++tries) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(phase)
        : "memory");
    if (!done && tries > (1u << 24)) __trap();
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

// Orders earlier generic-proxy accesses of shared memory before later async-proxy (TMA) ones.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

// Table-driven residual pass for models with E > 1 frames per sub-pixel phase (integer shifts, so all
// entries of a phase sample the same Bx element): z = sum_e (Bx - y_e), cost += sum_e (Bx - y_e)^2.
// Entry 0 comes from the TFast item, entries 1..E-1 from P.fast_y.  Lives in its own kernel
// instantiations (k_tile<..., FE>): its register needs must not disturb the one-frame-per-phase path.
template <int KH, bool FRAC, int TH, int E>
__device__ __forceinline__ double fast_residuals_multi(const TileParams& P, const double* __restrict__ bx,
                                                       double* __restrict__ zs, const double* __restrict__ ytile,
                                                       int tid) {
  using D = TileDims<KH, FRAC, TH>;
  constexpr int NT = D::NT;
  constexpr int NRING = 2 * KH * D::ZW + TH * 2 * KH;
  constexpr int RB = (4 / E) > 0 ? (4 / E) : 1;  // rows per batch (4 loads in flight per thread)
  const int s = P.s, sh = P.sshift;
  const int nitem_a = FT_W * (TH / 32) * s;
  const int nj = 32 >> sh;
  const int bstep = s * D::BP, zstep = s * D::ZP;
  const size_t ystep = (size_t)P.w;
  const int4* __restrict__ tab = reinterpret_cast<const int4*>(P.fast);
  double cost = 0.0;
  for (int id = tid; id < nitem_a; id += NT) {
    const int4 f = __ldg(tab + id);
    const double* yp[E];
    yp[0] = ytile + (((long long)f.y << 32) | (unsigned)f.x);
#pragma unroll
    for (int e = 1; e < E; ++e) yp[e] = ytile + __ldg(P.fast_y + (size_t)(e - 1) * P.fast_items + id);
    const double* __restrict__ bp = bx + f.z;
    double* __restrict__ zp = zs + f.w;
    for (int j = 0; j < nj; j += RB) {
      double o[E][RB];
#pragma unroll
      for (int t = 0; t < RB; ++t)
#pragma unroll
        for (int e = 0; e < E; ++e) o[e][t] = (j + t < nj) ? __ldg(yp[e] + (size_t)(j + t) * ystep) : 0.0;
#pragma unroll
      for (int t = 0; t < RB; ++t) {
        if (j + t < nj) {
          const double b = bp[(j + t) * bstep];
          double z = 0.0;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const double r = b - o[e][t];
            z += r;
            cost = fma(r, r, cost);
          }
          zp[(j + t) * zstep] = z;
        }
      }
    }
  }
  for (int id = tid; id < NRING; id += NT) {
    const int4 f = __ldg(tab + nitem_a + id);
    double o[E];
    o[0] = __ldg(ytile + (((long long)f.y << 32) | (unsigned)f.x));
#pragma unroll
    for (int e = 1; e < E; ++e) o[e] = __ldg(ytile + __ldg(P.fast_y + (size_t)(e - 1) * P.fast_items + nitem_a + id));
    const double b = bx[f.z];
    double z = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) z += b - o[e];
    zs[f.w] = z;
  }
  return cost;
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

// The whole evaluation of one tile (the body of k_tile; k_tile_z runs it for its border tiles).
template <int KH, bool FRAC, int TH, int FE>
__device__ __forceinline__ void tile_body(const TileParams& P, const CUtensorMap& map_x, const CUtensorMap& map_w) {
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
  const bool fastpath = !FRAC && interior && P.fast != nullptr;  // the host passes the table of THIS FE only

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
  if (FE > 1 && fastpath) {
    const double* __restrict__ ytile = ych + ((long long)(ty0 >> sh) * P.w + (tx0 >> sh));
    cost_data += fast_residuals_multi<KH, FRAC, TH, (FE > 1 ? FE : 2)>(P, bx, zs, ytile, tid);
  } else if (fastpath) {
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

template <int KH, bool FRAC, int TH, int FE>
__global__ void __launch_bounds__(TH * (FT_W / 8), TH == 32 ? (KH <= 3 ? 4 : 3) : 2)
k_tile(const TileParams P, const __grid_constant__ CUtensorMap map_x,
       const __grid_constant__ CUtensorMap map_w) {
  tile_body<KH, FRAC, TH, FE>(P, map_x, map_w);
}

// ---- row-major "Z layout" variant (round 1; kept behind SRB_ZT=0 for A/B runs -- the default Z-layout kernel is
//      k_tile_zt, srb_kernels_tilez.cuh) ---------------------------------------
// For models with integer shifts and exactly one frame per sub-pixel phase, every HR pixel receives
// exactly one regular LR sample, and that sample reads the Bx element under it.  With the
// observations gathered ONCE (at upload, k_build_yz) onto the HR grid, the residual pass of an
// interior tile is Z = Bx - yz elementwise, and the yz tile arrives by TMA while the horizontal PSF
// pass runs instead of being fetched sample by sample afterwards:
//     x, w boxes -> A, B | TV | vertical pass A -> B | yz box -> A (async)  ||  horizontal pass IN
//     PLACE in B (the K-1 inputs a thread shares with its right neighbour are read before a
//     barrier) | Z = B - A in place in B, cost | adjoint horizontal B -> A | adjoint vertical -> g
// Shared memory and registers are those of k_tile (4 CTAs / SM); tiles that touch the border band run
// tile_body unchanged.
//   HOLES: sub-pixel phases without a frame (a frame shard of a multi-GPU run
//   holds N / G of the s^2 phases) are NaN in yz and contribute Z = 0.
template <int KH, bool HOLES>
__global__ void __launch_bounds__(256, KH <= 3 ? 4 : 3)
k_tile_z(const TileParams P, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
         const __grid_constant__ CUtensorMap map_y) {
  constexpr int TH = 32;
  using D = TileDims<KH, false, TH>;
  constexpr int K = D::K;
  constexpr int NT = D::NT;
  static_assert(KH >= 1 && NT == 256, "k_tile_z: PSF of at least 3x3, 256 threads");
  constexpr int HYC = (KH + 1) & ~1;       // even column halo of the yz box (FLOAT64 TMA alignment)
  constexpr int YW = FT_W + 2 * HYC;       // yz box, dense
  constexpr int YH = TH + 2 * KH;
  static_assert(YH * YW <= D::A_DOUBLES, "yz box fits buffer A");
  static_assert(D::TR * D::TP <= D::B_DOUBLES && D::ZH * D::T2P <= D::A_DOUBLES, "buffer sizes");
  static_assert(D::BW == D::ZW && D::TR == D::ZH && D::TR == YH, "integer shifts: Bx region == Z region");

  const int unit = P.unit_begin + blockIdx.y;
  const int ch = unit / P.tile_rows;
  const int tx0 = blockIdx.x * FT_W, ty0 = (unit - ch * P.tile_rows) * TH;
  const int s = P.s, sh = P.sshift;
  {
    const int mr_lo = floordiv_scale(ty0 - KH, s, sh), mr_hi = floordiv_scale(ty0 + TH + KH - 1, s, sh);
    const int mc_lo = floordiv_scale(tx0 - KH, s, sh), mc_hi = floordiv_scale(tx0 + FT_W + KH - 1, s, sh);
    const bool interior = mr_lo + P.qoff_min_r >= P.lo_r && mr_hi + P.qoff_max_r < P.hi_r &&
                          mc_lo + P.qoff_min_c >= P.lo_c && mc_hi + P.qoff_max_c < P.hi_c &&
                          ty0 - KH >= 0 && tx0 - KH >= 0 && ty0 + TH + KH <= P.H && tx0 + FT_W + KH <= P.W;
    if (!interior) {  // block-uniform
      tile_body<KH, false, TH, 1>(P, map_x, map_w);
      return;
    }
  }

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* bufA = reinterpret_cast<double*>(smem_raw);  // xs [XH][XW] -> yz [YH][YW] -> t2 [ZH][T2P]
  double* bufB = bufA + D::A_DOUBLES;                  // ws [WH][WW] -> tmp -> bx -> z, all [TR][TP]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + D::FIXED_BYTES - 16);
  const int tid = threadIdx.x;
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

  // ---- 1. 2-D TV (as in tile_body) ----------------------------------------------------------------
  constexpr int ESEG = NT / FT_W;  // 4
  constexpr int EL = TH / ESEG;    // 8
  const int ec = tid % FT_W, er0 = (tid / FT_W) * EL;
  const int gc = tx0 + ec;
  double tvg[EL];
  double cost_reg = 0.0;
#pragma unroll
  for (int l = 0; l < EL; ++l) tvg[l] = 0.0;
  if (P.reg_fused) {
    // an interior tile has its right / bottom neighbours (KH >= 1); only the row band can cut it
    if (ty0 >= P.row0 && ty0 + TH <= P.row1)
      tile_tv<KH, false, TH, EL, false>(P, bufA, bufB, ty0, gc, ec, er0, tvg, cost_reg);
    else
      tile_tv<KH, false, TH, EL, true>(P, bufA, bufB, ty0, gc, ec, er0, tvg, cost_reg);
    __syncthreads();  // ws (B) is overwritten by the vertical pass
  }

  // ---- 2a. vertical PSF pass: tmp[r][c] = sum_i u[i] * xs[r+i][c]   (A -> B) ------------------------
  {
    constexpr int NSEG = (NT / D::TW) > 0 ? (NT / D::TW) : 1;
    constexpr int L = (D::TR + NSEG - 1) / NSEG;
    const double* __restrict__ xo = bufA + D::XOR * D::XW + D::XOC;
    for (int id = tid; id < D::TW * NSEG; id += NT) {
      const int c = id % D::TW, seg = id / D::TW;
      const int r0 = seg * L;
      const double* __restrict__ src = xo + r0 * D::XW + c;
      double* __restrict__ dst = bufB + r0 * D::TP + c;
      slide_correlate<K, L, SRB_SLIDE_B>(P.u, D::TR - r0,
                                [&](int i) { return src[i * D::XW]; },
                                [&](int l, double v) { dst[l * D::TP] = v; });
    }
  }
  __syncthreads();

  // x is consumed: the observations of the Z region land in A while the horizontal pass runs
  if (tid == 0) {
    fence_proxy_async();
    mbar_expect_tx(bar, (unsigned)(YH * YW * sizeof(double)));
    tma_load_3d(bufA, &map_y, tx0 - HYC, ty0 - KH, P.c0 + ch, bar);
  }

  // ---- 2b. horizontal PSF pass, in place: bx[r][c] = sum_j v[j] * tmp[r][c+j]   (B -> B) ----------
  {
    constexpr int NSEG = (NT / D::TR) > 0 ? (NT / D::TR) : 1;
    constexpr int L = (D::BW + NSEG - 1) / NSEG;
    static_assert(D::TR * NSEG <= NT && L >= K - 1, "one row segment per thread");
    const bool active = tid < D::TR * NSEG;
    const int r = tid % D::TR, seg = tid / D::TR;
    const int c0 = seg * L;
    double* row = bufB + r * D::TP + c0;
    double tail[K - 1];  // inputs [c0+L, c0+L+K-1): the next segment overwrites them
#pragma unroll
    for (int i = 0; i < K - 1; ++i) tail[i] = (active && c0 + L + i < D::TW) ? row[L + i] : 0.0;
    __syncthreads();
    if (active)
      slide_correlate<K, L, SRB_SLIDE_B>(P.v, D::BW - c0,
                                [&](int j) { return j < L ? row[j < L ? j : 0] : tail[j < L ? 0 : j - L]; },
                                [&](int l, double v) { row[l] = v; });
  }
  __syncthreads();

  // ---- 3. residuals: Z = Bx - yz, elementwise and in place; cost over the pixels the tile owns ------
  mbar_wait_bounded(bar, 1);
  double cost_data = 0.0;
  {
    const int c = tid & (FT_W - 1), rq = tid / FT_W;  // rows rq, rq + 4, ...
    double* __restrict__ bp = bufB + rq * D::TP + KH + c;
    const double* __restrict__ yp = bufA + rq * YW + HYC + c;
#pragma unroll
    for (int it = 0; it < (YH + ESEG - 1) / ESEG; ++it) {
      const int r = rq + ESEG * it;
      if (ESEG * it + ESEG - 1 < YH || r < YH) {
        if (!HOLES) {
          const double res = bp[it * ESEG * D::TP] - yp[it * ESEG * YW];
          bp[it * ESEG * D::TP] = res;
          if (r >= KH && r < KH + TH) cost_data = fma(res, res, cost_data);
        } else {
          const double yv = yp[it * ESEG * YW];
          const double res = (yv == yv) ? bp[it * ESEG * D::TP] - yv : 0.0;
          bp[it * ESEG * D::TP] = res;
          if (r >= KH && r < KH + TH) cost_data = fma(res, res, cost_data);
        }
      }
    }
    // left / right halo columns of the Z region
    for (int id = tid; id < 2 * KH * YH; id += NT) {
      const int r = id / (2 * KH), hc = id - r * (2 * KH);
      const int cz = hc < KH ? hc : hc + FT_W;
      if (!HOLES) {
        bufB[r * D::TP + cz] -= bufA[r * YW + cz - KH + HYC];
      } else {
        const double yv = bufA[r * YW + cz - KH + HYC];
        bufB[r * D::TP + cz] = (yv == yv) ? bufB[r * D::TP + cz] - yv : 0.0;
      }
    }
  }
  __syncthreads();

  if (P.g != nullptr) {
    // ---- 4a. adjoint horizontal pass: t2[r][c] = sum_j u[j] * Z[r][c+j]   (B -> A) ----------------
    {
      constexpr int NSEG = (NT / D::ZH) > 0 ? (NT / D::ZH) : 1;
      constexpr int L = (FT_W + NSEG - 1) / NSEG;
      for (int id = tid; id < D::ZH * NSEG; id += NT) {
        const int r = id % D::ZH, seg = id / D::ZH;
        const int c0 = seg * L;
        const double* __restrict__ src = bufB + r * D::TP + c0;
        double* __restrict__ dst = bufA + r * D::T2P + c0;
        slide_correlate<K, L, SRB_SLIDE_B>(P.u, FT_W - c0, [&](int j) { return src[j]; },
                                  [&](int l, double v) { dst[l] = v; });
      }
    }
    __syncthreads();

    // ---- 4b. adjoint vertical pass + regularizer part + store (the tile is inside the image) ------
    {
      const double* __restrict__ t2 = bufA;
      double win[K];
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = t2[(er0 + i) * D::T2P + ec];
      double* __restrict__ gp = P.g + (size_t)ch * HW + (size_t)(ty0 + er0) * P.W + gc;
      const size_t gstep = (size_t)P.W;
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
    }
  }

  const size_t cta = (size_t)unit * gridDim.x + blockIdx.x;
  block_sum2<NT>(cost_data, cost_reg);
  if (tid == 0) {
    P.part_data[cta] = P.s2 * cost_data;
    P.part_reg[cta] = cost_reg;
  }
}

// yz(c, p) = the observation of the one (frame, LR pixel) sample landing on HR pixel p; 0 where that
// sample lies outside the LR image (such pixels belong to tiles that never take the Z path), NaN
// where the pixel's sub-pixel phase has no frame at all.
// grid: (ceil(W/256), H, Ct)
__global__ void __launch_bounds__(256)
k_build_yz(int H, int W, int h, int w, int s, const TEntry* __restrict__ entries,
           const int* __restrict__ phase_begin, const double* __restrict__ y, double* __restrict__ yz) {
  const int pc = blockIdx.x * 256 + threadIdx.x, pr = blockIdx.y, c = blockIdx.z;
  if (pc >= W) return;
  const int mr = pr / s, mc = pc / s;
  const int ph = (pr - mr * s) * s + (pc - mc * s);
  double v = 0.0;
  if (phase_begin[ph + 1] == phase_begin[ph]) {
    v = __longlong_as_double(0x7ff8000000000000LL);  // no frame on this phase (k_tile_z<.., HOLES>)
  } else {
    const TEntry e = entries[phase_begin[ph]];
    const int qr = mr + (int)(short)(e.qoff & 0xffff), qc = mc + (e.qoff >> 16);
    if (qr >= 0 && qr < h && qc >= 0 && qc < w)
      v = y[(size_t)c * ((size_t)h * w) + e.yoff + (long long)mr * w + mc];
  }
  yz[((size_t)c * H + pr) * W + pc] = v;
}

}  // namespace srb
