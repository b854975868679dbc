// srb_kernels_fused.cuh -- the fused tile kernel of the MAP objective (fast path).
//
// Algorithm.  The reference evaluates, frame by frame,
//     r_k = D B M_k x - y_k,      g += 2 s^2 M_k^T B^T D^T r_k          (objective_data_term.cpp:15-75)
// with B (PSF correlation) applied at full HR resolution twice per frame.  B and the translation
// M_k are both convolutions, so away from the image border they commute:
//     r_k = D M_k (B x) - y_k,    g = 2 s^2 B^T ( sum_k M_k^T D^T r_k ).
// One CTA owns one HR tile of one channel and does, entirely in shared memory:
//     1. stage the x tile (+halo, zero outside the image)            HBM -> smem, x read ONCE
//     2. Bx = separable PSF correlation of the tile                   (2 passes, sliding register windows)
//     3. Z(p) = sum over the LR samples (k,q) that land on HR pixel p of (Bx-sample - y_k(q));
//        every LR observation is read exactly once; cost += r^2     (warp-shuffle/block reduction)
//     4. g = 2 s^2 B^T Z                                              (2 passes)
//     5. + IRLS-weighted TV gradient from the same x tile, single store of g.
// Per-frame cost drops from 2*K^2/s^2 MACs per HR pixel to ~1 load, and the PSF work no longer
// scales with the number of frames.  Border exactness: see the "special samples" notes below.
#pragma once
#include <algorithm>
#include <cmath>

#include "srb_common.cuh"

namespace srb {

constexpr int FT_H = 32;    // tile rows
constexpr int FT_W = 64;    // tile columns
constexpr int FT_NT = 256;  // threads per CTA

// One way an LR sample can land on an HR pixel of a given sub-pixel phase.
struct FEntry {
  int k;             // frame
  int qoff_r, qoff_c;  // LR index = floor(p / s) + qoff
  short dr, dc;      // Bx sampling offset relative to p (integer part)
  short fy, fx;      // forward-warp bilinear fractions (1/32 px)
  int owner;         // 1 for the (0,0) transpose tap: counts the sample's cost
  double wT;         // transpose-warp bilinear weight of this tap
};

struct FusedParams {
  int H, W, h, w, s, N, Ca, Ct, c0;
  const double* x;
  const double* y;
  double* g;           // may be NULL (cost only)
  const double* wts;   // IRLS weights
  const FEntry* entries;
  const int* phase_begin;  // [s*s + 1]
  double u[9], v[9];   // psf[i][j] = u[i] * v[j]
  double two_s2;       // 2 * s^2
  double s2;           // s^2
  double lambda;
  int reg_fused;       // 1: 2-D TV term evaluated in the epilogue
  int row0, row1;      // HR row band of the regularizer term on this rank
  double* part_data;   // per-CTA partial sums of the data cost
  double* part_reg;    // per-CTA partial sums of the regularization cost
};

struct FusedState {
  bool supported = false;
  bool frac = false;
  int KH = 0;
  double u[9], v[9];
  FEntry* d_entries = nullptr;
  int* d_phase_begin = nullptr;
  int num_entries = 0;
  std::string why;  // why the fused kernel does not cover this model
};

inline FusedState*& fused_state(srb_ctx* c) {
  static_assert(sizeof(void*) == sizeof(FusedState*), "");
  return reinterpret_cast<FusedState*&>(c->fused);
}
inline const FusedState* fused_state(const srb_ctx* c) {
  return reinterpret_cast<const FusedState*>(c->fused);
}

template <int KH, bool FRAC>
struct FusedDims {
  static constexpr int K = 2 * KH + 1;
  static constexpr int HB = KH + (FRAC ? 1 : 0);  // halo of Bx around the tile
  static constexpr int HX = KH + HB;              // halo of x around the tile
  static constexpr int XH = FT_H + 2 * HX, XW = FT_W + 2 * HX, XP = XW;
  static constexpr int TR = FT_H + 2 * HB, TP = XW | 1;   // vertical-pass output
  static constexpr int BW = FT_W + 2 * HB, BP = BW | 1;   // Bx
  static constexpr int ZH = FT_H + 2 * KH, ZW = FT_W + 2 * KH, ZP = ZW | 1;  // Z (aliases tmp)
  static constexpr int T2P = FT_W | 1;                    // adjoint horizontal pass (aliases Bx)
  static constexpr int SMEM_DOUBLES = XH * XP + TR * TP + TR * BP;
  static constexpr size_t SMEM_BYTES = SMEM_DOUBLES * sizeof(double);
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

template <int KH, bool FRAC>
__global__ void __launch_bounds__(FT_NT)
k_fused(const FusedParams P) {
  using D = FusedDims<KH, FRAC>;
  constexpr int K = D::K;
  extern __shared__ __align__(16) double smem[];
  double* xs = smem;                       // [XH][XP]
  double* tmp = xs + D::XH * D::XP;        // [TR][TP]   later Z [ZH][ZP]
  double* bx = tmp + D::TR * D::TP;        // [TR][BP]   later t2 [ZH][T2P]
  double* zs = tmp;
  double* t2 = bx;

  const int tid = threadIdx.x;
  const int tx0 = blockIdx.x * FT_W, ty0 = blockIdx.y * FT_H;
  const int ch = blockIdx.z;
  const size_t HW = (size_t)P.H * P.W;
  const double* __restrict__ xc = P.x + (size_t)ch * HW;

  // ---- 1. stage x tile + halo (zero outside the image) ----------------------------------------
  for (int id = tid; id < D::XH * D::XW; id += FT_NT) {
    const int r = id / D::XW, c = id - r * D::XW;
    const int gr = ty0 - D::HX + r, gc = tx0 - D::HX + c;
    double v = 0.0;
    if (gr >= 0 && gr < P.H && gc >= 0 && gc < P.W) v = xc[(size_t)gr * P.W + gc];
    xs[r * D::XP + c] = v;
  }
  __syncthreads();

  // ---- 2a. vertical PSF pass: tmp[r][c] = sum_i u[i] * xs[r+i][c] ------------------------------
  {
    constexpr int NSEG = (FT_NT / D::XW) > 0 ? (FT_NT / D::XW) : 1;
    constexpr int L = (D::TR + NSEG - 1) / NSEG;
    for (int id = tid; id < D::XW * NSEG; id += FT_NT) {
      const int c = id % D::XW, seg = id / D::XW;
      const int r0 = seg * L;
      double win[K];
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = (r0 + i < D::XH) ? xs[(r0 + i) * D::XP + c] : 0.0;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int r = r0 + l;
        if (r < D::TR) {
          win[K - 1] = xs[(r + K - 1) * D::XP + c];
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < K; ++i) acc = fma(P.u[i], win[i], acc);
          tmp[r * D::TP + c] = acc;
#pragma unroll
          for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
        }
      }
    }
  }
  __syncthreads();

  // ---- 2b. horizontal PSF pass: bx[r][c] = sum_j v[j] * tmp[r][c+j] ----------------------------
  {
    constexpr int NSEG = (FT_NT / D::TR) > 0 ? (FT_NT / D::TR) : 1;
    constexpr int L = (D::BW + NSEG - 1) / NSEG;
    for (int id = tid; id < D::TR * NSEG; id += FT_NT) {
      const int r = id % D::TR, seg = id / D::TR;
      const int c0 = seg * L;
      const double* __restrict__ row = tmp + r * D::TP;
      double win[K];
#pragma unroll
      for (int j = 0; j < K - 1; ++j) win[j] = (c0 + j < D::XW) ? row[c0 + j] : 0.0;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int c = c0 + l;
        if (c < D::BW) {
          win[K - 1] = row[c + K - 1];
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j) acc = fma(P.v[j], win[j], acc);
          bx[r * D::BP + c] = acc;
#pragma unroll
          for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
        }
      }
    }
  }
  __syncthreads();

  // ---- 3. residuals of the LR samples landing in the tile, accumulated per HR pixel -------------
  double cost_data = 0.0;
  {
    const int s = P.s;
    const size_t hw = (size_t)P.h * P.w;
    const bool last_row_tile = ty0 + FT_H >= P.H, last_col_tile = tx0 + FT_W >= P.W;
    for (int id = tid; id < D::ZH * D::ZW; id += FT_NT) {
      const int r = id / D::ZW, c = id - r * D::ZW;
      const int pr = ty0 - KH + r, pc = tx0 - KH + c;  // global HR position (may be outside)
      const int mr = floordiv(pr, s), mc = floordiv(pc, s);
      const int phase = (pr - mr * s) * s + (pc - mc * s);
      const bool own = ((r >= KH && r < KH + FT_H) || (pr < 0 && ty0 == 0) || (pr >= P.H && last_row_tile)) &&
                       ((c >= KH && c < KH + FT_W) || (pc < 0 && tx0 == 0) || (pc >= P.W && last_col_tile));
      double z = 0.0;
      const int e1 = P.phase_begin[phase + 1];
      for (int e = P.phase_begin[phase]; e < e1; ++e) {
        const FEntry en = P.entries[e];
        const int qr = mr + en.qoff_r, qc = mc + en.qoff_c;
        if (qr < 0 || qr >= P.h || qc < 0 || qc >= P.w) continue;
        double pred;
        if (!FRAC) {
          pred = bx[(r + en.dr) * D::BP + (c + en.dc)];
        } else {
          const double* b = bx + (r + 1 + en.dr) * D::BP + (c + 1 + en.dc);
          const double wy1 = en.fy * (1.0 / 32.0), wy0 = 1.0 - wy1;
          const double wx1 = en.fx * (1.0 / 32.0), wx0 = 1.0 - wx1;
          pred = b[0] * (wy0 * wx0) + b[1] * (wy0 * wx1) + b[D::BP] * (wy1 * wx0) + b[D::BP + 1] * (wy1 * wx1);
        }
        const double obs = __ldg(P.y + ((size_t)en.k * P.Ct + P.c0 + ch) * hw + (size_t)qr * P.w + qc);
        const double res = pred - obs;
        z = fma(en.wT, res, z);
        if (own && en.owner) cost_data = fma(res, res, cost_data);
      }
      zs[r * D::ZP + c] = z;
    }
  }
  __syncthreads();

  double cost_reg = 0.0;
  if (P.g != nullptr) {
    // ---- 4a. adjoint horizontal pass: t2[r][c] = sum_j u[j] * Z[r][c+j] --------------------------
    {
      constexpr int NSEG = (FT_NT / D::ZH) > 0 ? (FT_NT / D::ZH) : 1;
      constexpr int L = (FT_W + NSEG - 1) / NSEG;
      for (int id = tid; id < D::ZH * NSEG; id += FT_NT) {
        const int r = id % D::ZH, seg = id / D::ZH;
        const int c0 = seg * L;
        const double* __restrict__ row = zs + r * D::ZP;
        double win[K];
#pragma unroll
        for (int j = 0; j < K - 1; ++j) win[j] = (c0 + j < D::ZW) ? row[c0 + j] : 0.0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const int c = c0 + l;
          if (c < FT_W) {
            win[K - 1] = row[c + K - 1];
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) acc = fma(P.u[j], win[j], acc);
            t2[r * D::T2P + c] = acc;
#pragma unroll
            for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- 4b. adjoint vertical pass + 5. regularizer epilogue + store ------------------------------
  {
    constexpr int NSEG = FT_NT / FT_W;          // 4
    constexpr int L = FT_H / NSEG;              // 8
    static_assert(NSEG * FT_W == FT_NT && L * NSEG == FT_H, "tile/thread shape");
    const int c = tid % FT_W, seg = tid / FT_W;
    const int r0 = seg * L;
    const int gc = tx0 + c;
    double win[K];
    if (P.g != nullptr) {
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = t2[(r0 + i) * D::T2P + c];
    }
    const double* __restrict__ wc = P.wts + (size_t)ch * HW;
    double* __restrict__ gcn = P.g ? P.g + (size_t)ch * HW : nullptr;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const int r = r0 + l;
      const int gr = ty0 + r;
      double out = 0.0;
      if (P.g != nullptr) {
        win[K - 1] = t2[(r + K - 1) * D::T2P + c];
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) acc = fma(P.v[i], win[i], acc);
#pragma unroll
        for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];
        out = P.two_s2 * acc;
      }
      const bool inside = gr < P.H && gc < P.W;
      if (P.reg_fused && inside && gr >= P.row0 && gr < P.row1) {
        // 2-D TV, IRLS weighted (tv_regularizer.cpp:134-227, objective_irls_regularization_term.cpp)
        const double* xp = xs + (r + D::HX) * D::XP + (c + D::HX);
        const double x0 = xp[0];
        const bool has_r = gc + 1 < P.W, has_b = gr + 1 < P.H;
        const double gx = has_r ? xp[1] - x0 : 0.0;
        const double gy = has_b ? xp[D::XP] - x0 : 0.0;
        const double v0 = fabs(gy) + fabs(gx);
        const size_t gi = (size_t)gr * P.W + gc;
        const double c0w = P.lambda * __ldg(wc + gi);
        double didi = 0.0;
        didi += (gx < 0.0) ? 1.0 : (gx > 0.0 ? -1.0 : 0.0);
        didi += (gy < 0.0) ? 1.0 : (gy > 0.0 ? -1.0 : 0.0);
        double part = 2.0 * c0w * v0 * didi;
        if (gc > 0) {
          const double xl = xp[-1];
          const double gxl = x0 - xl;                       // left pixel always has a right neighbour
          const double gyl = has_b ? xp[D::XP - 1] - xl : 0.0;
          const double vl = fabs(gyl) + fabs(gxl);
          const double sg = gxl > 0.0 ? 1.0 : (gxl < 0.0 ? -1.0 : 0.0);
          part += 2.0 * (P.lambda * __ldg(wc + gi - 1)) * vl * sg;
        }
        if (gr > 0) {
          const double xa = xp[-D::XP];
          const double gya = x0 - xa;
          const double gxa = has_r ? xp[-D::XP + 1] - xa : 0.0;
          const double va = fabs(gya) + fabs(gxa);
          const double sg = gya > 0.0 ? 1.0 : (gya < 0.0 ? -1.0 : 0.0);
          part += 2.0 * (P.lambda * __ldg(wc + gi - P.W)) * va * sg;
        }
        out += part;
        cost_reg = fma(c0w * v0, v0, cost_reg);
      }
      if (gcn && inside) gcn[(size_t)gr * P.W + gc] = out;
    }
  }

  // ---- cost partial sums (deterministic: fixed per-CTA slot, fixed-order final reduction) --------
  const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const double cd = block_sum(cost_data);
  const double cr = block_sum(cost_reg);
  if (tid == 0) {
    P.part_data[cta] = P.s2 * cd;
    P.part_reg[cta] = cr;
  }
}

// out[slot0] = sum(a), out[slot1] = sum(b): two fixed-order reductions in one launch.
__global__ void k_reduce_partials2(const double* __restrict__ a, const double* __restrict__ b, size_t n,
                                   double* __restrict__ out) {
  const double* src = blockIdx.x == 0 ? a : b;
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += src[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// ---- host side ---------------------------------------------------------------------------------
inline bool fused_supported(const srb_ctx* c) {
  const FusedState* st = fused_state(c);
  return st && st->supported;
}

inline void fused_teardown(srb_ctx* c) {
  FusedState*& st = fused_state(c);
  if (!st) return;
  if (st->d_entries) cudaFree(st->d_entries);
  if (st->d_phase_begin) cudaFree(st->d_phase_begin);
  delete st;
  st = nullptr;
}

inline void fused_reg_changed(srb_ctx*) {}

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

inline srb_status fused_setup(srb_ctx* c) {
  FusedState* st = new FusedState();
  fused_state(c) = st;
  const Geometry& G = c->g;
  if (G.K > 9) { st->why = "PSF larger than 9x9"; return SRB_OK; }
  if (!c->warps_uniform) { st->why = "a shift sits on a fixed-point rounding boundary"; return SRB_OK; }
  if (!factor_separable(c->psf_h, G.K, st->u, st->v)) { st->why = "PSF is not separable (rank 1)"; return SRB_OK; }
  st->KH = G.hk;
  st->frac = !c->warps_integer;
  const int s = G.s, h = G.hk;
  // Special samples (window crossing the image border with content shifted across it, or living
  // farther than the PSF half width outside the image) are not handled by this kernel yet.
  for (int k = 0; k < G.N; ++k) {
    for (int dim = 0; dim < 2; ++dim) {
      const int n32 = dim == 0 ? c->warp_fwd[k].nY : c->warp_fwd[k].nX;
      const int t32 = dim == 0 ? c->warp_tr[k].nY : c->warp_tr[k].nX;
      const int n = n32 >> 5, nt = t32 >> 5;
      const int amax = (n32 & 31) ? 1 : 0, tmax = (t32 & 31) ? 1 : 0;
      const int size = dim == 0 ? G.H : G.W, lsize = dim == 0 ? G.h : G.w;
      const bool far = nt > h || nt < -(s - 1 + h);
      const bool low_cross = (h > 0) && (n + amax >= 1 || nt <= -1);
      const bool high_cross = (s * (lsize - 1) + h > size - 1) && (n <= -1 || nt + tmax >= 1);
      if (far || low_cross || high_cross) {
        st->why = "shifted content crosses the image border inside a PSF window (special samples)";
        return SRB_OK;
      }
    }
  }
  // phase lists
  std::vector<std::vector<FEntry>> lists((size_t)s * s);
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
            FEntry e;
            e.k = k;
            e.qoff_r = (pr + t_r + a) / s;
            e.qoff_c = (pc + t_c + b) / s;
            e.dr = (short)(t_r + n_r + a);
            e.dc = (short)(t_c + n_c + b);
            e.fy = (short)fy;
            e.fx = (short)fx;
            e.owner = (a == 0 && b == 0) ? 1 : 0;
            const double wy = a ? ty / 32.0 : (32 - ty) / 32.0, wx = b ? tx / 32.0 : (32 - tx) / 32.0;
            e.wT = wy * wx;
            lists[(size_t)pr * s + pc].push_back(e);
          }
  }
  std::vector<FEntry> flat;
  std::vector<int> begin((size_t)s * s + 1, 0);
  for (size_t ph = 0; ph < lists.size(); ++ph) {
    begin[ph] = (int)flat.size();
    flat.insert(flat.end(), lists[ph].begin(), lists[ph].end());
  }
  begin[(size_t)s * s] = (int)flat.size();
  st->num_entries = (int)flat.size();
  if (cudaMalloc((void**)&st->d_entries, (flat.size() + 1) * sizeof(FEntry)) != cudaSuccess ||
      cudaMalloc((void**)&st->d_phase_begin, begin.size() * sizeof(int)) != cudaSuccess)
    return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (fused tables)");
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_entries, flat.data(), flat.size() * sizeof(FEntry), cudaMemcpyHostToDevice));
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_phase_begin, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice));
  st->supported = true;
  return SRB_OK;
}

template <int KH, bool FRAC>
inline srb_status fused_launch(srb_ctx* c, const FusedParams& P, dim3 grid) {
  using D = FusedDims<KH, FRAC>;
  SRB_CUDA_CHECK(c, cudaFuncSetAttribute(k_fused<KH, FRAC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)D::SMEM_BYTES));
  k_fused<KH, FRAC><<<grid, FT_NT, D::SMEM_BYTES, c->stream>>>(P);
  return SRB_OK;
}

// Data term (+ 2-D TV term when fused) for the active channel range.  Leaves the data cost in
// d_cost[0] and the fused regularization cost in d_cost[1]; returns whether the regularizer was
// handled here through *reg_done.
inline srb_status fused_eval(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, bool* reg_done) {
  const FusedState* st = fused_state(c);
  const Geometry& G = c->g;
  FusedParams P;
  P.H = G.H; P.W = G.W; P.h = G.h; P.w = G.w; P.s = G.s; P.N = G.N; P.Ca = c->Ca(); P.Ct = G.Ct; P.c0 = c->c0;
  P.x = d_x; P.y = c->d_y; P.g = d_g; P.wts = c->d_w;
  P.entries = st->d_entries; P.phase_begin = st->d_phase_begin;
  for (int i = 0; i < 9; ++i) P.u[i] = i < G.K ? st->u[i] : 0.0, P.v[i] = i < G.K ? st->v[i] : 0.0;
  P.s2 = (double)G.s * G.s;
  P.two_s2 = 2.0 * P.s2;
  P.lambda = c->lambda;
  P.reg_fused = (do_reg && c->reg_kind == SRB_REG_TV) ? 1 : 0;
  P.row0 = c->reg_row0; P.row1 = c->reg_row1;
  *reg_done = P.reg_fused != 0;
  const dim3 grid((G.W + FT_W - 1) / FT_W, (G.H + FT_H - 1) / FT_H, c->Ca());
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (2 * nblocks > c->partial_capacity) {
    if (c->d_partial) cudaFree(c->d_partial);
    c->d_partial = nullptr;
    c->partial_capacity = 0;
    if (cudaMalloc((void**)&c->d_partial, 2 * nblocks * sizeof(double)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (cost partials)");
    c->partial_capacity = 2 * nblocks;
  }
  P.part_data = c->d_partial;
  P.part_reg = c->d_partial + nblocks;
  srb_status rc = SRB_OK;
  const int key = st->KH * 2 + (st->frac ? 1 : 0);
  switch (key) {
    case 0: rc = fused_launch<0, false>(c, P, grid); break;
    case 1: rc = fused_launch<0, true>(c, P, grid); break;
    case 2: rc = fused_launch<1, false>(c, P, grid); break;
    case 3: rc = fused_launch<1, true>(c, P, grid); break;
    case 4: rc = fused_launch<2, false>(c, P, grid); break;
    case 5: rc = fused_launch<2, true>(c, P, grid); break;
    case 6: rc = fused_launch<3, false>(c, P, grid); break;
    case 7: rc = fused_launch<3, true>(c, P, grid); break;
    case 8: rc = fused_launch<4, false>(c, P, grid); break;
    case 9: rc = fused_launch<4, true>(c, P, grid); break;
    default: return c->fail(SRB_ERR_STATE, "fused kernel: unsupported PSF size");
  }
  if (rc != SRB_OK) return rc;
  k_reduce_partials2<<<2, 1024, 0, c->stream>>>(P.part_data, P.part_reg, nblocks, c->d_cost);
  c->timing.kernel_launches += 2;
  return SRB_OK;
}

}  // namespace srb
