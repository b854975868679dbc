// srb_kernels_band.cuh -- the border band of "special" LR samples (exact, reference operation order)
// and the fixed-order final reduction of the cost partial sums.  See srb_kernels_tile.cuh for why a
// thin band of samples near the image border is evaluated apart from the fused tile kernel.
#pragma once
#include "srb_common.cuh"
#include "srb_kernels_generic.cuh"

namespace srb {

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

}  // namespace srb
