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

// Frames with the same shift merged into groups (srb_kernels_tilez.cuh: n ||A x - mean||^2 + constant): the band
// kernels then walk the groups instead of the frames.  frame[g] = a frame of group g (its warp tables stand for
// the group), yband = the group means of the band samples [groups][Ct][count], n = frames per group.
// groups == 0: no merging, group g is frame g and the observations come from the full LR stack.
struct BandGroups {
  int groups, n;
  const int* frame;
  const double* yband;
};

// Row-band partition (every device evaluates the whole objective on its own gradient rows): the band kernels
// restricted to the rows of one device.  The device owns, of active channel c, HR rows [lo(c), hi(c)) with
//   lo(c) = c == ch_first ? row_first : 0,   hi(c) = c == ch_last ? row_last : H     (nothing outside [ch_first, ch_last]).
// The adjoint kernel adds into those rows only; the forward kernel evaluates the samples those rows can reach
// (s * qr within `pad` HR rows of them) and counts the cost of the samples whose own HR row s * qr they contain.
// accumulate = 1: cost slots are added to (several launches per evaluation; the caller zeroes them first).
struct BandRows {
  int ch_first, row_first, ch_last, row_last, pad, accumulate;
  __host__ __device__ bool rows_of(int c, int H, int* lo, int* hi) const {
    if (c < ch_first || c > ch_last) return false;
    *lo = c == ch_first ? row_first : 0;
    *hi = c == ch_last ? row_last : H;
    return *hi > *lo;
  }
};

// Forward model + residual of the special samples in the reference's operation order
// (forward_pixel): pooled[(k*Ca + c) * count + i] = s^2-fold sum of r, cost partials s^2 r^2.
// grid: (ceil(count/256), N*Ca)  (N = number of groups when frames are merged)
__global__ void __launch_bounds__(256)
k_band_forward(GenericParams P, BandGeom B, BandGroups M, BandRows RW, const double* __restrict__ x,
               const double* __restrict__ y, double* __restrict__ pooled, double* __restrict__ cost_partial) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long cnt = B.count();
  const int kc = blockIdx.y;
  const int kg = kc / P.Ca, c = kc % P.Ca;
  const int k = M.groups ? M.frame[kg] : kg;
  double cost = 0.0;
  int rlo = 0, rhi = P.H;
  const bool have_rows = RW.rows_of(c, P.H, &rlo, &rhi);
  int qr = 0, qc = 0;
  if (i < cnt) B.sample_of(i, &qr, &qc);
  const int hr = qr * P.s;  // the sample's own HR row
  if (i < cnt && have_rows && hr >= rlo - RW.pad && hr < rhi + RW.pad) {
    const size_t HW = (size_t)P.H * P.W, hw = (size_t)P.h * P.w;
    const double pred = forward_pixel(P, x + (size_t)c * HW, k, qr, qc);
    const double obs = M.groups ? M.yband[((size_t)kg * P.Ct + P.c0 + c) * cnt + i]
                                : y[((size_t)k * P.Ct + P.c0 + c) * hw + (size_t)qr * P.w + qc];
    const double r = __dadd_rn(pred, -obs);
    double acc = 0.0;
    const int reps = P.s * P.s;
    for (int t = 0; t < reps; ++t) acc = __dadd_rn(acc, r);
    const double nf = M.groups ? (double)M.n : 1.0;
    pooled[(size_t)kc * cnt + i] = nf * acc;
    if (hr >= rlo && hr < rhi) cost = nf * ((double)reps * (r * r));
  }
  const double bs = block_sum(cost);
  if (threadIdx.x == 0) {
    double* slot = cost_partial + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    *slot = RW.accumulate ? *slot + bs : bs;
  }
}

// B^T D^T restricted to the special samples of one frame, at HR pixel (pr, pc) (cf.
// backproject_pixel).
//   sshift = log2(s) when s is a power of two (shifts and masks instead of integer divisions), else -1
__device__ __forceinline__ double band_backproject(const GenericParams& P, const BandGeom& B, int sshift,
                                                   const double* __restrict__ pooled_kc, int pr, int pc) {
  if (pr < 0 || pr >= P.H || pc < 0 || pc >= P.W) return 0.0;
  const int s = P.s, K = P.K, hk = P.hk;
  int i0, j0;
  if (sshift >= 0) {
    i0 = (hk - pr) & (s - 1);
    j0 = (hk - pc) & (s - 1);
  } else {
    i0 = (hk - pr) % s;
    if (i0 < 0) i0 += s;
    j0 = (hk - pc) % s;
    if (j0 < 0) j0 += s;
  }
  double acc = 0.0;
  for (int i = i0; i < K; i += s) {
    const int zr = pr + i - hk;
    if (zr < 0 || zr >= P.H) continue;
    const int qr = sshift >= 0 ? zr >> sshift : zr / s;
    // rows of regular samples contribute through the column bands only
    const bool row_regular = qr >= B.lo_r && qr < B.hi_r;
    for (int j = j0; j < K; j += s) {
      const double kv = P.psf[j * K + i];  // transposed kernel
      const int zc = pc + j - hk;
      if (kv == 0.0 || zc < 0 || zc >= P.W) continue;
      const int qcol = sshift >= 0 ? zc >> sshift : zc / s;
      if (row_regular && qcol >= B.lo_c && qcol < B.hi_c) continue;
      const long long idx = B.index_of(qr, qcol);
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
k_band_adjoint(GenericParams P, BandGeom B, BandGeom R, BandGroups M, BandRows RW, int sshift,
               const double* __restrict__ pooled, double* __restrict__ g) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= R.count()) return;
  int pr, pc;
  R.sample_of(i, &pr, &pc);
  const int c = blockIdx.y;
  int rlo, rhi;
  if (!RW.rows_of(c, P.H, &rlo, &rhi) || pr < rlo || pr >= rhi) return;
  const long long cnt = B.count();
  double acc = 0.0;
  const int ng = M.groups ? M.groups : P.N;
  for (int kg = 0; kg < ng; ++kg) {
    const int k = M.groups ? M.frame[kg] : kg;
    const double* __restrict__ pk = pooled + ((size_t)kg * P.Ca + c) * cnt;
    const int Y = P.rowY[(size_t)k * P.H + pr];
    const int X = 32 * pc + P.nX[k];
    const int sy = Y >> 5, fy = Y & 31, sx = X >> 5, fx = X & 31;
    double back;
    if ((fy | fx) == 0) {
      back = band_backproject(P, B, sshift, pk, sy, sx);
    } else if (sx >= P.W || sx + 1 < 0 || sy >= P.H || sy + 1 < 0) {
      back = 0.0;
    } else {
      const double wy1 = fy * (1.0 / 32.0), wy0 = (32 - fy) * (1.0 / 32.0);
      const double wx1 = fx * (1.0 / 32.0), wx0 = (32 - fx) * (1.0 / 32.0);
      back = __dmul_rn(band_backproject(P, B, sshift, pk, sy, sx), wy0 * wx0);
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, sshift, pk, sy, sx + 1), wy0 * wx1));
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, sshift, pk, sy + 1, sx), wy1 * wx0));
      back = __dadd_rn(back, __dmul_rn(band_backproject(P, B, sshift, pk, sy + 1, sx + 1), wy1 * wx1));
    }
    acc = __dadd_rn(acc, __dmul_rn(2.0, back));
  }
  const size_t o = (size_t)c * P.H * P.W + (size_t)pr * P.W + pc;
  g[o] += acc;
}

// Group means of the band samples and the constant part of the merged data cost (once per srb_set_observations):
//   yband[(g * Ct + c) * count + i] = mean over the frames e of group g of y_e(sample i),
//   var_part[c * var_stride + var_offset + g * gridDim.x + blockIdx.x] = block sum of sum_e (y_e - mean)^2.
// The frames of group g are the entries of sub-pixel phase group_phase[g] (TEntry::yoff - its sample offset =
// the frame's base offset in the LR stack).   grid: (ceil(count / 256), groups * Ct)
struct TEntry;
template <class Entry>
__global__ void __launch_bounds__(256)
k_band_merge(BandGeom B, int Ct, int w, const Entry* __restrict__ entries, const int* __restrict__ phase_begin,
             const int* __restrict__ group_phase, const double* __restrict__ y, double* __restrict__ yband,
             double* __restrict__ var_part, size_t var_stride, size_t var_offset) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long cnt = B.count();
  const int g = blockIdx.y / Ct, c = blockIdx.y % Ct;
  const int ph = group_phase[g];
  const int e0 = phase_begin[ph], n = phase_begin[ph + 1] - e0;
  double var = 0.0;
  if (i < cnt) {
    int qr, qc;
    B.sample_of(i, &qr, &qc);
    const size_t hw = (size_t)B.h * B.w;
    const long long at = (long long)c * (long long)hw + (long long)qr * w + qc;
    double sum = 0.0;
    for (int e = 0; e < n; ++e) {
      const Entry en = entries[e0 + e];
      const long long base = en.yoff - ((long long)(short)(en.qoff & 0xffff) * w + (en.qoff >> 16));
      sum += y[base + at];
    }
    const double m = sum / (double)n;
    for (int e = 0; e < n; ++e) {
      const Entry en = entries[e0 + e];
      const long long base = en.yoff - ((long long)(short)(en.qoff & 0xffff) * w + (en.qoff >> 16));
      const double d = y[base + at] - m;
      var = fma(d, d, var);
    }
    yband[((size_t)g * Ct + c) * cnt + i] = m;
  }
  var = block_sum(var);
  if (threadIdx.x == 0) var_part[(size_t)c * var_stride + var_offset + (size_t)g * gridDim.x + blockIdx.x] = var;
}

// First stage of the cost reduction for large tile counts: block b sums a fixed contiguous chunk of the data
// partials and of the regularization partials into stage[b] / stage[gridDim.x + b] (fixed order: deterministic);
// k_finish_partials then closes over the 2 * gridDim.x stage values.
__global__ void __launch_bounds__(1024)
k_stage_partials(const double* __restrict__ pd, size_t nd, const double* __restrict__ pr, size_t nr,
                 double* __restrict__ stage) {
  const size_t cd = (nd + gridDim.x - 1) / gridDim.x, cr = (nr + gridDim.x - 1) / gridDim.x;
  const size_t d0 = (size_t)blockIdx.x * cd, d1 = d0 + cd < nd ? d0 + cd : nd;
  const size_t r0 = (size_t)blockIdx.x * cr, r1 = r0 + cr < nr ? r0 + cr : nr;
  double a = 0.0, b = 0.0;
  for (size_t i = d0 + threadIdx.x; i < d1; i += blockDim.x) a += pd[i];
  for (size_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) b += pr[i];
  a = block_sum(a);
  b = block_sum(b);
  if (threadIdx.x == 0) {
    stage[blockIdx.x] = a;
    stage[gridDim.x + blockIdx.x] = b;
  }
}

// k_finish_partials with the data partials in two ranges (a device's own tiles + the border band slots).
__global__ void __launch_bounds__(1024)
k_finish_partials3(const double* __restrict__ pd, size_t nd, const double* __restrict__ pb, size_t nb,
                   const double* __restrict__ pr, size_t nr, double* __restrict__ cost, double* __restrict__ tail) {
  double a = 0.0, b = 0.0;
  for (size_t i = threadIdx.x; i < nd; i += blockDim.x) a += pd[i];
  for (size_t i = threadIdx.x; i < nb; i += blockDim.x) a += pb[i];
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
