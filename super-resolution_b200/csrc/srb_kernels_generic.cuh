// srb_kernels_generic.cuh -- "reference-order" CUDA kernels of the data term.
//
// These kernels evaluate the image formation model frame by frame in exactly the operation order
// of the reference (objective_data_term.cpp:15-75 and the OpenCV calls behind it), with explicit
// round-to-nearest multiplies/adds (no FMA contraction), so the LR prediction and the data-term
// gradient are BIT-IDENTICAL to the CPU reference restatement.  They accept every model the
// reference accepts (any odd PSF size, non-separable / asymmetric PSFs, fractional shifts, no
// blur, no motion, forward-only images whose size is not divisible by the scale).  The fused tile
// kernel (srb_kernels_fused.cuh) is the fast path; it falls back to these only for models it
// does not cover -- never to the CPU.
#pragma once
#include "srb_common.cuh"

namespace srb {

__device__ __forceinline__ double at0(const double* __restrict__ img, int H, int W, int r, int c) {
  return (r >= 0 && r < H && c >= 0 && c < W) ? img[(size_t)r * W + c] : 0.0;
}

// One destination pixel of cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) for a pure translation
// (motion_module.cpp:18-24): Y / X are the fixed-point source coordinates (1/32 px).
__device__ __forceinline__ double warp_sample(const double* __restrict__ img, int H, int W, int Y,
                                              int X) {
  const int sy = Y >> 5, fy = Y & 31, sx = X >> 5, fx = X & 31;
  if ((fy | fx) == 0) return at0(img, H, W, sy, sx);  // weights (1,0,0,0)
  if (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0) return 0.0;
  // BilinearTab_f entries: products of multiples of 1/32, exact in float and in double.
  const double wy1 = fy * (1.0 / 32.0), wy0 = (32 - fy) * (1.0 / 32.0);
  const double wx1 = fx * (1.0 / 32.0), wx0 = (32 - fx) * (1.0 / 32.0);
  const double v0 = at0(img, H, W, sy, sx), v1 = at0(img, H, W, sy, sx + 1);
  const double v2 = at0(img, H, W, sy + 1, sx), v3 = at0(img, H, W, sy + 1, sx + 1);
  double v = __dmul_rn(v0, wy0 * wx0);
  v = __dadd_rn(v, __dmul_rn(v1, wy0 * wx1));
  v = __dadd_rn(v, __dmul_rn(v2, wy1 * wx0));
  v = __dadd_rn(v, __dmul_rn(v3, wy1 * wx1));
  return v;
}

// LR prediction of frame k, channel plane x_c, at LR pixel (qr, qc):
//   D B M_k x  =  sum_ij psf[i][j] * [p in image] * warp_k(x)(p),  p = src(q) + (i,j) - hk
// (image_model.cpp:86-91 -> warpAffine, filter2D with zero border, nearest decimation).
__device__ __forceinline__ double forward_pixel(const GenericParams& P, const double* __restrict__ x_c,
                                                int k, int qr, int qc) {
  const int sr = P.src_r[qr], sc = P.src_c[qc];
  const int nX = P.nX[k];
  const int* __restrict__ rowY = P.rowY + (size_t)k * P.H;
  double acc = 0.0;
  for (int i = 0; i < P.K; ++i) {
    const int pr = sr + i - P.hk;
    if (pr < 0 || pr >= P.H) continue;
    const int Y = rowY[pr];
    for (int j = 0; j < P.K; ++j) {
      const double kv = P.psf[i * P.K + j];
      const int pc = sc + j - P.hk;
      if (kv == 0.0 || pc < 0 || pc >= P.W) continue;
      const double v = warp_sample(x_c, P.H, P.W, Y, 32 * pc + nX);
      acc = __dadd_rn(acc, __dmul_rn(kv, v));
    }
  }
  return acc;
}

// Forward model + residual for every (frame, active channel, LR pixel).
//   mode 0: out = D B M_k x                          (ImageModel::ApplyToImage)
//   mode 1: out = additive-pooled residual, i.e. the s*s-fold sequential sum of
//           r = (D B M_k x) - y  (objective_data_term.cpp:29-59), and per-block partial sums of
//           s^2 * r^2 (the data cost) into cost_partial[blockIdx linear].
// grid: (ceil(w/32), ceil(h/8), N*Ca)
template <int kMode>
__global__ void __launch_bounds__(256)
k_forward_generic(GenericParams P, const double* __restrict__ x, const double* __restrict__ y,
                  double* __restrict__ out, double* __restrict__ cost_partial, double* __restrict__ resid = nullptr) {
  const int qc = blockIdx.x * 32 + threadIdx.x;
  const int qr = blockIdx.y * 8 + threadIdx.y;
  const int kc = blockIdx.z;
  const int k = kc / P.Ca, c = kc % P.Ca;
  double cost = 0.0;
  if (qc < P.w && qr < P.h) {
    const size_t HW = (size_t)P.H * P.W, hw = (size_t)P.h * P.w;
    const double pred = forward_pixel(P, x + (size_t)c * HW, k, qr, qc);
    const size_t o = ((size_t)k * P.Ca + c) * hw + (size_t)qr * P.w + qc;
    if (kMode == 0) {
      out[o] = pred;
    } else {
      const double obs = y[((size_t)k * P.Ct + P.c0 + c) * hw + (size_t)qr * P.w + qc];
      const double r = __dadd_rn(pred, -obs);
      // ResizeAdditiveInterpolation (image_data.cpp:116-133): the s*s replicated residuals are
      // added one by one into the LR pixel.
      double pooled = 0.0;
      const int reps = P.s * P.s;
      for (int t = 0; t < reps; ++t) pooled = __dadd_rn(pooled, r);
      out[o] = pooled;
      if (resid) resid[o] = r;   // strict-order cost (k_strict_data_cost)
      cost = (double)reps * (r * r);
    }
  }
  if (kMode == 1) {
    const double bs = block_sum(cost);
    if (threadIdx.x == 0 && threadIdx.y == 0)
      cost_partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = bs;
  }
}

// The data cost summed in the REFERENCE'S order (objective_data_term.cpp:36-50, 104-114): per frame one
// running sum over channels and HR pixels of residual * residual -- the residual image is the nearest-
// upsampled LR residual, so every LR residual enters s*s times, in raster order -- and the frame sums added
// in frame order.  One thread per frame; a parity device (srb_set_strict_cost), not a fast path.
__global__ void k_strict_data_cost(const double* __restrict__ resid, int N, int Ca, int H, int W, int h, int w, int s,
                                   double* __restrict__ frame_sums, double* __restrict__ cost_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) {
    double sum = 0.0;
    for (int c = 0; c < Ca; ++c) {
      const double* __restrict__ rc = resid + ((size_t)k * Ca + c) * ((size_t)h * w);
      for (int pr = 0; pr < H; ++pr) {
        const double* __restrict__ rr = rc + (size_t)(pr / s) * w;
        for (int pc = 0; pc < W; ++pc) {
          const double r = rr[pc / s];
          sum = __dadd_rn(sum, __dmul_rn(r, r));
        }
      }
    }
    frame_sums[k] = sum;
  }
}
__global__ void k_strict_sum_frames(const double* __restrict__ frame_sums, int N, double* __restrict__ cost_out) {
  double total = 0.0;
  for (int k = 0; k < N; ++k) total = __dadd_rn(total, frame_sums[k]);
  *cost_out = total;
}

// B^T D^T of one frame at HR pixel (pr, pc): filter2D with blur_kernel_.t() (blur_module.cpp:
// 30-36) over the zero-inserted LR image (image_data.cpp:99-115).  Only taps that land on an
// inserted sample (multiples of s) are visited; the skipped taps add +0.0 in the reference.
__device__ __forceinline__ double backproject_pixel(const GenericParams& P,
                                                    const double* __restrict__ lr, int pr, int pc) {
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
      acc = __dadd_rn(acc, __dmul_rn(kv, lr[(size_t)qr * P.w + zc / s]));
    }
  }
  return acc;
}

// Transpose model summed over frames at every HR pixel:
//   g[c][p] (+)= sum_k outer * warp_{-shift_k}( B^T D^T lr_k )(p)      (image_model.cpp:93-101,
//   objective_data_term.cpp:60-71 with outer = 2).
// P.rowY / P.nX hold the TRANSPOSE warp tables.  accumulate = false writes, true adds to g.
// grid: (ceil(W/32), ceil(H/8), Ca)
__global__ void __launch_bounds__(256)
k_adjoint_generic(GenericParams P, const double* __restrict__ lr, double* __restrict__ g,
                  double outer, int accumulate) {
  const int pc = blockIdx.x * 32 + threadIdx.x;
  const int pr = blockIdx.y * 8 + threadIdx.y;
  const int c = blockIdx.z;
  if (pc >= P.W || pr >= P.H) return;
  const size_t HW = (size_t)P.H * P.W, hw = (size_t)P.h * P.w;
  const size_t o = (size_t)c * HW + (size_t)pr * P.W + pc;
  double acc = accumulate ? g[o] : 0.0;
  for (int k = 0; k < P.N; ++k) {
    const double* __restrict__ lrk = lr + ((size_t)k * P.Ca + c) * hw;
    const int Y = P.rowY[(size_t)k * P.H + pr];
    const int X = 32 * pc + P.nX[k];
    const int sy = Y >> 5, fy = Y & 31, sx = X >> 5, fx = X & 31;
    double back;
    if ((fy | fx) == 0) {
      back = backproject_pixel(P, lrk, sy, sx);
    } else if (sx >= P.W || sx + 1 < 0 || sy >= P.H || sy + 1 < 0) {
      back = 0.0;
    } else {
      const double wy1 = fy * (1.0 / 32.0), wy0 = (32 - fy) * (1.0 / 32.0);
      const double wx1 = fx * (1.0 / 32.0), wx0 = (32 - fx) * (1.0 / 32.0);
      back = __dmul_rn(backproject_pixel(P, lrk, sy, sx), wy0 * wx0);
      back = __dadd_rn(back, __dmul_rn(backproject_pixel(P, lrk, sy, sx + 1), wy0 * wx1));
      back = __dadd_rn(back, __dmul_rn(backproject_pixel(P, lrk, sy + 1, sx), wy1 * wx0));
      back = __dadd_rn(back, __dmul_rn(backproject_pixel(P, lrk, sy + 1, sx + 1), wy1 * wx1));
    }
    acc = __dadd_rn(acc, __dmul_rn(outer, back));
  }
  g[o] = acc;
}

}  // namespace srb
