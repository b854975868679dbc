/*
 * sr_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See sr_oracle.h.
 *
 * Each function restates one piece of the reference hot path and cites the reference
 * file:line it follows.  The arithmetic of the four OpenCV entry points the reference
 * delegates to (not vendored, not pinned: CMakeLists.txt:5) is restated from OpenCV's
 * published algorithms and pinned against cv2 4.13 fixtures (tests/golden/).
 *
 * Build: oracle/Makefile  ->  oracle/_build/libsr_oracle.so   (gcc -O2, strict IEEE, no FMA
 * contraction, so results do not depend on the host's vector ISA).
 */
#include "sr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double* sro_alloc(size_t n) {
  double* p = (double*)malloc((n ? n : 1) * sizeof(double));
  if (!p) abort();
  return p;
}

/* ------------------------------------------------------------------------------------------
 * cv::getGaussianKernel(n, sigma, CV_64F) for sigma > 0: t_i = exp(-0.5 (i-(n-1)/2)^2 / sigma^2),
 * normalised to sum 1.  Reference call site: src/image_model/blur_module.cpp:20-21.
 * ---------------------------------------------------------------------------------------- */
void sro_gaussian_kernel(int n, double sigma, double* out) {
  const double scale2x = -0.5 / (sigma * sigma);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double x = i - (n - 1) * 0.5;
    const double t = exp(scale2x * x * x);
    out[i] = t;
    sum += t;
  }
  sum = 1.0 / sum;
  for (int i = 0; i < n; ++i) out[i] *= sum;
}

/* blur_kernel_ = kernel_x * kernel_y.t()  (blur_module.cpp:22): an outer product. */
void sro_gaussian_psf(int n, double sigma, double* out) {
  double* g = sro_alloc((size_t)n);
  sro_gaussian_kernel(n, sigma, g);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[i * n + j] = g[i] * g[j];
  free(g);
}

/* ------------------------------------------------------------------------------------------
 * cv::warpAffine with M = [1 0 dx; 0 1 dy], INTER_LINEAR, BORDER_CONSTANT 0, no WARP_INVERSE_MAP
 * (motion_module.cpp:18-24,29-51).  OpenCV inverts M, walks destination pixels and derives the
 * source coordinate in 1/1024 fixed point (AB_BITS = 10), adds round_delta = 16 and drops to
 * 1/32 px (INTER_BITS = 5); the bilinear weights are therefore exact multiples of 1/32.
 * ---------------------------------------------------------------------------------------- */
int sro_warp_quantize(double d) {
  /* X0 = saturate_cast<int>(-d * 1024) + 16;  (X0 + 1024 p) >> 5  ==  32 p + (X0 >> 5). */
  const long x0 = lrint(-d * 1024.0) + 16;
  return (int)(x0 >> 5); /* arithmetic shift: floor */
}

static inline double sro_at0(const double* src, int H, int W, int r, int c) {
  return (r >= 0 && r < H && c >= 0 && c < W) ? src[(size_t)r * W + c] : 0.0;
}

void sro_warp_shift(const double* src, int H, int W, double dx, double dy, double* dst) {
  const double m2 = -dx, m5 = -dy; /* translation of the inverted matrix */
  const long X0 = lrint((-0.0 * 0 + m2) * 1024.0) + 16;
  for (int y = 0; y < H; ++y) {
    const long Y0 = lrint((1.0 * y + m5) * 1024.0) + 16;
    const long Y = Y0 >> 5;
    const int sy = (int)(Y >> 5);
    const int fy = (int)(Y & 31);
    for (int x = 0; x < W; ++x) {
      const long X = (X0 + 1024L * x) >> 5;
      const int sx = (int)(X >> 5);
      const int fx = (int)(X & 31);
      /* BilinearTab_f entries (float, exact for 1/32 steps). */
      const double w0 = (double)((float)(32 - fy) / 32.0f * ((float)(32 - fx) / 32.0f));
      const double w1 = (double)((float)(32 - fy) / 32.0f * ((float)fx / 32.0f));
      const double w2 = (double)((float)fy / 32.0f * ((float)(32 - fx) / 32.0f));
      const double w3 = (double)((float)fy / 32.0f * ((float)fx / 32.0f));
      double v;
      if (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0) {
        v = 0.0;
      } else {
        const double v0 = sro_at0(src, H, W, sy, sx);
        const double v1 = sro_at0(src, H, W, sy, sx + 1);
        const double v2 = sro_at0(src, H, W, sy + 1, sx);
        const double v3 = sro_at0(src, H, W, sy + 1, sx + 1);
        v = v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
      }
      dst[(size_t)y * W + x] = v;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * cv::filter2D: correlation, anchor at the kernel centre (ksize/2), BORDER_CONSTANT 0
 * (matrix_util.cpp:12-29).  OpenCV's direct engine walks the NON-ZERO coefficients in row-major
 * order starting from delta = 0; same order here.  (For K*K >= 50 taps OpenCV switches to a
 * DFT-based path whose result differs from this sum by ~1e-16 relative.)
 * ---------------------------------------------------------------------------------------- */
void sro_filter2d(const double* src, int H, int W, const double* kernel, int kh, int kw,
                  double* dst) {
  const int ay = kh / 2, ax = kw / 2;
  for (int r = 0; r < H; ++r) {
    for (int c = 0; c < W; ++c) {
      double s = 0.0;
      for (int i = 0; i < kh; ++i) {
        const int rr = r + i - ay;
        if (rr < 0 || rr >= H) continue;
        for (int j = 0; j < kw; ++j) {
          const double kv = kernel[i * kw + j];
          const int cc = c + j - ax;
          if (kv == 0.0 || cc < 0 || cc >= W) continue;
          s += kv * src[(size_t)rr * W + cc];
        }
      }
      dst[(size_t)r * W + c] = s;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * cv::resize INTER_NEAREST: src index = min(floor(q * (1 / (n_dst / n_src))), n_src - 1), all in
 * fp64 (image_data.cpp:341-347).  This integer map must be reproduced bit-exactly.
 * ---------------------------------------------------------------------------------------- */
int sro_nearest_index(int q, int n_src, int n_dst) {
  const double inv_scale = (double)n_dst / (double)n_src;
  const double ifx = 1.0 / inv_scale;
  int s = (int)floor(q * ifx);
  if (s > n_src - 1) s = n_src - 1;
  return s;
}

void sro_resize_nearest(const double* src, int H, int W, double* dst, int H2, int W2) {
  int* xo = (int*)malloc(sizeof(int) * (size_t)(W2 ? W2 : 1));
  for (int x = 0; x < W2; ++x) xo[x] = sro_nearest_index(x, W, W2);
  for (int y = 0; y < H2; ++y) {
    const int sy = sro_nearest_index(y, H, H2);
    for (int x = 0; x < W2; ++x) dst[(size_t)y * W2 + x] = src[(size_t)sy * W + xo[x]];
  }
  free(xo);
}

/* ResizeAdditiveInterpolation (image_data.cpp:80-134). */
void sro_resize_additive(const double* src, int H, int W, double* dst, int H2, int W2) {
  memset(dst, 0, sizeof(double) * (size_t)H2 * W2);
  if (W <= W2 && H <= H2) { /* upsample: zero insertion at (row*ys, col*xs) */
    const int ys = H2 / H, xs = W2 / W;
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) dst[(size_t)(r * ys) * W2 + c * xs] = src[(size_t)r * W + c];
  } else { /* downsample: every HR pixel added into (row/ys, col/xs), row-major order */
    const int ys = H / H2, xs = W / W2;
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) dst[(size_t)(r / ys) * W2 + c / xs] += src[(size_t)r * W + c];
  }
}

/* ImageModel::ApplyToImage(ImageData*, k): operators in insertion order M, B, D
 * (image_model.cpp:17-61,86-91).  D: ResizeImage(1/s, NEAREST) with new size int(W * (1/s))
 * (downsampling_module.cpp:19-27, image_data.cpp:353-364). */
static void sro_lr_size(int s, int H, int W, int* h, int* w) {
  const double f = 1.0 / (double)s;
  *w = (int)(W * f);
  *h = (int)(H * f);
}

void sro_forward(const sro_model* m, int k, const double* hr, int H, int W, double* lr_out) {
  const size_t P = (size_t)H * W;
  double* a = sro_alloc(P);
  double* b = sro_alloc(P);
  memcpy(a, hr, P * sizeof(double));
  if (m->shifts) {
    sro_warp_shift(a, H, W, m->shifts[2 * k], m->shifts[2 * k + 1], b);
    double* t = a; a = b; b = t;
  }
  if (m->psf_size > 0) {
    sro_filter2d(a, H, W, m->psf, m->psf_size, m->psf_size, b);
    double* t = a; a = b; b = t;
  }
  int h, w;
  sro_lr_size(m->scale, H, W, &h, &w);
  sro_resize_nearest(a, H, W, lr_out, h, w);
  free(a);
  free(b);
}

/* ImageModel::ApplyTransposeToImage: reverse order D^T, B^T, M^T (image_model.cpp:93-101).
 * D^T = ResizeImage(s, ADDITIVE) zero insertion (downsampling_module.cpp:29-39);
 * B^T = filter2D with blur_kernel_.t() (blur_module.cpp:30-36) -- a transposed, NOT flipped, kernel;
 * M^T = warpAffine with (-dx, -dy) (motion_module.cpp:40-51). */
void sro_transpose(const sro_model* m, int k, const double* lr, int h, int w, double* hr_out) {
  const int s = m->scale;
  const int H = h * s, W = w * s;
  const size_t P = (size_t)H * W;
  double* a = sro_alloc(P);
  double* b = sro_alloc(P);
  sro_resize_additive(lr, h, w, a, H, W);
  if (m->psf_size > 0) {
    const int K = m->psf_size;
    double* kt = sro_alloc((size_t)K * K);
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) kt[i * K + j] = m->psf[j * K + i];
    sro_filter2d(a, H, W, kt, K, K, b);
    free(kt);
    double* t = a; a = b; b = t;
  }
  if (m->shifts) {
    sro_warp_shift(a, H, W, -m->shifts[2 * k], -m->shifts[2 * k + 1], b);
    double* t = a; a = b; b = t;
  }
  memcpy(hr_out, a, P * sizeof(double));
  free(a);
  free(b);
}

/* ------------------------------------------------------------------------------------------
 * ComputeTermForObservation (objective_data_term.cpp:15-75) for one frame and one channel:
 * continues the frame's running residual sum `sum` through this channel's pixels (:36-50 keeps ONE
 * accumulator over all channels of a frame) and returns it; adds 2 * A^T(...) into grad_c (may be NULL).
 * ---------------------------------------------------------------------------------------- */
static double sro_data_term_frame_channel(const sro_model* m, int k, const double* x_c, int H,
                                          int W, const double* obs_c, double* grad_c, double sum) {
  const size_t P = (size_t)H * W;
  int h, w;
  sro_lr_size(m->scale, H, W, &h, &w);
  double* lr = sro_alloc((size_t)h * w);
  double* up = sro_alloc(P);
  /* :27-29  degrade, then re-upsample with nearest interpolation */
  sro_forward(m, k, x_c, H, W, lr);
  sro_resize_nearest(lr, h, w, up, H, W);
  /* :36-50  residuals and their squared sum, pixel order */
  for (size_t p = 0; p < P; ++p) {
    const double r = up[p] - obs_c[p];
    up[p] = r;
    sum += r * r;
  }
  /* :55-71  additive downsample, transpose model, g += 2 * result */
  if (grad_c) {
    const int s = m->scale;
    const int h2 = H / s, w2 = W / s;
    double* rl = sro_alloc((size_t)h2 * w2);
    double* back = sro_alloc((size_t)h2 * s * w2 * s);
    sro_resize_additive(up, H, W, rl, h2, w2);
    sro_transpose(m, k, rl, h2, w2, back);
    for (size_t p = 0; p < P; ++p) grad_c[p] += 2 * back[p];
    free(rl);
    free(back);
  }
  free(lr);
  free(up);
  return sum;
}

/* ObjectiveDataTerm::Compute (objective_data_term.cpp:98-116): frames in order, channels in
 * order; serially the cost is summed exactly as the reference does (one running sum per frame over
 * its channels and pixels, frame sums added in frame order).  With num_threads > 1 frames x
 * channels run in parallel into private gradient buffers that are then summed in frame order, and
 * the cost is the sum of per-(frame, channel) sums: equal to the serial one up to re-association. */
double sro_data_term(const sro_model* m, const double* x, int H, int W, int C,
                     const double* obs_hr, int C_total, int channel_start, double* gradient,
                     int num_threads) {
  const size_t P = (size_t)H * W;
  const int N = m->num_frames; /* observations; shifts == NULL => model has no motion operator */
  const int jobs = N * C;
  double* costs = sro_alloc((size_t)jobs);
  if (num_threads <= 1) {
    /* the reference's order: frame-major, channels inside, gradient accumulated in place */
    double total = 0.0;
    for (int k = 0; k < N; ++k) {
      double frame = 0.0;
      for (int c = 0; c < C; ++c)
        frame = sro_data_term_frame_channel(
            m, k, x + c * P, H, W, obs_hr + ((size_t)k * C_total + channel_start + c) * P,
            gradient ? gradient + c * P : NULL, frame);
      total += frame;
    }
    free(costs);
    return total;
  } else {
    /* threaded: per-job private gradient buffers, reduced afterwards in frame order */
    double* priv = NULL;
    if (gradient) {
      priv = (double*)calloc((size_t)jobs * P, sizeof(double));
      if (!priv) abort();
    }
#ifdef _OPENMP
#pragma omp parallel for num_threads(num_threads) schedule(dynamic)
#endif
    for (int j = 0; j < jobs; ++j) {
      const int k = j / C, c = j % C;
      costs[j] = sro_data_term_frame_channel(
          m, k, x + c * P, H, W, obs_hr + ((size_t)k * C_total + channel_start + c) * P,
          priv ? priv + (size_t)j * P : NULL, 0.0);
    }
    if (gradient) {
      for (int j = 0; j < jobs; ++j) {
        double* g = gradient + (size_t)(j % C) * P;
        const double* pj = priv + (size_t)j * P;
        for (size_t p = 0; p < P; ++p) g[p] += pj[p];
      }
      free(priv);
    }
  }
  double total = 0.0;
  for (int k = 0; k < N; ++k) {
    double frame = 0.0;
    for (int c = 0; c < C; ++c) frame += costs[k * C + c];
    total += frame;
  }
  free(costs);
  return total;
}

/* ------------------------------------------------------------------------------------------
 * Regularizers.
 * ---------------------------------------------------------------------------------------- */
#define IDX(c, r, col) ((size_t)(c) * P + (size_t)(r) * W + (col)) /* util.cpp:81-89 */

/* tv_regularizer.cpp:21-55 */
static inline double tv_gx(const double* x, int H, int W, size_t P, int c, int r, int col) {
  (void)H;
  if (col >= 0 && col + 1 < W) return x[IDX(c, r, col + 1)] - x[IDX(c, r, col)];
  return 0;
}
static inline double tv_gy(const double* x, int H, int W, size_t P, int c, int r, int col) {
  if (r >= 0 && r + 1 < H) return x[IDX(c, r + 1, col)] - x[IDX(c, r, col)];
  return 0;
}
/* tv_regularizer.cpp:57-70 */
static inline double tv_gz(const double* x, int W, size_t P, int c, int r, int col) {
  return x[IDX(c + 1, r, col)] - x[IDX(c, r, col)];
}
/* tv_regularizer.cpp:72-107: y variation first, then x, then (3-D) z */
static inline double tv_value(const double* x, int H, int W, size_t P, int C, int use3d, int c,
                              int r, int col) {
  const double yv = fabs(tv_gy(x, H, W, P, c, r, col));
  const double xv = fabs(tv_gx(x, H, W, P, c, r, col));
  double tv = yv + xv;
  if (use3d && c + 1 < C) tv += fabs(tv_gz(x, W, P, c, r, col));
  return tv;
}

/* btv_regularizer.cpp:19-46: inclusive window 0..R in both directions, std::pow per tap */
static inline double btv_value(const double* x, int H, int W, size_t P, int c, int r, int col,
                               int R, double decay) {
  double tv = 0.0;
  const size_t index = IDX(c, r, col);
  for (int i = 0; i <= R; ++i) {
    for (int j = 0; j <= R; ++j) {
      const int orow = r + i, ocol = col + j;
      if (orow >= H || ocol >= W) continue;
      const double d = pow(decay, i + j);
      tv += d * fabs(x[index] - x[IDX(c, orow, ocol)]);
    }
  }
  return tv;
}

void sro_reg_apply(int kind, int R, double decay, const double* x, int H, int W, int C,
                   double* values) {
  const size_t P = (size_t)H * W;
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < H; ++r)
      for (int col = 0; col < W; ++col)
        values[IDX(c, r, col)] = (kind == 2) ? btv_value(x, H, W, P, c, r, col, R, decay)
                                             : tv_value(x, H, W, P, C, kind == 1, c, r, col);
}

static inline double sgn_pos(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

void sro_reg_apply_diff(int kind, int R, double decay, const double* x, const double* cst, int H,
                        int W, int C, double* values, double* partials) {
  const size_t P = (size_t)H * W;
  sro_reg_apply(kind, R, decay, x, H, W, C, values);
  for (int c = 0; c < C; ++c) {
    for (int r = 0; r < H; ++r) {
      for (int col = 0; col < W; ++col) {
        const size_t index = IDX(c, r, col);
        double g = 0.0;
        if (kind != 2) {
          /* tv_regularizer.cpp:152-170: self term; note the 3-D z self term is absent */
          double didi = 0.0;
          const double gx = tv_gx(x, H, W, P, c, r, col);
          if (gx < 0.0) didi += 1.0; else if (gx > 0.0) didi -= 1.0;
          const double gy = tv_gy(x, H, W, P, c, r, col);
          if (gy < 0.0) didi += 1.0; else if (gy > 0.0) didi -= 1.0;
          g += 2 * cst[index] * values[index] * didi;
          /* :171-184 left neighbour */
          if (col - 1 >= 0) {
            const size_t li = IDX(c, r, col - 1);
            g += 2 * cst[li] * values[li] * sgn_pos(tv_gx(x, H, W, P, c, r, col - 1));
          }
          /* :185-201 above neighbour */
          if (r - 1 >= 0) {
            const size_t ai = IDX(c, r - 1, col);
            g += 2 * cst[ai] * values[ai] * sgn_pos(tv_gy(x, H, W, P, c, r - 1, col));
          }
          /* :202-220 previous channel (3-D TV) */
          if (kind == 1 && c > 0) {
            const size_t bi = IDX(c - 1, r, col);
            g += 2 * cst[bi] * values[bi] * sgn_pos(tv_gz(x, W, P, c - 1, r, col));
          }
        } else {
          /* btv_regularizer.cpp:113-136: self term, EXCLUSIVE window 0..R-1 */
          double didi = 0.0;
          for (int i = 0; i < R; ++i) {
            for (int j = 0; j < R; ++j) {
              const int orow = r + i, ocol = col + j;
              if (orow >= H || ocol >= W) continue;
              const double diff = x[index] - x[IDX(c, orow, ocol)];
              didi += pow(decay, i + j) * sgn_pos(diff);
            }
          }
          g += 2 * cst[index] * values[index] * didi;
          /* :137-165: pixels whose window covers this one; skips image pixel (0,0) */
          for (int i = 0; i < R; ++i) {
            for (int j = 0; j < R; ++j) {
              const int orow = r - i, ocol = col - j;
              if ((orow == 0 && ocol == 0) || orow < 0 || ocol < 0) continue;
              const size_t oi = IDX(c, orow, ocol);
              const double diff = x[oi] - x[index];
              double didj = 0.0;
              if (diff < 0.0) didj = 1.0; else if (diff > 0.0) didj = -1.0;
              didj *= pow(decay, i + j);
              g += 2 * cst[oi] * values[oi] * didj;
            }
          }
        }
        partials[index] = g;
      }
    }
  }
}

/* ObjectiveIRLSRegularizationTerm::Compute (objective_irls_regularization_term.cpp:10-58) */
double sro_irls_term(int kind, int R, double decay, double lambda, const double* weights,
                     const double* x, int H, int W, int C, double* gradient) {
  if (lambda <= 0.0) return 0.0; /* :15-18 */
  const size_t n = (size_t)H * W * C;
  double* cst = sro_alloc(n);
  double* values = sro_alloc(n);
  double* partials = sro_alloc(n);
  for (size_t i = 0; i < n; ++i) cst[i] = lambda * weights[i]; /* :27-32 */
  sro_reg_apply_diff(kind, R, decay, x, cst, H, W, C, values, partials);
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i) { /* :46-55 */
    const double r = values[i];
    sum += lambda * weights[i] * r * r;
    if (gradient) gradient[i] += partials[i];
  }
  free(cst);
  free(values);
  free(partials);
  return sum;
}

/* irls_map_solver.cpp:35,128-143 */
void sro_reweight(int kind, int R, double decay, const double* x, int H, int W, int C,
                  double* weights) {
  const size_t n = (size_t)H * W * C;
  sro_reg_apply(kind, R, decay, x, H, W, C, weights);
  for (size_t i = 0; i < n; ++i) {
    const double r = weights[i];
    weights[i] = 1.0 / (r > 0.00001 ? r : 0.00001);
  }
}

/* ObjectiveFunction::ComputeAllTerms (objective_function.cpp:5-20) */
double sro_eval(const sro_model* m, const double* x, int H, int W, int C, const double* obs_hr,
                int reg_kind, int R, double decay, double lambda, const double* weights,
                double* gradient, int num_threads) {
  const size_t n = (size_t)H * W * C;
  if (gradient) memset(gradient, 0, n * sizeof(double));
  double sum = 0.0;
  sum += sro_data_term(m, x, H, W, C, obs_hr, C, 0, gradient, num_threads);
  if (weights && lambda > 0.0)
    sum += sro_irls_term(reg_kind, R, decay, lambda, weights, x, H, W, C, gradient);
  return sum;
}
