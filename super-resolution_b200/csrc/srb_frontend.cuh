// srb_frontend.cuh -- the steps either side of the hot path, on the device and
// behind the same C ABI, so that a whole run -- read the cube, reduce its bands, synthesise / load the LR stack,
// initial estimate, solve, score -- needs the host only for file IO and scalars.  Included at the end of srb_api.cu.
//
//   N2  LR stack synthesis: ImageModel::ApplyToImage over every frame (image_model.cpp:76-84, generate_data.cpp:
//       83-127) + AdditiveNoiseModule (additive_noise_module.cpp:19-36: N(0, (sigma/255)^2) per sample).  The
//       reference draws from cv::randn on OpenCV's global, unseeded-by-the-program RNG, so the noise values have
//       no reference to match: parity is statistical, and determinism comes from a counter-based generator
//       (Philox4x32-10 + Box-Muller, pinned against Random123's known-answer vectors in the oracle).
//   N3  initial estimate: cv::resize(INTER_LINEAR) of LR frame 0 (super_resolution.cpp:371-373,
//       image_data.cpp:310-364), and the two scores of src/evaluation (peak_signal_to_noise_ratio.cpp:11-54,
//       structural_similarity.cpp:9-103: global mean / variance / covariance SSIM, not the windowed one).
//   N4  ENVI BSQ float32 reader / writer (hyperspectral_data_loader.cpp:68-118, 120-196, 226-270) and SpectralPCA
//       (spectral_pca.cpp:27-199: cv::PCA on 10*C sub-sampled pixel vectors, project / backProject per pixel).
#pragma once
#include <cstdio>
#include <fstream>
#include <sstream>

namespace srb {

// ---- N3: bilinear resize (cv::resize INTER_LINEAR, CV_64F) -----------------------------------------------
// OpenCV's geometry (resize.cpp, resizeGeneric_ / HResizeLinear / VResizeLinear): destination index d samples
// the source at f = (d + 0.5) * (n_src / n_dst) - 0.5, i = floor(f), weight w = f - i; i < 0 -> (0, w = 0);
// i >= n_src - 1 -> (n_src - 1, w = 0).  Rows are interpolated horizontally first, then vertically, each as
// a * (1 - w) + b * w without contraction.
__device__ __forceinline__ void linear_coef(int d, int n_dst, int n_src, int* i0, int* i1, double* w) {
  const double scale = (double)n_src / (double)n_dst;
  double f = __dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  int i = (int)floor(f);
  f = __dadd_rn(f, -(double)i);
  if (i < 0) { i = 0; f = 0.0; }
  if (i >= n_src - 1) { i = n_src - 1; f = 0.0; }
  *i0 = i;
  *i1 = i + 1 < n_src ? i + 1 : n_src - 1;
  *w = f;
}

// grid: (ceil(W / 32), ceil(H / 8), C); src [C][h][w] with plane stride src_plane -> dst [C][H][W]
__global__ void __launch_bounds__(256)
k_resize_linear(const double* __restrict__ src, size_t src_plane, int h, int w, double* __restrict__ dst, int H, int W) {
  const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y, c = blockIdx.z;
  if (X >= W || Y >= H) return;
  int x0, x1, y0, y1;
  double wx, wy;
  linear_coef(X, W, w, &x0, &x1, &wx);
  linear_coef(Y, H, h, &y0, &y1, &wy);
  const double* __restrict__ s = src + (size_t)c * src_plane;
  const double ax = __dadd_rn(1.0, -wx), ay = __dadd_rn(1.0, -wy);
  const double r0 = __dadd_rn(__dmul_rn(s[(size_t)y0 * w + x0], ax), __dmul_rn(s[(size_t)y0 * w + x1], wx));
  const double r1 = __dadd_rn(__dmul_rn(s[(size_t)y1 * w + x0], ax), __dmul_rn(s[(size_t)y1 * w + x1], wx));
  dst[((size_t)c * H + Y) * W + X] = __dadd_rn(__dmul_rn(r0, ay), __dmul_rn(r1, wy));
}

// ---- N3: PSNR / SSIM sums ----------------------------------------------------------------------------------
// pass 1: partial sums of a and b;  pass 2 (means known): partial sums of (a-ma)^2, (b-mb)^2, (a-ma)(b-mb),
// (b-a)^2.  One slot per block and quantity, reduced in fixed order by k_score_finish (deterministic).
__global__ void __launch_bounds__(256)
k_score_sums(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* __restrict__ part) {
  double sa = 0.0, sb = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    sa += a[i];
    sb += b[i];
  }
  sa = block_sum(sa);
  sb = block_sum(sb);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = sa;
    part[gridDim.x + blockIdx.x] = sb;
  }
}
__global__ void __launch_bounds__(256)
k_score_moments(const double* __restrict__ a, const double* __restrict__ b, size_t n, const double* __restrict__ means,
                double* __restrict__ part) {
  const double ma = means[0], mb = means[1];
  double va = 0.0, vb = 0.0, cab = 0.0, sd = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const double x = a[i], y = b[i];
    const double da = x - ma, db = y - mb, d = y - x;
    va = fma(da, da, va);
    vb = fma(db, db, vb);
    cab = fma(da, db, cab);
    sd = fma(d, d, sd);
  }
  va = block_sum(va);
  vb = block_sum(vb);
  cab = block_sum(cab);
  sd = block_sum(sd);
  if (threadIdx.x == 0) {
    part[0 * gridDim.x + blockIdx.x] = va;
    part[1 * gridDim.x + blockIdx.x] = vb;
    part[2 * gridDim.x + blockIdx.x] = cab;
    part[3 * gridDim.x + blockIdx.x] = sd;
  }
}
// out[q] = sum(part[q][0 .. nb)) * scale for q < nq; grid: nq blocks
__global__ void __launch_bounds__(256)
k_score_finish(const double* __restrict__ part, int nb, double scale, double* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += part[(size_t)blockIdx.x * nb + i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = acc * scale;
}

// ---- N2: additive Gaussian noise -----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123): counter = (index of the 4-sample group, stream), key = seed.
__host__ __device__ inline void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k1;
    const unsigned n3 = (unsigned)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// data[i] += sigma * z_i, z_i ~ N(0, 1): samples 4g .. 4g+3 come from counter (g, stream) as two Box-Muller pairs
// with u = (r + 0.5) * 2^-32 in (0, 1).
__global__ void __launch_bounds__(256)
k_add_noise(double* __restrict__ data, size_t n, double sigma, unsigned long long seed, unsigned long long stream) {
  const size_t groups = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
    unsigned c[4] = {(unsigned)g, (unsigned)(g >> 32), (unsigned)stream, (unsigned)(stream >> 32)};
    philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    double z[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double u1 = ((double)c[2 * p] + 0.5) * 2.3283064365386963e-10;
      const double u2 = ((double)c[2 * p + 1] + 0.5) * 2.3283064365386963e-10;
      const double r = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      z[2 * p] = r * cs;
      z[2 * p + 1] = r * sn;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * g + j < n) data[4 * g + j] = fma(sigma, z[j], data[4 * g + j]);
  }
}

// ---- N4: ENVI float32 BSQ -> planar doubles; SpectralPCA projections -----------------------------------------
// raw: the bytes of rows [r0, r1) x ALL columns of the selected bands, band after band; out [bands][r1-r0][c1-c0]
__global__ void __launch_bounds__(256)
k_envi_convert(const unsigned* __restrict__ raw, int rows, int cols_file, int c0, int cols, int swap,
               double* __restrict__ out, size_t total) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const size_t plane = (size_t)rows * cols;
    const size_t b = i / plane, rem = i - b * plane;
    const int r = (int)(rem / cols), c = (int)(rem - (size_t)r * cols);
    unsigned v = raw[(b * rows + r) * (size_t)cols_file + c0 + c];
    if (swap) v = __byte_perm(v, 0, 0x0123);
    out[i] = (double)__uint_as_float(v);
  }
}

// cv::PCA::project per pixel (spectral_pca.cpp:137-141): out[j][p] = sum_c (in[c][p] - mean[c]) * E[j][c];
// backProject: out[c][p] = sum_j in[j][p] * E[j][c] + mean[c].  Thread = pixel (planar layout: coalesced per
// band); the basis is read through the constant / L1 path.  Up to 32 outputs per pass are kept in registers.
template <bool FORWARD>
__global__ void __launch_bounds__(256)
k_pca_convert(const double* __restrict__ in, size_t P, int C, int Kc, const double* __restrict__ mean,
              const double* __restrict__ E, double* __restrict__ out, int o0, int o1) {
  const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  double acc[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) acc[o] = 0.0;
  const int n_in = FORWARD ? C : Kc;
  for (int i = 0; i < n_in; ++i) {
    const double v = FORWARD ? in[(size_t)i * P + p] - mean[i] : in[(size_t)i * P + p];
#pragma unroll
    for (int o = 0; o < 32; ++o)
      if (o0 + o < o1) acc[o] = fma(v, FORWARD ? E[(size_t)(o0 + o) * C + i] : E[(size_t)i * C + (o0 + o)], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 32; ++o)
    if (o0 + o < o1) out[(size_t)(o0 + o) * P + p] = FORWARD ? acc[o] : acc[o] + mean[o0 + o];
}

inline int stream_blocks(const srb_ctx* c, size_t n) {
  const size_t b = (n + 255) / 256;
  const size_t cap = (size_t)c->num_sms * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

struct DevBuf {  // RAII device scratch
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  template <class T>
  T* as() { return (T*)p; }
};

// Scores of `image` against `truth` (device buffers, n doubles each) -> host scalars.
inline srb_status score_dev(srb_ctx* c, const double* d_img, const double* d_truth, size_t n, double k1, double k2,
                            double image_scale, double* psnr, double* ssim) {
  const int nb = stream_blocks(c, n);
  DevBuf part, res;
  SRB_CUDA_CHECK(c, cudaMalloc(&part.p, (size_t)4 * nb * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&res.p, 8 * sizeof(double)));
  double* d_part = part.as<double>();
  double* d_res = res.as<double>();
  const double inv_n = 1.0 / (double)n;
  k_score_sums<<<nb, 256, 0, c->stream>>>(d_truth, d_img, n, d_part);
  k_score_finish<<<2, 256, 0, c->stream>>>(d_part, nb, inv_n, d_res);            // means: truth, image
  k_score_moments<<<nb, 256, 0, c->stream>>>(d_truth, d_img, n, d_res, d_part);
  k_score_finish<<<4, 256, 0, c->stream>>>(d_part, nb, inv_n, d_res + 2);        // var truth, var image, cov, mse
  c->timing.kernel_launches += 4;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  double h[6];
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(h, d_res, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  const double mean_t = h[0], mean_i = h[1], var_t = h[2], var_i = h[3], cov = h[4], mse = h[5];
  if (psnr) *psnr = 20.0 * log10(1.0) - 10.0 * log10(mse);  // peak_signal_to_noise_ratio.cpp:44-52 (max value 1.0)
  if (ssim) {  // structural_similarity.cpp:60-101
    double c1 = k1 * image_scale, c2 = k2 * image_scale;
    c1 = c1 * c1;
    c2 = c2 * c2;
    const double n1 = 2 * mean_t * mean_i + c1, n2 = 2 * cov + c2;
    const double d1 = mean_t * mean_t + mean_i * mean_i + c1, d2 = var_t + var_i + c2;
    *ssim = (n1 * n2) / (d1 * d2);
  }
  return SRB_OK;
}

// Symmetric eigen-decomposition by cyclic Jacobi rotations (what cv::eigen does for the C x C covariance of
// cv::PCA): A is overwritten, V's ROWS are the eigenvectors, sorted by descending eigenvalue.
inline void jacobi_eigen(std::vector<double>& A, int n, std::vector<double>& eval, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {  // columns p, q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = cs * akp - sn * akq;
          A[(size_t)k * n + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = cs * apk - sn * aqk;
          A[(size_t)q * n + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < n; ++k) {  // eigenvectors (rows of V)
          const double vpk = V[(size_t)p * n + k], vqk = V[(size_t)q * n + k];
          V[(size_t)p * n + k] = cs * vpk - sn * vqk;
          V[(size_t)q * n + k] = sn * vpk + cs * vqk;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] > A[(size_t)b * n + b]; });
  std::vector<double> Vs((size_t)n * n);
  eval.resize(n);
  for (int i = 0; i < n; ++i) {
    eval[i] = A[(size_t)order[i] * n + order[i]];
    for (int k = 0; k < n; ++k) Vs[(size_t)i * n + k] = V[(size_t)order[i] * n + k];
  }
  V.swap(Vs);
}

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

}  // namespace srb

struct srb_pca {
  int num_bands = 0;        // C: spectral bands
  int num_components = 0;   // k: retained components
  std::vector<double> mean, eigenvectors, eigenvalues;  // [C], [k][C], [k]
};

extern "C" {

// ---- N3 ------------------------------------------------------------------------------------------------------
srb_status srb_resize_linear(srb_ctx* c, const double* src_host, int C, int h, int w, int H, int W, double* dst_host) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!src_host || !dst_host) return c->fail(SRB_ERR_INVALID, "null buffer");
  // image_data.cpp:318-320: CHECK_GT on the new size
  if (C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return c->fail(SRB_ERR_INVALID, "images must have a positive size");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  DevBuf in, out;
  const size_t n_in = (size_t)C * h * w, n_out = (size_t)C * H * W;
  SRB_CUDA_CHECK(c, cudaMalloc(&in.p, n_in * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&out.p, n_out * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(in.p, src_host, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_resize_linear<<<grid2d(W, H, C), dim3(32, 8), 0, c->stream>>>(in.as<double>(), (size_t)h * w, h, w, out.as<double>(), H, W);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(dst_host, out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_initial_estimate_dev(srb_ctx* c, int frame, double* x_dev_out) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev_out) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  if (frame < 0 || frame >= c->g.N) return c->fail(SRB_ERR_INVALID, "frame index out of range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const Geometry& G = c->g;
  const double* src = c->d_y + ((size_t)frame * G.Ct + c->c0) * c->p;
  k_resize_linear<<<grid2d(G.W, G.H, c->Ca()), dim3(32, 8), 0, c->stream>>>(src, c->p, G.h, G.w, x_dev_out, G.H, G.W);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

srb_status srb_initial_estimate(srb_ctx* c, int frame, double* x_host_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host_out) return c->fail(SRB_ERR_INVALID, "null buffer");
  srb_status st = srb_initial_estimate_dev(c, frame, c->d_x);
  if (st != SRB_OK) return st;
  c->x_resident = true;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(x_host_out, c->d_x, c->n_active() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_score_dev(srb_ctx* c, const double* image_dev, const double* truth_dev, unsigned long long n, double k1,
                         double k2, double image_scale, double* psnr, double* ssim) {
  if (!c) return SRB_ERR_INVALID;
  if (!image_dev || !truth_dev || n == 0) return c->fail(SRB_ERR_INVALID, "null or empty image");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  return srb::score_dev(c, image_dev, truth_dev, (size_t)n, k1, k2, image_scale, psnr, ssim);
}

srb_status srb_score(srb_ctx* c, const double* image_host, const double* truth_host, unsigned long long n, double k1,
                     double k2, double image_scale, double* psnr, double* ssim) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!image_host || !truth_host || n == 0) return c->fail(SRB_ERR_INVALID, "null or empty image");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  DevBuf a, b;
  SRB_CUDA_CHECK(c, cudaMalloc(&a.p, (size_t)n * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&b.p, (size_t)n * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(a.p, image_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(b.p, truth_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return score_dev(c, a.as<double>(), b.as<double>(), (size_t)n, k1, k2, image_scale, psnr, ssim);
}

// ---- N2 ------------------------------------------------------------------------------------------------------
srb_status srb_add_noise_dev(srb_ctx* c, double* data_dev, unsigned long long n, double sigma, unsigned long long seed,
                             unsigned long long stream_id) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!data_dev) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (!(sigma > 0.0)) return c->fail(SRB_ERR_INVALID, "noise sigma must be positive");  // additive_noise_module.cpp:15-17
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  // additive_noise_module.cpp:26-27: pixels are in [0, 1], sigma is given on the 0..255 scale
  k_add_noise<<<stream_blocks(c, ((size_t)n + 3) / 4), 256, 0, c->stream>>>(data_dev, (size_t)n, sigma / 255.0, seed, stream_id);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

srb_status srb_add_noise(srb_ctx* c, double* data_host, unsigned long long n, double sigma, unsigned long long seed,
                         unsigned long long stream_id) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!data_host) return c->fail(SRB_ERR_INVALID, "null buffer");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  DevBuf d;
  SRB_CUDA_CHECK(c, cudaMalloc(&d.p, (size_t)n * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(d.p, data_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  srb_status st = srb_add_noise_dev(c, d.as<double>(), n, sigma, seed, stream_id);
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(data_host, d.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_generate_observations(srb_ctx* c, const double* hr_host, const double* hr_dev, double noise_sigma,
                                     unsigned long long seed, double* lr_out_host, int keep_as_observations) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if ((hr_host == nullptr) == (hr_dev == nullptr)) return c->fail(SRB_ERR_INVALID, "give the HR image on the host or on the device");
  if (noise_sigma < 0.0) return c->fail(SRB_ERR_INVALID, "noise sigma must not be negative");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const Geometry& G = c->g;
  const size_t n_hr = (size_t)G.Ct * c->P, n_lr = (size_t)G.N * G.Ct * c->p;
  const double* src = hr_dev;
  if (hr_host) {
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, hr_host, n_hr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->x_resident = false;
    src = c->d_x;
  }
  double* dst = keep_as_observations ? c->d_y : nullptr;
  if (!dst) {
    srb_status st = dev_alloc(c, &c->d_pooled, n_lr);
    if (st != SRB_OK) return st;
    dst = c->d_pooled;
  }
  GenericParams P = make_params(c, false);
  P.Ca = G.Ct;  // every channel, whatever the active channel range is
  P.c0 = 0;
  if (n_lr > 0) {
    k_forward_generic<0><<<grid2d(G.w, G.h, G.N * G.Ct), dim3(32, 8), 0, c->stream>>>(P, src, nullptr, dst, nullptr);
    c->timing.kernel_launches += 1;
    SRB_CUDA_CHECK(c, cudaGetLastError());
    if (noise_sigma > 0.0) {  // image_model.cpp:76-84 applies the operators in order: the noise module comes last
      srb_status st = srb_add_noise_dev(c, dst, n_lr, noise_sigma, seed, 0);
      if (st != SRB_OK) return st;
    }
  }
  if (keep_as_observations) {
    srb_status st = fused_observations_changed(c);
    if (st != SRB_OK) return st;
    c->have_obs = true;
  }
  if (lr_out_host)
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(lr_out_host, dst, n_lr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

// ---- N4: ENVI ------------------------------------------------------------------------------------------------
srb_status srb_envi_read_header(const char* header_path, srb_envi_header* out) {
  if (!header_path || !out) return SRB_ERR_INVALID;
  std::ifstream fin(header_path);
  if (!fin.is_open()) return SRB_ERR_INVALID;
  // HSIBinaryDataParameters::ReadHeaderFromFile (hyperspectral_data_loader.cpp:226-270): "key = value" lines,
  // split at the first '=', both sides trimmed (config_reader.cpp:16-36); unknown interleave / data type fall
  // back to bsq / float with a warning
  srb_envi_header h{};
  h.interleave_bsq = 1;
  h.data_type = 4;
  std::string line;
  while (std::getline(fin, line)) {
    if (line.find("#") == 0) continue;
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    const std::string key = srb::trim(line.substr(0, eq)), value = srb::trim(line.substr(eq + 1));
    if (key == "interleave") h.interleave_bsq = value == "bsq" ? 1 : 0;
    else if (key == "data type") h.data_type = atoi(value.c_str());
    else if (key == "byte order") h.big_endian = value == "1" ? 1 : 0;
    else if (key == "header offset") h.header_offset = atoi(value.c_str());
    else if (key == "samples") h.num_data_rows = atoi(value.c_str());   // (sic: the reference maps samples -> rows)
    else if (key == "lines") h.num_data_cols = atoi(value.c_str());
    else if (key == "bands") h.num_data_bands = atoi(value.c_str());
  }
  *out = h;
  return SRB_OK;
}

static srb_status envi_read_impl(srb_ctx* c, const char* path, const srb_envi_header* hd, int r0, int r1, int c0, int c1,
                                 int b0, int b1, double* out_host, double* out_dev) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!path || !hd || (!out_host && !out_dev)) return c->fail(SRB_ERR_INVALID, "null argument");
  if (hd->num_data_rows <= 0 || hd->num_data_cols <= 0 || hd->num_data_bands <= 0 || hd->header_offset < 0)
    return c->fail(SRB_ERR_INVALID, "ENVI data size must be positive and the header offset non-negative");
  if (!hd->interleave_bsq || hd->data_type != 4) return c->fail(SRB_ERR_INVALID, "only float32 BSQ ENVI data is supported");
  if (r0 < 0 || r1 > hd->num_data_rows || r1 <= r0 || c0 < 0 || c1 > hd->num_data_cols || c1 <= c0 || b0 < 0 ||
      b1 > hd->num_data_bands || b1 <= b0)
    return c->fail(SRB_ERR_INVALID, "ENVI data range outside the file's size");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  FILE* f = fopen(path, "rb");
  if (!f) return c->fail(SRB_ERR_INVALID, std::string("file '") + path + "' could not be opened for reading");
  const int rows = r1 - r0, cols = c1 - c0, bands = b1 - b0;
  const size_t row_bytes = (size_t)hd->num_data_cols * 4, band_bytes = (size_t)rows * row_bytes;
  // whole rows of the selected range are read (one contiguous run per band) into pinned memory, band by band,
  // and go to the device while the next band is being read
  float* h_raw = nullptr;
  DevBuf raw, out;
  if (cudaMallocHost((void**)&h_raw, (size_t)bands * band_bytes) != cudaSuccess) {
    fclose(f);
    (void)cudaGetLastError();
    return c->fail(SRB_ERR_NOMEM, "cudaMallocHost failed (ENVI staging buffer)");
  }
  srb_status st = SRB_OK;
  if (cudaMalloc(&raw.p, (size_t)bands * band_bytes) != cudaSuccess) st = c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (ENVI raw data)");
  const size_t total = (size_t)bands * rows * cols;
  double* d_out = out_dev;
  if (st == SRB_OK && !d_out) {
    if (cudaMalloc(&out.p, total * sizeof(double)) != cudaSuccess) st = c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (ENVI image)");
    d_out = out.as<double>();
  }
  const size_t plane = (size_t)hd->num_data_rows * hd->num_data_cols;
  for (int b = 0; st == SRB_OK && b < bands; ++b) {
    // the header offset counts BYTES (the ENVI convention); element (band, row, col) sits at
    // offset + 4 * (band * rows * cols + row * cols + col)   (hyperspectral_data_loader.cpp:88-101)
    const long long pos = (long long)hd->header_offset + 4LL * ((long long)(b0 + b) * (long long)plane + (long long)r0 * hd->num_data_cols);
    char* dstp = (char*)h_raw + (size_t)b * band_bytes;
    if (fseek(f, (long)pos, SEEK_SET) != 0 || fread(dstp, 1, band_bytes, f) != band_bytes)
      st = c->fail(SRB_ERR_INVALID, "ENVI file is shorter than its header says");
    else if (cudaMemcpyAsync((char*)raw.p + (size_t)b * band_bytes, dstp, band_bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      st = c->fail(SRB_ERR_CUDA, "cudaMemcpyAsync failed (ENVI raw data)");
  }
  fclose(f);
  if (st == SRB_OK) {
    const unsigned one = 1;
    const bool machine_big = *(const unsigned char*)&one != 1;  // IsMachineBigEndian (:48-63)
    const int swap = (hd->big_endian != 0) != machine_big ? 1 : 0;
    k_envi_convert<<<stream_blocks(c, total), 256, 0, c->stream>>>(raw.as<unsigned>(), rows, hd->num_data_cols, c0, cols, swap, d_out, total);
    c->timing.kernel_launches += 1;
    if (cudaGetLastError() != cudaSuccess) st = c->fail(SRB_ERR_CUDA, "k_envi_convert launch failed");
    if (st == SRB_OK && out_host &&
        cudaMemcpyAsync(out_host, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      st = c->fail(SRB_ERR_CUDA, "cudaMemcpyAsync failed (ENVI image)");
  }
  cudaStreamSynchronize(c->stream);
  cudaFreeHost(h_raw);
  return st;
}

srb_status srb_envi_read(srb_ctx* c, const char* path, const srb_envi_header* hd, int r0, int r1, int c0, int c1, int b0,
                         int b1, double* out_host) {
  return envi_read_impl(c, path, hd, r0, r1, c0, c1, b0, b1, out_host, nullptr);
}
srb_status srb_envi_read_dev(srb_ctx* c, const char* path, const srb_envi_header* hd, int r0, int r1, int c0, int c1,
                             int b0, int b1, double* out_dev) {
  return envi_read_impl(c, path, hd, r0, r1, c0, c1, b0, b1, nullptr, out_dev);
}

srb_status srb_envi_write(const char* path, const double* image_host, int bands, int rows, int cols) {
  if (!path || !image_host || bands <= 0 || rows <= 0 || cols <= 0) return SRB_ERR_INVALID;
  // WriteBinaryFileBSQ<float> (hyperspectral_data_loader.cpp:120-196): float32, machine byte order, then the
  // .hdr and the .config that lets LoadImageFromENVIFile read the file back
  FILE* f = fopen(path, "wb");
  if (!f) return SRB_ERR_INVALID;
  std::vector<float> row((size_t)cols);
  bool ok = true;
  for (size_t r = 0; ok && r < (size_t)bands * rows; ++r) {
    for (int q = 0; q < cols; ++q) row[q] = (float)image_host[r * cols + q];
    ok = fwrite(row.data(), 4, cols, f) == (size_t)cols;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return SRB_ERR_INVALID;
  std::ofstream hdr(std::string(path) + ".hdr");
  if (!hdr.is_open()) return SRB_ERR_INVALID;
  hdr << "ENVI\ndescription = {File generated by HyperspectralDataLoader.}\n"
      << "samples = " << rows << "\nlines = " << cols << "\nbands = " << bands << "\n"
      << "header offset = 0\nfile type = ENVI Standard\ndata type = 4\ninterleave = bsq\nbyte order = 0\n";
  hdr.close();
  std::ofstream cfg(std::string(path) + ".config");
  if (!cfg.is_open()) return SRB_ERR_INVALID;
  cfg << "# Configuration file for reading '" << path << "', generated by HyperspectralDataLoader.\n"
      << "file " << path << "\ninterleave bsq\ndata_type float\nbig_endian false\nheader_offset 0\n"
      << "num_data_rows " << rows << "\nnum_data_cols " << cols << "\nnum_data_bands " << bands << "\n"
      << "start_row 0\nend_row " << rows << "\nstart_col 0\nend_col " << cols << "\nstart_band 0\nend_band " << bands << "\n";
  cfg.close();
  return SRB_OK;
}

// ---- N4: SpectralPCA -----------------------------------------------------------------------------------------
srb_status srb_pca_create(const double* const* images_host, int num_images, int num_bands, unsigned long long num_pixels,
                          int num_pca_bands, double retained_variance, srb_pca** out) {
  using namespace srb;
  if (!out) return SRB_ERR_INVALID;
  *out = nullptr;
  // spectral_pca.cpp:27-33: at least one image, at least one channel
  if (!images_host || num_images < 1 || num_bands < 1 || num_pixels < 1) return SRB_ERR_INVALID;
  if (num_pca_bands < 0 || num_pca_bands > num_bands || (num_pca_bands == 0 && !(retained_variance > 0.0 && retained_variance <= 1.0)))
    return SRB_ERR_INVALID;
  // GetPCAInputData (spectral_pca.cpp:27-96): 10 * C samples in all, taken every num_pixels / per_image pixels
  const int C = num_bands;
  const long long P = (long long)num_pixels;
  long long per_image = (long long)C * 10 / num_images;
  if (per_image > P) per_image = P;
  if (per_image < 1) return SRB_ERR_INVALID;  // (the reference divides by zero here)
  const long long skip = P / per_image;
  const long long count = per_image * num_images;
  std::vector<double> data((size_t)count * C);
  for (int im = 0; im < num_images; ++im) {
    if (!images_host[im]) return SRB_ERR_INVALID;
    for (int ch = 0; ch < C; ++ch)
      for (long long sm = 0; sm < per_image; ++sm)
        data[(size_t)(im * per_image + sm) * C + ch] = images_host[im][(size_t)ch * P + (size_t)(sm * skip)];
  }
  // cv::PCA (DATA_AS_ROW): mean over the samples, covariance scaled by 1 / count, eigenvectors by descending
  // eigenvalue
  srb_pca* p = new (std::nothrow) srb_pca();
  if (!p) return SRB_ERR_NOMEM;
  p->num_bands = C;
  p->mean.assign(C, 0.0);
  for (long long r = 0; r < count; ++r)
    for (int ch = 0; ch < C; ++ch) p->mean[ch] += data[(size_t)r * C + ch];
  for (int ch = 0; ch < C; ++ch) p->mean[ch] /= (double)count;
  std::vector<double> cov((size_t)C * C, 0.0);
  for (long long r = 0; r < count; ++r)
    for (int i = 0; i < C; ++i) {
      const double di = data[(size_t)r * C + i] - p->mean[i];
      for (int j = i; j < C; ++j) cov[(size_t)i * C + j] += di * (data[(size_t)r * C + j] - p->mean[j]);
    }
  for (int i = 0; i < C; ++i)
    for (int j = i; j < C; ++j) {
      cov[(size_t)i * C + j] /= (double)count;
      cov[(size_t)j * C + i] = cov[(size_t)i * C + j];
    }
  std::vector<double> eval, V;
  jacobi_eigen(cov, C, eval, V);
  int k = num_pca_bands;
  if (k == 0) {
    // cv::PCA with retainedVariance (OpenCV pca.cpp, computeCumulativeEnergy): L = the first index whose
    // cumulative energy EXCEEDS the fraction -- the component that crosses it is not kept -- then max(2, L)
    double total = 0.0;
    for (int i = 0; i < C; ++i) total += eval[i];
    double acc = 0.0;
    int L = 0;
    for (; L < C; ++L) {
      acc += eval[L];
      if (acc / total > retained_variance) break;
    }
    k = std::min(std::max(L, 2), C);
  }
  p->num_components = k;
  p->eigenvalues.assign(eval.begin(), eval.begin() + k);
  p->eigenvectors.assign(V.begin(), V.begin() + (size_t)k * C);
  *out = p;
  return SRB_OK;
}

void srb_pca_destroy(srb_pca* p) { delete p; }
int srb_pca_num_components(const srb_pca* p) { return p ? p->num_components : 0; }
int srb_pca_num_bands(const srb_pca* p) { return p ? p->num_bands : 0; }
srb_status srb_pca_get(const srb_pca* p, double* mean_out, double* eigenvectors_out, double* eigenvalues_out) {
  if (!p) return SRB_ERR_INVALID;
  if (mean_out) std::copy(p->mean.begin(), p->mean.end(), mean_out);
  if (eigenvectors_out) std::copy(p->eigenvectors.begin(), p->eigenvectors.end(), eigenvectors_out);
  if (eigenvalues_out) std::copy(p->eigenvalues.begin(), p->eigenvalues.end(), eigenvalues_out);
  return SRB_OK;
}

static srb_status pca_convert(srb_ctx* c, const srb_pca* p, const double* in_host, unsigned long long num_pixels,
                              double* out_host, bool forward) {
  using namespace srb;
  if (!c) return SRB_ERR_INVALID;
  if (!p || !in_host || !out_host || num_pixels == 0) return c->fail(SRB_ERR_INVALID, "null or empty argument");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const int C = p->num_bands, K = p->num_components;
  const size_t P = (size_t)num_pixels;
  const int n_in = forward ? C : K, n_out = forward ? K : C;
  DevBuf in, out, mean, ev;
  SRB_CUDA_CHECK(c, cudaMalloc(&in.p, (size_t)n_in * P * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&out.p, (size_t)n_out * P * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&mean.p, (size_t)C * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&ev.p, (size_t)K * C * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(in.p, in_host, (size_t)n_in * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(mean.p, p->mean.data(), (size_t)C * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(ev.p, p->eigenvectors.data(), (size_t)K * C * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int blocks = (int)((P + 255) / 256);
  for (int o0 = 0; o0 < n_out; o0 += 32) {
    const int o1 = std::min(n_out, o0 + 32);
    if (forward)
      k_pca_convert<true><<<blocks, 256, 0, c->stream>>>(in.as<double>(), P, C, K, mean.as<double>(), ev.as<double>(), out.as<double>(), o0, o1);
    else
      k_pca_convert<false><<<blocks, 256, 0, c->stream>>>(in.as<double>(), P, C, K, mean.as<double>(), ev.as<double>(), out.as<double>(), o0, o1);
    c->timing.kernel_launches += 1;
  }
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(out_host, out.p, (size_t)n_out * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_pca_project(srb_ctx* c, const srb_pca* p, const double* image_host, unsigned long long num_pixels,
                           double* pca_image_out_host) {
  return pca_convert(c, p, image_host, num_pixels, pca_image_out_host, true);
}
srb_status srb_pca_reconstruct(srb_ctx* c, const srb_pca* p, const double* pca_image_host, unsigned long long num_pixels,
                               double* image_out_host) {
  return pca_convert(c, p, pca_image_host, num_pixels, image_out_host, false);
}

}  // extern "C"
