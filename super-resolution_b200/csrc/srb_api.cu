// srb_api.cu -- C ABI (include/srb200.h) of the B200-native MAP super-resolution gradient engine.
//
// Host-side orchestration only: geometry tables, buffer management, kernel launches, timing.
// All arithmetic of the hot path runs in the CUDA kernels of srb_kernels_*.cuh; there is no CPU
// fallback anywhere in this library.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "srb_common.cuh"
#include "srb_geometry.h"
#include "srb_kernels_generic.cuh"
#include "srb_kernels_reg.cuh"
#include "srb_kernels_tile.cuh"
#include "srb_kernels_band.cuh"
#include "srb_kernels_peer.cuh"
#include "srb_tile_plan.cuh"

using namespace srb;

namespace {

template <class T>
srb_status dev_alloc(srb_ctx* ctx, T** ptr, size_t count) {
  if (*ptr) return SRB_OK;
  cudaError_t e = cudaMalloc((void**)ptr, (count ? count : 1) * sizeof(T));
  if (e != cudaSuccess) {
    *ptr = nullptr;
    (void)cudaGetLastError();
    return ctx->fail(SRB_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  }
  return SRB_OK;
}
template <class T>
void dev_free(T** ptr) {
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
}

struct TempBuf {  // RAII device scratch for the one-off forward / transpose entry points
  void* p = nullptr;
  ~TempBuf() {
    if (p) cudaFree(p);
  }
  template <class T>
  T* get() { return (T*)p; }
};

inline dim3 grid2d(int W, int H, int Z) { return dim3((W + 31) / 32, (H + 7) / 8, Z); }
inline int fill_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b ? b : 1));
}

GenericParams make_params(const srb_ctx* c, bool transpose_warp) {
  GenericParams P;
  P.H = c->g.H; P.W = c->g.W; P.h = c->g.h; P.w = c->g.w; P.s = c->g.s; P.K = c->g.K; P.hk = c->g.hk;
  P.N = c->g.N; P.Ca = c->Ca(); P.Ct = c->g.Ct; P.c0 = c->c0;
  P.src_r = c->d_src_r; P.src_c = c->d_src_c; P.psf = c->d_psf;
  P.rowY = transpose_warp ? c->d_rowY_tr : c->d_rowY_fwd;
  P.nX = transpose_warp ? c->d_nX_tr : c->d_nX_fwd;
  return P;
}
RegParams make_reg_params(const srb_ctx* c, int C) {
  RegParams R;
  R.H = c->g.H; R.W = c->g.W; R.C = C; R.kind = c->reg_kind; R.R = c->btv_R; R.decay = c->d_decay;
  return R;
}

srb_status ensure_partials(srb_ctx* c, size_t n) {
  if (n <= c->partial_capacity) return SRB_OK;
  dev_free(&c->d_partial);
  srb_status st = dev_alloc(c, &c->d_partial, n);
  if (st == SRB_OK) c->partial_capacity = n;
  return st;
}

bool reg_active(const srb_ctx* c) { return c->reg_kind != SRB_REG_NONE && c->lambda > 0.0; }

int resolve_path(const srb_ctx* c) {
  if (c->path == SRB_PATH_REFERENCE_ORDER) return SRB_PATH_REFERENCE_ORDER;
  return fused_supported(c) ? SRB_PATH_FUSED : SRB_PATH_REFERENCE_ORDER;
}

// ---- evaluation core ---------------------------------------------------------------------------
// ObjectiveFunction::ComputeAllTerms on device buffers.  d_g may be NULL (cost only).  On return
// (stream-ordered) c->d_cost[0..2] hold data cost, regularization cost and their sum; if `tail` is
// non-NULL the sum is also written there.
srb_status eval_core(srb_ctx* c, const double* d_x, double* d_g, double* tail, bool data_term,
                     bool reg_term, bool accumulate) {
  if (data_term && !c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  const Geometry& G = c->g;
  const int Ca = c->Ca();
  const bool do_reg = reg_term && reg_active(c) && c->reg_row1 > c->reg_row0;

  bool reg_done = false;
  if (resolve_path(c) == SRB_PATH_FUSED && data_term && !accumulate) {
    // the tile kernel's finishing launch also closes the cost when nothing follows it: the regularization term
    // is evaluated inside the tile kernel (2-D TV) or by a tiled kernel right behind it (BTV, 3-D TV)
    const bool last = !do_reg || fused_reg_covered(c);
    srb_status st = fused_eval(c, d_x, d_g, do_reg, last ? tail : nullptr, &reg_done);
    if (st != SRB_OK) return st;
    if (last) {
      SRB_CUDA_CHECK(c, cudaGetLastError());
      c->timing.num_evals += 1;
      return SRB_OK;
    }
  } else if (data_term) {
    SRB_CUDA_CHECK(c, cudaMemsetAsync(c->d_cost, 0, 4 * sizeof(double), c->stream));
    srb_status st = dev_alloc(c, &c->d_pooled, (size_t)G.N * G.Ct * c->p);
    if (st != SRB_OK) return st;
    const dim3 grid = grid2d(G.w, G.h, G.N * Ca);
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if ((st = ensure_partials(c, nblocks)) != SRB_OK) return st;
    if (c->strict_cost && (st = dev_alloc(c, &c->d_resid, (size_t)G.N * G.Ct * c->p)) != SRB_OK) return st;
    if (c->strict_cost && (st = ensure_partials(c, nblocks + (size_t)G.N)) != SRB_OK) return st;
    k_forward_generic<1><<<grid, dim3(32, 8), 0, c->stream>>>(make_params(c, false), d_x, c->d_y,
                                                             c->d_pooled, c->d_partial, c->strict_cost ? c->d_resid : nullptr);
    if (c->strict_cost) {
      // the reference's summation order (objective_data_term.cpp:36-50, 104-114): bit-identical cost
      k_strict_data_cost<<<(G.N + 31) / 32, 32, 0, c->stream>>>(c->d_resid, G.N, Ca, G.H, G.W, G.h, G.w, G.s,
                                                               c->d_partial + nblocks, nullptr);
      k_strict_sum_frames<<<1, 1, 0, c->stream>>>(c->d_partial + nblocks, G.N, c->d_cost);
      c->timing.kernel_launches += 1;
    } else {
      k_reduce_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, nblocks, c->d_cost, 0);
    }
    c->timing.kernel_launches += 2;
    if (d_g) {
      k_adjoint_generic<<<grid2d(G.W, G.H, Ca), dim3(32, 8), 0, c->stream>>>(
          make_params(c, true), c->d_pooled, d_g, 2.0, accumulate ? 1 : 0);
      c->timing.kernel_launches += 1;
    }
  } else {
    SRB_CUDA_CHECK(c, cudaMemsetAsync(c->d_cost, 0, 4 * sizeof(double), c->stream));
    if (d_g && !accumulate)
      SRB_CUDA_CHECK(c, cudaMemsetAsync(d_g, 0, c->n_active() * sizeof(double), c->stream));
  }
  if (do_reg && !reg_done) {
    srb_status st = dev_alloc(c, &c->d_vals, (size_t)G.Ct * c->P);
    if (st != SRB_OK) return st;
    const RegParams R = make_reg_params(c, Ca);
    k_reg_values<0><<<grid2d(G.W, G.H, Ca), dim3(32, 8), 0, c->stream>>>(R, d_x, c->d_vals);
    const int rows = c->reg_row1 - c->reg_row0;
    const dim3 grid = grid2d(G.W, rows, Ca);
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if ((st = ensure_partials(c, nblocks)) != SRB_OK) return st;
    k_reg_partials<1><<<grid, dim3(32, 8), 0, c->stream>>>(R, d_x, c->d_vals, c->d_w, c->lambda,
                                                          c->reg_row0, c->reg_row1, d_g, c->d_partial);
    if (c->strict_cost && c->reg_row0 == 0 && c->reg_row1 == G.H)
      k_strict_reg_cost<<<1, 1, 0, c->stream>>>(c->d_vals, c->d_w, c->lambda, c->n_active(), c->d_cost + 1);
    else
      k_reduce_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, nblocks, c->d_cost, 1);
    c->timing.kernel_launches += 3;
  }
  k_finish_cost<<<1, 1, 0, c->stream>>>(c->d_cost, tail);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  c->timing.num_evals += 1;
  return SRB_OK;
}

srb_status fetch_cost(srb_ctx* c, int slot, double* cost) {
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->h_cost, c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost,
                                    c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (cost) *cost = c->h_cost[slot];
  return SRB_OK;
}

void update_timing(srb_ctx* c) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->timing.last_eval_h2d_ms = ms;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->timing.last_eval_kernel_ms = ms;
  if (cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->timing.last_eval_d2h_ms = ms;
}

// Host part of srb_create, shared with srb_plan: validates the description the way the reference
// CHECK-fails (map_solver.cpp:52-76, blur_module.cpp:13-18, downsampling_module.cpp:13-17,
// objective_data_term.cpp:91-95), fills the geometry and quantises the warps.  No CUDA calls.
srb_status build_host_model(srb_ctx* c, const srb_model_desc* d, bool allow_empty_shard = false) {
  if (!d) return c->fail(SRB_ERR_INVALID, "null model description");
  if (d->lr_height <= 0 || d->lr_width <= 0 || d->num_channels <= 0)
    return c->fail(SRB_ERR_INVALID, "observation size and channel count must be positive");
  // map_solver.cpp:56-57 refuses an empty observation list; a frame shard of a multi-GPU run may be empty
  // (more devices than frames): it then contributes only its band of the regularization term
  if (d->num_frames < 0 || (d->num_frames == 0 && !allow_empty_shard))
    return c->fail(SRB_ERR_INVALID, "cannot solve with 0 observations");
  if (d->scale < 1) return c->fail(SRB_ERR_INVALID, "downsampling scale must be >= 1");
  if (d->psf_size < 0 || (d->psf_size > 0 && (d->psf_size % 2 == 0 || !d->psf)))
    return c->fail(SRB_ERR_INVALID, "blur kernel size must be odd and the kernel non-null");
  if (d->psf_size > 63) return c->fail(SRB_ERR_INVALID, "blur kernel larger than 63x63 is not supported");

  Geometry& G = c->g;
  G.h = d->lr_height; G.w = d->lr_width; G.s = d->scale; G.N = d->num_frames; G.Ct = d->num_channels;
  const long long Hl = (long long)G.h * G.s, Wl = (long long)G.w * G.s;
  // map_solver.cpp:96-101 guards C*H*W <= INT_MAX
  if (Hl * Wl * G.Ct > 2147483647LL) return c->fail(SRB_ERR_INVALID, "C*H*W exceeds INT_MAX");
  G.H = (int)Hl; G.W = (int)Wl;
  c->has_blur = d->psf_size > 0;
  G.K = c->has_blur ? d->psf_size : 1;
  G.hk = G.K / 2;
  c->P = (size_t)G.H * G.W;
  c->p = (size_t)G.h * G.w;
  c->psf_h.assign(1, 1.0);
  if (c->has_blur) c->psf_h.assign(d->psf, d->psf + (size_t)G.K * G.K);
  c->has_motion = d->shifts != nullptr;
  c->shifts_h.assign((size_t)2 * G.N, 0.0);
  if (c->has_motion) c->shifts_h.assign(d->shifts, d->shifts + (size_t)2 * G.N);
  for (double v : c->shifts_h)
    if (!(std::fabs(v) < 1.0e6)) return c->fail(SRB_ERR_INVALID, "motion shift out of range");

  // cv::resize must act block-regularly between LR and HR, which is what the fused residual /
  // sum-pool / zero-insert chain of the reference (objective_data_term.cpp:29,57-59) assumes.
  {
    int h2, w2;
    lr_size(G.s, G.H, G.W, &h2, &w2);
    bool ok = (h2 == G.h && w2 == G.w);
    for (int q = 0; ok && q < G.h; ++q) ok = nearest_index(q, G.H, G.h) == q * G.s;
    for (int q = 0; ok && q < G.w; ++q) ok = nearest_index(q, G.W, G.w) == q * G.s;
    for (int r = 0; ok && r < G.H; ++r) ok = nearest_index(r, G.h, G.H) == r / G.s;
    for (int r = 0; ok && r < G.W; ++r) ok = nearest_index(r, G.w, G.W) == r / G.s;
    if (!ok) return c->fail(SRB_ERR_GEOMETRY, "cv::resize nearest index map is not block-regular for this size/scale");
  }

  c->warp_fwd.resize(G.N);
  c->warp_tr.resize(G.N);
  c->warps_uniform = true;
  c->warps_integer = true;
  for (int k = 0; k < G.N; ++k) {
    const double dx = c->shifts_h[2 * k], dy = c->shifts_h[2 * k + 1];
    c->warp_fwd[k] = quantize_warp(dx, dy, G.H, nullptr);
    c->warp_tr[k] = quantize_warp(-dx, -dy, G.H, nullptr);
    c->warps_uniform = c->warps_uniform && c->warp_fwd[k].uniform && c->warp_tr[k].uniform;
    const int frac = (c->warp_fwd[k].nX | c->warp_fwd[k].nY | c->warp_tr[k].nX | c->warp_tr[k].nY) & 31;
    c->warps_integer = c->warps_integer && frac == 0;
  }
  return SRB_OK;
}

// w = 1 / max(1e-5, reg(x)) for a device-resident estimate (irls_map_solver.cpp:128-143), stream-ordered
srb_status reweight_dev(srb_ctx* c, const double* d_x) {
  k_reg_values<1><<<grid2d(c->g.W, c->g.H, c->Ca()), dim3(32, 8), 0, c->stream>>>(
      make_reg_params(c, c->Ca()), d_x, c->d_w);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

}  // namespace

#include "srb_cg_device.cuh"  // device-resident CG (needs eval_core)

// ================================================================================================
extern "C" {

const char* srb_version(void) { return "srb200 0.1 (sm_100a)"; }

int srb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

srb_status srb_plan(const srb_model_desc* d, srb_plan_info* out) {
  if (!out) return SRB_ERR_INVALID;
  memset(out, 0, sizeof *out);
  srb_ctx tmp;   // host-only use: no CUDA call is made on it
  srb_status st = build_host_model(&tmp, d);
  if (st != SRB_OK) {
    snprintf(out->why, sizeof out->why, "%s", tmp.err.c_str());
    return st;
  }
  TilePlan plan;
  plan_tile_model(tmp.g, tmp.psf_h, tmp.warp_fwd, tmp.warp_tr, tmp.warps_uniform, tmp.warps_integer, &plan);
  out->hr_height = tmp.g.H;
  out->hr_width = tmp.g.W;
  out->warps_uniform = tmp.warps_uniform ? 1 : 0;
  out->warps_integer = tmp.warps_integer ? 1 : 0;
  out->fused = plan.supported ? 1 : 0;
  snprintf(out->why, sizeof out->why, "%s", plan.why.c_str());
  if (!plan.supported) return SRB_OK;
  out->fractional = plan.frac ? 1 : 0;
  out->psf_half = plan.KH;
  out->num_entries = (int)plan.entries.size();
  out->min_entries_per_phase = out->num_entries;
  for (size_t ph = 0; ph + 1 < plan.phase_begin.size(); ++ph) {
    const int n = plan.phase_begin[ph + 1] - plan.phase_begin[ph];
    out->min_entries_per_phase = std::min(out->min_entries_per_phase, n);
    out->max_entries_per_phase = std::max(out->max_entries_per_phase, n);
  }
  out->band_lo_r = plan.band.lo_r; out->band_hi_r = plan.band.hi_r;
  out->band_lo_c = plan.band.lo_c; out->band_hi_c = plan.band.hi_c;
  out->table_driven = plan.fast[0].empty() ? 0 : plan.fast_E;
  out->zlayout = plan.zlayout;
  out->zt_frames = plan.zt ? plan.zt_n : 0;
  return SRB_OK;
}

int srb_quantize_shift(double d) { return quantize_warp(d, 0.0, 1, nullptr).nX; }

int srb_sample_is_special(int q, int hr_size, int psf_half, int scale, double shift) {
  const WarpQ f = quantize_warp(shift, 0.0, 1, nullptr), t = quantize_warp(-shift, 0.0, 1, nullptr);
  return sample_is_special(q, hr_size, psf_half, scale, f.nX, t.nX) ? 1 : 0;
}

const char* srb_last_error(const srb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static srb_status create_ctx(const srb_model_desc* d, int device, bool allow_empty_shard, srb_ctx** out);
srb_status srb_create(const srb_model_desc* d, int device, srb_ctx** out) { return create_ctx(d, device, false, out); }
srb_status srb_create_shard(const srb_model_desc* d, int device, srb_ctx** out) { return create_ctx(d, device, true, out); }

static srb_status create_ctx(const srb_model_desc* d, int device, bool allow_empty_shard, srb_ctx** out) {
  if (!out) return SRB_ERR_INVALID;
  *out = nullptr;
  srb_ctx* c = new (std::nothrow) srb_ctx();
  if (!c) return SRB_ERR_NOMEM;
  *out = c;  // returned even on failure so the caller can read srb_last_error, then srb_destroy
  {
    srb_status hst = build_host_model(c, d, allow_empty_shard);
    if (hst != SRB_OK) return hst;
  }
  Geometry& G = c->g;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return c->fail(SRB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return c->fail(SRB_ERR_INVALID, "invalid CUDA device index");
  c->device = device;
  SRB_CUDA_CHECK(c, cudaSetDevice(device));
  SRB_CUDA_CHECK(c, cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  SRB_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) SRB_CUDA_CHECK(c, cudaEventCreate(&e));
  SRB_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  SRB_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (auto& e : c->ev_in) SRB_CUDA_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : c->ev_k) SRB_CUDA_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : c->ev_pipe) SRB_CUDA_CHECK(c, cudaEventCreate(&e));
  if (const char* e = getenv("SRB_PIPE_CHUNKS")) c->pipe_chunks = std::max(1, std::min(atoi(e), (int)srb_ctx::kMaxPipe));

  // tables
  std::vector<int> src_r(G.h), src_c(G.w);
  for (int q = 0; q < G.h; ++q) src_r[q] = nearest_index(q, G.H, G.h);
  for (int q = 0; q < G.w; ++q) src_c[q] = nearest_index(q, G.W, G.w);
  std::vector<int> rowY_f((size_t)G.N * G.H), rowY_t((size_t)G.N * G.H), nX_f(G.N), nX_t(G.N);
  for (int k = 0; k < G.N; ++k) {
    const double dx = c->shifts_h[2 * k], dy = c->shifts_h[2 * k + 1];
    quantize_warp(dx, dy, G.H, &rowY_f[(size_t)k * G.H]);
    quantize_warp(-dx, -dy, G.H, &rowY_t[(size_t)k * G.H]);
    nX_f[k] = c->warp_fwd[k].nX;
    nX_t[k] = c->warp_tr[k].nX;
  }
  srb_status st;
#define ALLOC_COPY(dptr, hvec)                                                                   \
  if ((st = dev_alloc(c, &dptr, hvec.size())) != SRB_OK) return st;                              \
  SRB_CUDA_CHECK(c, cudaMemcpy(dptr, hvec.data(), hvec.size() * sizeof(hvec[0]), cudaMemcpyHostToDevice));
  ALLOC_COPY(c->d_psf, c->psf_h)
  ALLOC_COPY(c->d_src_r, src_r)
  ALLOC_COPY(c->d_src_c, src_c)
  ALLOC_COPY(c->d_rowY_fwd, rowY_f)
  ALLOC_COPY(c->d_rowY_tr, rowY_t)
  ALLOC_COPY(c->d_nX_fwd, nX_f)
  ALLOC_COPY(c->d_nX_tr, nX_t)
#undef ALLOC_COPY

  const size_t n_all = (size_t)G.Ct * c->P;
  if ((st = dev_alloc(c, &c->d_y, (size_t)G.N * G.Ct * c->p)) != SRB_OK) return st;
  if ((st = dev_alloc(c, &c->d_x, n_all)) != SRB_OK) return st;
  if ((st = dev_alloc(c, &c->d_grad, n_all + 1)) != SRB_OK) return st;
  if ((st = dev_alloc(c, &c->d_w, n_all)) != SRB_OK) return st;
  if ((st = dev_alloc(c, &c->d_cost, 4)) != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaMallocHost((void**)&c->h_cost, 4 * sizeof(double)));
  k_fill<<<fill_blocks(n_all), 256, 0, c->stream>>>(c->d_w, n_all, 1.0);
  c->c0 = 0;
  c->c1 = G.Ct;
  c->reg_row0 = 0;
  c->reg_row1 = G.H;
  if ((st = fused_setup(c)) != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

void srb_destroy(srb_ctx* c) {
  if (!c) return;
  if (c->stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  fused_teardown(c);
  dev_free(&c->d_psf); dev_free(&c->d_src_r); dev_free(&c->d_src_c);
  dev_free(&c->d_rowY_fwd); dev_free(&c->d_rowY_tr); dev_free(&c->d_nX_fwd); dev_free(&c->d_nX_tr);
  dev_free(&c->d_y); dev_free(&c->d_decay); dev_free(&c->d_w); dev_free(&c->d_x); dev_free(&c->d_grad);
  dev_free(&c->d_resid); dev_free(&c->d_pooled); dev_free(&c->d_vals); dev_free(&c->d_aux); dev_free(&c->d_partial);
  dev_free(&c->d_cost);
  dev_free(&c->cg_store);
  if (c->cg_h_out) cudaFreeHost(c->cg_h_out);
  dev_free(&c->peer.d_err);
  for (auto& st : c->peer.s_copy)
    if (st) cudaStreamDestroy(st);
  for (auto& e : c->peer.ev_band)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->peer.ev_copy)
    if (e) cudaEventDestroy(e);
  if (c->h_cost) cudaFreeHost(c->h_cost);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_in)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_k)
    if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_pipe)
    if (e) cudaEventDestroy(e);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

srb_status srb_set_observations(srb_ctx* c, const double* lr_host) {
  if (!c) return SRB_ERR_INVALID;
  if (!lr_host) return c->fail(SRB_ERR_INVALID, "null observations");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (c->g.N > 0)
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_y, lr_host, (size_t)c->g.N * c->g.Ct * c->p * sizeof(double),
                                      cudaMemcpyHostToDevice, c->stream));
  srb_status zst = fused_observations_changed(c);
  if (zst != SRB_OK) return zst;
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  c->have_obs = true;
  return SRB_OK;
}

srb_status srb_set_observations_dev(srb_ctx* c, const double* lr_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!lr_dev) return c->fail(SRB_ERR_INVALID, "null observations");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_y, lr_dev, (size_t)c->g.N * c->g.Ct * c->p * sizeof(double),
                                    cudaMemcpyDeviceToDevice, c->stream));
  srb_status zst = fused_observations_changed(c);
  if (zst != SRB_OK) return zst;
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  c->have_obs = true;
  return SRB_OK;
}

static srb_status reset_weights(srb_ctx* c) {
  const size_t n = (size_t)c->g.Ct * c->P;
  k_fill<<<fill_blocks(n), 256, 0, c->stream>>>(c->d_w, n, 1.0);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

srb_status srb_set_channel_range(srb_ctx* c, int c0, int c1) {
  if (!c) return SRB_ERR_INVALID;
  // objective_data_term.cpp:91-95
  if (c0 < 0 || c1 > c->g.Ct || c1 <= c0) return c->fail(SRB_ERR_INVALID, "invalid channel range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  c->c0 = c0;
  c->c1 = c1;
  c->x_resident = false;
  return reset_weights(c);
}

srb_status srb_set_regularizer(srb_ctx* c, int kind, double lambda, int btv_range, double btv_decay) {
  if (!c) return SRB_ERR_INVALID;
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (kind == SRB_REG_NONE || !(lambda > 0.0)) {
    c->reg_kind = SRB_REG_NONE;
    c->lambda = 0.0;
    return reset_weights(c);
  }
  if (kind != SRB_REG_TV && kind != SRB_REG_TV3D && kind != SRB_REG_BTV)
    return c->fail(SRB_ERR_INVALID, "unknown regularizer kind");
  if (kind == SRB_REG_BTV) {
    // btv_regularizer.cpp:58-61
    if (btv_range < 1) return c->fail(SRB_ERR_INVALID, "BTV scale range must be at least 1");
    if (!(btv_decay > 0.0 && btv_decay <= 1.0))
      return c->fail(SRB_ERR_INVALID, "BTV spatial decay must be in (0, 1]");
    if (btv_range > 16) return c->fail(SRB_ERR_INVALID, "BTV scale range larger than 16 is not supported");
    std::vector<double> tab(2 * btv_range + 1);
    for (int t = 0; t <= 2 * btv_range; ++t) tab[t] = std::pow(btv_decay, t);  // btv_regularizer.cpp:39
    dev_free(&c->d_decay);
    srb_status st = dev_alloc(c, &c->d_decay, tab.size());
    if (st != SRB_OK) return st;
    SRB_CUDA_CHECK(c, cudaMemcpy(c->d_decay, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->btv_R = btv_range;
    c->btv_decay = btv_decay;
    c->decay_h = tab;
  }
  c->reg_kind = kind;
  c->lambda = lambda;
  return reset_weights(c);
}

srb_status srb_set_irls_weights(srb_ctx* c, const double* w) {
  if (!c) return SRB_ERR_INVALID;
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (!w) return reset_weights(c);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_w, w, c->n_active() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_reweight(srb_ctx* c, const double* x_host, double* w_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!reg_active(c)) return c->fail(SRB_ERR_STATE, "no regularizer configured");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (x_host) {
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, c->n_active() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->x_resident = true;
  } else if (!c->x_resident) {
    // only the host-buffer entry points (srb_eval, srb_cg_minimize, ...) leave the estimate in the context;
    // the *_dev forms work on the caller's device buffer, which this context does not keep
    return c->fail(SRB_ERR_STATE, "srb_reweight(x = NULL) needs a preceding evaluation or solve with a HOST estimate "
                                  "on the current channel range; use srb_reweight_dev for device-resident estimates");
  }
  srb_status rst = reweight_dev(c, c->d_x);
  if (rst != SRB_OK) return rst;
  if (w_out)
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(w_out, c->d_w, c->n_active() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_reweight_dev(srb_ctx* c, const double* x_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!reg_active(c)) return c->fail(SRB_ERR_STATE, "no regularizer configured");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  return reweight_dev(c, x_dev);
}

// ---- device-resident solver --------------------------------------------------------
static CgOptions cg_options_from(const srb_cg_options* o) {
  CgOptions opt;
  if (o) {
    opt.epsg = o->gradient_norm_threshold;
    opt.epsf = o->cost_decrease_threshold;
    opt.epsx = o->parameter_variation_threshold;
    opt.maxits = o->max_num_solver_iterations;
  }
  return opt;
}
static bool cg_options_valid(const srb_cg_options* o) {
  // mincgsetcond asserts finite, non-negative thresholds and a non-negative iteration limit
  return !o || (o->num_lbfgs_hessian_corrections >= 0 && o->num_lbfgs_hessian_corrections <= 64 &&
                std::isfinite(o->gradient_norm_threshold) && o->gradient_norm_threshold >= 0 &&
                std::isfinite(o->cost_decrease_threshold) && o->cost_decrease_threshold >= 0 &&
                std::isfinite(o->parameter_variation_threshold) && o->parameter_variation_threshold >= 0 &&
                o->max_num_solver_iterations >= 0);
}
static void cg_report_to(const CgReport& r, srb_cg_report* out) {
  if (!out) return;
  out->iterations = r.iterations;
  out->num_evaluations = r.nfev;
  out->termination_type = r.termination;
  out->num_restarts = r.restarts;
  out->final_cost = r.f;
}

srb_status srb_cg_minimize_dev(srb_ctx* c, double* x_dev, const srb_cg_options* options, srb_cg_report* report) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!cg_options_valid(options)) return c->fail(SRB_ERR_INVALID, "invalid solver thresholds");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  CgReport rep;
  srb_status st = cg_minimize_dev(c, x_dev, cg_options_from(options), &rep);
  if (st != SRB_OK) return st;
  cg_report_to(rep, report);
  return SRB_OK;
}

srb_status srb_lbfgs_minimize_dev(srb_ctx* c, double* x_dev, const srb_cg_options* options, srb_cg_report* report) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!options || !cg_options_valid(options) || options->num_lbfgs_hessian_corrections < 1)
    return c->fail(SRB_ERR_INVALID, "invalid solver options (L-BFGS needs 1..64 correction pairs)");
  if ((long long)options->num_lbfgs_hessian_corrections > (long long)c->n_active())
    return c->fail(SRB_ERR_INVALID, "more correction pairs than parameters");  // minlbfgscreate asserts M <= N
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  CgReport rep;
  srb_status st = lbfgs_minimize_dev(c, x_dev, options->num_lbfgs_hessian_corrections, cg_options_from(options), &rep);
  if (st != SRB_OK) return st;
  cg_report_to(rep, report);
  return SRB_OK;
}

srb_status srb_lbfgs_minimize(srb_ctx* c, double* x_host, const srb_cg_options* options, srb_cg_report* report) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host) return c->fail(SRB_ERR_INVALID, "null estimate");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t bytes = c->n_active() * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = true;
  srb_status st = srb_lbfgs_minimize_dev(c, c->d_x, options, report);
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(x_host, c->d_x, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_cg_minimize(srb_ctx* c, double* x_host, const srb_cg_options* options, srb_cg_report* report) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host) return c->fail(SRB_ERR_INVALID, "null estimate");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t bytes = c->n_active() * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = true;
  srb_status st = srb_cg_minimize_dev(c, c->d_x, options, report);
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(x_host, c->d_x, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_solve_irls(srb_ctx* c, double* x_host, const srb_cg_options* options, int max_num_irls_iterations,
                          double irls_cost_difference_threshold, srb_irls_report* report) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!cg_options_valid(options) || max_num_irls_iterations < 0 || !(irls_cost_difference_threshold >= 0))
    return c->fail(SRB_ERR_INVALID, "invalid solver options");
  // the outer loop runs while |cost difference| >= threshold (irls_map_solver.cpp:76-78,145-152): with a zero
  // threshold and no iteration limit it cannot end once a regularizer is configured
  if (max_num_irls_iterations == 0 && irls_cost_difference_threshold == 0.0 && reg_active(c))
    return c->fail(SRB_ERR_INVALID, "unlimited IRLS iterations with a zero cost-difference threshold never terminate "
                                    "(the reference's defaults are 20 and 1e-5, irls_map_solver.h:27,35)");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t bytes = c->n_active() * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = true;
  srb_status st = reset_weights(c);  // irls_map_solver.cpp:66-74: all weights 1
  if (st != SRB_OK) return st;
  IrlsReport rep;
  st = irls_solve_dev(c, c->d_x, cg_options_from(options), max_num_irls_iterations,
                      irls_cost_difference_threshold, reg_active(c),
                      options ? options->num_lbfgs_hessian_corrections : 0, &rep);
  if (st != SRB_OK) return st;
  srb_irls_report out{};
  out.num_irls_iterations = rep.irls_iterations;
  out.num_solver_iterations = rep.solver_iterations;
  out.num_evaluations = rep.nfev;
  out.last_termination_type = rep.last_termination;
  out.final_cost = rep.f;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(x_host, c->d_x, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (report) *report = out;
  return SRB_OK;
}

srb_status srb_set_path(srb_ctx* c, int path) {
  if (!c) return SRB_ERR_INVALID;
  if (path != SRB_PATH_AUTO && path != SRB_PATH_REFERENCE_ORDER && path != SRB_PATH_FUSED)
    return c->fail(SRB_ERR_INVALID, "unknown path");
  if (path == SRB_PATH_FUSED && !fused_supported(c))
    return c->fail(SRB_ERR_INVALID, std::string("the fused kernel does not cover this model: ") +
                                        (tile_state(c) ? tile_state(c)->why : std::string("not initialised")));
  c->path = path;
  return SRB_OK;
}
int srb_active_path(const srb_ctx* c) { return c ? resolve_path(c) : -1; }
srb_status srb_set_strict_cost(srb_ctx* c, int on) {
  if (!c) return SRB_ERR_INVALID;
  c->strict_cost = on != 0;
  return SRB_OK;
}
int srb_zlayout_active(const srb_ctx* c) {
  const TileState* st = c ? tile_state(c) : nullptr;
  return (st && st->supported && (st->d_yz || st->d_yzt) && st->yz_valid) ? 1 : 0;
}

srb_status srb_set_regularizer_rows(srb_ctx* c, int row_begin, int row_end) {
  if (!c) return SRB_ERR_INVALID;
  if (row_begin < 0 || row_end > c->g.H || row_end < row_begin)
    return c->fail(SRB_ERR_INVALID, "invalid regularizer row band");
  c->reg_row0 = row_begin;
  c->reg_row1 = row_end;
  return SRB_OK;
}

// ---- hot path -----------------------------------------------------------------------------------
static bool units_pipelined(const srb_ctx* c);
static bool host_slices_ok(const srb_ctx* c);
static bool unit_ranges_ok(const srb_ctx* c);
// Slices of the host <-> device pipeline of srb_eval: at most pipe_chunks (SRB_PIPE_CHUNKS, default 16), at most one
// per unit, and no slice below 4 MB -- a small problem (cfg2: 2 MB) is faster as one copy, one launch, one copy
// than as a train of tiny launches.
static int host_pipe_chunks(const srb_ctx* c) {
  const int nu = tile_rows_per_channel(c) * c->Ca();
  const long long by_size = (long long)(c->n_active() * sizeof(double)) / (4ll << 20);
  return (int)std::max(1ll, std::min<long long>(std::min(c->pipe_chunks, nu), by_size));
}

// srb_eval, pipelined: the estimate goes to the device in contiguous slices on a copy-in stream, the
// tile kernel evaluates the units whose rows (and the halo rows of the next slice) have arrived, and
// every finished gradient slice returns on a copy-out stream -- H2D, compute and D2H overlap, so the
// call costs about one PCIe direction instead of two (both directions of the link work at once).
static srb_status eval_host_pipelined(srb_ctx* c, const double* x_host, double* g_host, double* cost) {
  c->x_resident = true;
  const int nu = tile_rows_per_channel(c) * c->Ca();
  const TileLayout L0 = tile_layout(c);
  if (L0.nband) {  // the border band is evaluated slice by slice too: its cost slots accumulate over the slices
    srb_status pst = ensure_partials(c, 2 * L0.nblocks + L0.nband);
    if (pst != SRB_OK) return pst;
    SRB_CUDA_CHECK(c, cudaMemsetAsync(c->d_partial + L0.nblocks, 0, L0.nband * sizeof(double), c->stream));
  }
  const int nch = host_pipe_chunks(c);
  unsigned long long b[srb_ctx::kMaxPipe + 1];
  int u[srb_ctx::kMaxPipe + 1];
  {
    const int tr = tile_rows_per_channel(c), TH = tile_height(c);
    for (int i = 0; i <= nch; ++i) {
      u[i] = (int)((long long)i * nu / nch);
      const int ch = u[i] / tr, t = u[i] - ch * tr;   // first element of unit u[i] (memory order)
      const int row = t * TH < c->g.H ? t * TH : c->g.H;
      b[i] = (unsigned long long)ch * c->P + (unsigned long long)row * c->g.W;
    }
  }
  b[nch] = c->n_active();
  const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
  cudaEventRecord(c->ev_pipe[0], c->s_in);
  for (int i = 0; i < nch; ++i) {
    SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x + b[i], x_host + b[i], (b[i + 1] - b[i]) * sizeof(double),
                                      cudaMemcpyHostToDevice, c->s_in));
    cudaEventRecord(c->ev_in[i], c->s_in);
  }
  cudaEventRecord(c->ev_pipe[1], c->s_in);
  cudaEventRecord(c->ev[1], c->stream);
  for (int i = 0; i < nch; ++i) {
    // rows of slice i need the halo rows below them: the head of slice i + 1
    SRB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_in[std::min(i + 1, nch - 1)], 0));
    bool reg_done = false;
    srb_status st = fused_eval_units(c, c->d_x, g_host ? c->d_grad : nullptr, do_reg, u[i], u[i + 1], &reg_done);
    if (st != SRB_OK) return st;
    if ((st = fused_band_units(c, c->d_x, g_host ? c->d_grad : nullptr, u[i], u[i + 1])) != SRB_OK) return st;
    if (g_host) {
      cudaEventRecord(c->ev_k[i], c->stream);
      SRB_CUDA_CHECK(c, cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0));
      if (i == 0) cudaEventRecord(c->ev_pipe[2], c->s_out);
      SRB_CUDA_CHECK(c, cudaMemcpyAsync(g_host + b[i], c->d_grad + b[i], (b[i + 1] - b[i]) * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->s_out));
    }
  }
  srb_status st = fused_eval_finish(c, c->d_x, g_host ? c->d_grad : nullptr, nullptr, /*run_band=*/false);
  if (st != SRB_OK) return st;
  cudaEventRecord(c->ev[2], c->stream);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->h_cost, c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (g_host) cudaEventRecord(c->ev_pipe[3], c->s_out);
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (g_host) SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->s_out));
  c->timing.num_evals += 1;
  if (cost) *cost = c->h_cost[2];
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev_pipe[0], c->ev_pipe[1]) == cudaSuccess) c->timing.last_eval_h2d_ms = ms;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->timing.last_eval_kernel_ms = ms;
  if (g_host && cudaEventElapsedTime(&ms, c->ev_pipe[2], c->ev_pipe[3]) == cudaSuccess) c->timing.last_eval_d2h_ms = ms;
  (void)cudaGetLastError();
  return SRB_OK;
}

srb_status srb_eval(srb_ctx* c, const double* x_host, double* g_host, double* cost) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (unit_ranges_ok(c) && host_slices_ok(c) && host_pipe_chunks(c) > 1) return eval_host_pipelined(c, x_host, g_host, cost);
  const size_t bytes = c->n_active() * sizeof(double);
  cudaEventRecord(c->ev[0], c->stream);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = true;
  cudaEventRecord(c->ev[1], c->stream);
  srb_status st = eval_core(c, c->d_x, g_host ? c->d_grad : nullptr, nullptr, true, true, false);
  if (st != SRB_OK) return st;
  cudaEventRecord(c->ev[2], c->stream);
  if (g_host) SRB_CUDA_CHECK(c, cudaMemcpyAsync(g_host, c->d_grad, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->h_cost, c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  cudaEventRecord(c->ev[3], c->stream);
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (cost) *cost = c->h_cost[2];
  update_timing(c);
  return SRB_OK;
}

srb_status srb_eval_dev(srb_ctx* c, const double* x_dev, double* g_dev, double* cost) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null estimate");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  cudaEventRecord(c->ev[0], c->stream);
  cudaEventRecord(c->ev[1], c->stream);
  srb_status st = eval_core(c, x_dev, g_dev, nullptr, true, true, false);
  if (st != SRB_OK) return st;
  cudaEventRecord(c->ev[2], c->stream);
  cudaEventRecord(c->ev[3], c->stream);
  if (cost) {
    st = fetch_cost(c, 2, cost);
    if (st != SRB_OK) return st;
    update_timing(c);
  }
  return SRB_OK;
}

srb_status srb_eval_partial_dev(srb_ctx* c, const double* x_dev, double* gc_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev || !gc_dev) return c->fail(SRB_ERR_INVALID, "null buffer");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  return eval_core(c, x_dev, gc_dev, gc_dev + c->n_active(), true, true, false);
}

// ---- pipelined multi-GPU form -------------------------------------------------------------------
static bool units_pipelined(const srb_ctx* c) {
  // units need the tile kernel to be the only writer of the gradient: no border band, and a
  // regularizer that is fused (2-D TV) or absent
  const TileState* st = tile_state(c);
  const bool reg_ok = !reg_active(c) || fused_reg_covered(c);
  return resolve_path(c) == SRB_PATH_FUSED && st && !st->has_band && reg_ok;
}
// srb_eval / srb_multi_eval stream x to the device slice by slice in memory order; 3-D TV reads the next
// channel's plane at the same rows, which has not arrived when a slice is evaluated
static bool host_slices_ok(const srb_ctx* c) { return !(reg_active(c) && c->reg_kind == SRB_REG_TV3D); }

srb_status srb_num_units(srb_ctx* c, int* num_units, int* rows_per_unit) {
  if (!c || !num_units) return SRB_ERR_INVALID;
  if (units_pipelined(c)) {
    *num_units = tile_rows_per_channel(c) * c->Ca();
    if (rows_per_unit) *rows_per_unit = tile_height(c);
  } else {
    *num_units = 1;
    if (rows_per_unit) *rows_per_unit = c->g.H * c->Ca();
  }
  return SRB_OK;
}

srb_status srb_unit_range(srb_ctx* c, int u0, int u1, unsigned long long* begin, unsigned long long* end) {
  if (!c || !begin || !end) return SRB_ERR_INVALID;
  int nu = 0;
  srb_num_units(c, &nu, nullptr);
  if (u0 < 0 || u1 > nu || u1 < u0) return c->fail(SRB_ERR_INVALID, "invalid unit range");
  if (!units_pipelined(c)) {
    *begin = 0;
    *end = u1 > u0 ? c->n_active() : 0;
    return SRB_OK;
  }
  const int tr = tile_rows_per_channel(c), TH = tile_height(c);
  auto first_elem = [&](int u) -> unsigned long long {
    const int ch = u / tr, t = u - ch * tr;
    const int row = t * TH < c->g.H ? t * TH : c->g.H;
    return (unsigned long long)ch * c->P + (unsigned long long)row * c->g.W;
  };
  *begin = first_elem(u0);
  *end = first_elem(u1);
  return SRB_OK;
}

srb_status srb_eval_units_dev(srb_ctx* c, const double* x_dev, double* gc_dev, int u0, int u1) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev || !gc_dev) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  int nu = 0;
  srb_num_units(c, &nu, nullptr);
  if (u0 < 0 || u1 > nu || u1 <= u0) return c->fail(SRB_ERR_INVALID, "invalid unit range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (!units_pipelined(c))  // single unit: the whole evaluation happens in srb_eval_finish_dev
    return SRB_OK;
  const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
  bool reg_done = false;
  srb_status st = fused_eval_units(c, x_dev, gc_dev, do_reg, u0, u1, &reg_done);
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

srb_status srb_eval_finish_dev(srb_ctx* c, const double* x_dev, double* gc_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev || !gc_dev) return c->fail(SRB_ERR_INVALID, "null buffer");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  if (!units_pipelined(c)) return eval_core(c, x_dev, gc_dev, gc_dev + c->n_active(), true, true, false);
  srb_status st = fused_eval_finish(c, x_dev, gc_dev, gc_dev + c->n_active());
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  c->timing.num_evals += 1;
  return SRB_OK;
}

// Row-band partition, one process per GPU: units [u0, u1) of the WHOLE objective (this context holds every frame)
// and the cost of exactly those units -- a rank's share; the gradient rows written are final.
// units of the whole objective can be evaluated range by range: fused path with a regularizer it covers (a border
// band is evaluated by rows too)
static bool unit_ranges_ok(const srb_ctx* c) {
  const bool reg_ok = !reg_active(c) || fused_reg_covered(c);
  return resolve_path(c) == SRB_PATH_FUSED && tile_state(c) && reg_ok;
}

srb_status srb_eval_unit_range_dev(srb_ctx* c, const double* x_dev, double* g_dev, int u0, int u1, double* cost_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  if (!unit_ranges_ok(c)) return c->fail(SRB_ERR_STATE, "unit ranges need the fused tile kernel and a regularizer it covers");
  const int nu = tile_rows_per_channel(c) * c->Ca();
  if (u0 < 0 || u1 > nu || u1 < u0) return c->fail(SRB_ERR_INVALID, "invalid unit range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const TileLayout L = tile_layout(c);
  {
    srb_status st = ensure_partials(c, 2 * L.nblocks + L.nband);
    if (st != SRB_OK) return st;
  }
  if (L.nband) SRB_CUDA_CHECK(c, cudaMemsetAsync(c->d_partial + L.nblocks, 0, L.nband * sizeof(double), c->stream));
  if (u1 > u0) {
    const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
    bool reg_done = false;
    srb_status st = fused_eval_units(c, x_dev, g_dev, do_reg, u0, u1, &reg_done);
    if (st != SRB_OK) return st;
    if ((st = fused_band_units(c, x_dev, g_dev, u0, u1)) != SRB_OK) return st;
  }
  // the cost slots of units [u0, u1) are contiguous: slot = unit * tiles_per_row + tile column
  const size_t per_unit = L.nblocks / (size_t)nu;
  const size_t first = (size_t)u0 * per_unit, count = (size_t)(u1 - u0) * per_unit;
  k_finish_partials3<<<1, 1024, 0, c->stream>>>(c->d_partial + first, count, c->d_partial + L.nblocks, L.nband,
                                                c->d_partial + L.nblocks + L.nband + first, count, c->d_cost, cost_dev);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  c->timing.num_evals += 1;
  return SRB_OK;
}

int srb_halo_rows(const srb_ctx* c) { return (c && tile_state(c)) ? stencil_halo_rows(c) : 0; }

// ---- multi-GPU peer path ------------------------------------------------------------------------
srb_status srb_dev_alloc(void** ptr, unsigned long long bytes) {
  if (!ptr || !bytes) return SRB_ERR_INVALID;
  if (cudaMalloc(ptr, (size_t)bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    *ptr = nullptr;
    return SRB_ERR_NOMEM;
  }
  cudaMemset(*ptr, 0, (size_t)bytes);
  return SRB_OK;
}
srb_status srb_dev_free(void* ptr) {
  if (!ptr) return SRB_ERR_INVALID;
  return cudaFree(ptr) == cudaSuccess ? SRB_OK : SRB_ERR_CUDA;
}
srb_status srb_ipc_export(const void* dev_ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!dev_ptr || !handle) return SRB_ERR_INVALID;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)) != cudaSuccess) {
    (void)cudaGetLastError();
    return SRB_ERR_CUDA;
  }
  memcpy(handle, &h, 64);
  return SRB_OK;
}
srb_status srb_ipc_open(const unsigned char handle[64], void** dev_ptr) {
  if (!handle || !dev_ptr) return SRB_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  if (cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void)cudaGetLastError();
    *dev_ptr = nullptr;
    return SRB_ERR_CUDA;
  }
  return SRB_OK;
}
srb_status srb_ipc_close(void* dev_ptr) {
  if (!dev_ptr) return SRB_ERR_INVALID;
  return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? SRB_OK : SRB_ERR_CUDA;
}

// Contiguous ownership: rank o owns units [o*nu/world, (o+1)*nu/world), a contiguous element range.
static void peer_bands(const srb_ctx* c, int world, int* band_unit, long long* band_elem, long long* cap) {
  const int nu = tile_rows_per_channel(c) * c->Ca();
  const int tr = tile_rows_per_channel(c), TH = tile_height(c);
  *cap = 0;
  for (int o = 0; o <= world; ++o) {
    const int u = (int)((long long)o * nu / world);
    band_unit[o] = u;
    const int ch = u / tr, t = u - ch * tr;
    const int row = t * TH < c->g.H ? t * TH : c->g.H;
    band_elem[o] = u >= nu ? (long long)c->n_active() : (long long)ch * (long long)c->P + (long long)row * c->g.W;
    if (o > 0 && band_elem[o] - band_elem[o - 1] > *cap) *cap = band_elem[o] - band_elem[o - 1];
  }
}

srb_status srb_peer_sizes(srb_ctx* c, int world, unsigned long long* slots_bytes, unsigned long long* out_bytes) {
  if (!c || !slots_bytes || !out_bytes) return SRB_ERR_INVALID;
  if (world < 1 || world > SRB_MAX_PEERS) return c->fail(SRB_ERR_INVALID, "world size must be 1..8");
  if (!units_pipelined(c)) return c->fail(SRB_ERR_STATE, "the peer path needs the fused tile kernel without a border band");
  int bu[SRB_MAX_PEERS + 1];
  long long be[SRB_MAX_PEERS + 1], cap;
  peer_bands(c, world, bu, be, &cap);
  *slots_bytes = (unsigned long long)world * (unsigned long long)cap * sizeof(double);
  // gradient, total cost, `world` partial costs, 2 x `world` barrier flags
  *out_bytes = (unsigned long long)(c->n_active() + 1 + 3 * world) * sizeof(double);
  return SRB_OK;
}

srb_status srb_peer_setup(srb_ctx* c, int rank, int world, double* const* slot_bases, double* const* out_bases) {
  if (!c || !slot_bases || !out_bases) return SRB_ERR_INVALID;
  if (world < 1 || world > SRB_MAX_PEERS || rank < 0 || rank >= world) return c->fail(SRB_ERR_INVALID, "bad rank / world");
  if (!units_pipelined(c)) return c->fail(SRB_ERR_STATE, "the peer path needs the fused tile kernel without a border band");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  srb_ctx::Peer& p = c->peer;
  p.rank = rank;
  p.world = world;
  peer_bands(c, world, p.band_unit, p.band_elem, &p.band_cap);
  for (int o = 0; o < world; ++o) {
    if (!slot_bases[o] || !out_bases[o]) return c->fail(SRB_ERR_INVALID, "null peer buffer");
    p.slots[o] = slot_bases[o];
    p.out[o] = out_bases[o];
  }
  p.epoch = 0;
  if (!p.d_err) {
    SRB_CUDA_CHECK(c, cudaMalloc((void**)&p.d_err, 2 * sizeof(int)));
    SRB_CUDA_CHECK(c, cudaMemset(p.d_err, 0, 2 * sizeof(int)));
    {
      // highest priority: whatever executes the peer copies (copy engine or a copy kernel) must get
      // going while the tile kernel of the next band still fills the SMs
      int lo = 0, hi = 0;
      SRB_CUDA_CHECK(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      for (auto& s : p.s_copy) SRB_CUDA_CHECK(c, cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    }
    for (auto& e : p.ev_band) SRB_CUDA_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : p.ev_copy) SRB_CUDA_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  p.active = true;
  return SRB_OK;
}

// Phase 1.  The tile kernel evaluates the gradient band by band, the bands of the OTHER ranks first
// (starting with the next rank, so that at any time every rank's link carries one band in each
// direction); as soon as a band is finished a copy engine pushes it over NVLink into slot [rank] of
// its owner, while the SMs are already computing the next band.  The own band is evaluated last and
// stays local.  After the last push the partial cost and the phase-0 flag go to every rank.
srb_status srb_peer_scatter_dev(srb_ctx* c, const double* x_dev) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_dev) return c->fail(SRB_ERR_INVALID, "null estimate");
  if (!c->peer.active) return c->fail(SRB_ERR_STATE, "srb_peer_setup has not been called");
  if (!c->have_obs) return c->fail(SRB_ERR_STATE, "srb_set_observations has not been called");
  if (!units_pipelined(c)) return c->fail(SRB_ERR_STATE, "configuration changed: the peer path no longer applies");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  srb_ctx::Peer& p = c->peer;
  const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
  static const bool trace = getenv("SRB_PEER_TRACE") != nullptr;  // one-off timeline on stderr
  cudaEvent_t tr[20] = {};
  if (trace) {
    for (auto& e : tr) cudaEventCreate(&e);
    cudaEventRecord(tr[0], c->stream);
  }
  // Bands are evaluated in groups of `grp` consecutive owners per tile-kernel launch and pushed as soon
  // as the group's launch has finished.  One band per launch measured best (8 GPUs, cfg3: 0.423 ms per
  // step vs 0.445 with two bands per launch, although a band is then only 1.3 waves of CTAs);
  // SRB_PEER_GROUP overrides.
  static const int grp_env = getenv("SRB_PEER_GROUP") ? atoi(getenv("SRB_PEER_GROUP")) : 0;
  const int grp = grp_env > 0 ? grp_env : 1;
  for (int i0 = 0; i0 < p.world; i0 += grp) {
    const int i1 = std::min(i0 + grp, p.world);
    // owners (rank + 1 + i) % world for i in [i0, i1): one or two contiguous unit ranges
    for (int i = i0; i < i1;) {
      const int o = (p.rank + 1 + i) % p.world;  // i == world - 1  <=>  o == rank
      int j = i;
      while (j + 1 < i1 && (p.rank + 1 + j + 1) % p.world == (p.rank + 1 + j) % p.world + 1) ++j;
      const int o_last = (p.rank + 1 + j) % p.world;
      if (p.band_unit[o_last + 1] > p.band_unit[o]) {
        bool reg_done = false;
        srb_status st = fused_eval_units(c, x_dev, c->d_grad, do_reg, p.band_unit[o], p.band_unit[o_last + 1], &reg_done);
        if (st != SRB_OK) return st;
      }
      i = j + 1;
    }
    SRB_CUDA_CHECK(c, cudaEventRecord(p.ev_band[i0 / grp], c->stream));
    for (int i = i0; i < i1; ++i) {
      const int o = (p.rank + 1 + i) % p.world;
      if (o == p.rank || p.band_elem[o + 1] <= p.band_elem[o]) continue;
      cudaStream_t sc = p.s_copy[i & 1];
      SRB_CUDA_CHECK(c, cudaStreamWaitEvent(sc, p.ev_band[i0 / grp], 0));
      SRB_CUDA_CHECK(c, cudaMemcpyAsync(p.slots[o] + (long long)p.rank * p.band_cap, c->d_grad + p.band_elem[o],
                                        (size_t)(p.band_elem[o + 1] - p.band_elem[o]) * sizeof(double),
                                        cudaMemcpyDeviceToDevice, sc));
      if (trace) cudaEventRecord(tr[10 + i], sc);
    }
    if (trace) cudaEventRecord(tr[1 + i0 / grp], c->stream);
  }
  for (int k = 0; k < 2; ++k) {  // the flag may only be raised once every push has been delivered
    SRB_CUDA_CHECK(c, cudaEventRecord(p.ev_copy[k], p.s_copy[k]));
    SRB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, p.ev_copy[k], 0));
  }
  GatherParams G;
  G.world = p.world;
  G.rank = p.rank;
  for (int r = 0; r < p.world; ++r) G.out[r] = p.out[r];
  const TileLayout L = tile_layout(c);
  p.epoch += 1;
  k_peer_finish_scatter<<<1, 1024, 0, c->stream>>>(c->d_partial, L.nblocks + L.nband, c->d_partial + L.nblocks + L.nband,
                                                   L.nblocks, c->d_cost, G, (long long)c->n_active() + 1,
                                                   (long long)c->n_active() + 1 + p.world, p.epoch);
  if (trace) {
    cudaEventRecord(tr[9], c->stream);
    cudaStreamSynchronize(c->stream);
    float ms = 0.f;
    fprintf(stderr, "[srb peer trace rank %d]", p.rank);
    for (int gi = 0; gi * grp < p.world; ++gi) {
      cudaEventElapsedTime(&ms, tr[0], tr[1 + gi]);
      fprintf(stderr, " group%d kernels done %.3f", gi, ms);
    }
    for (int i = 0; i + 1 < p.world; ++i)
      if (cudaEventElapsedTime(&ms, tr[0], tr[10 + i]) == cudaSuccess) fprintf(stderr, " push%d done %.3f", i, ms);
    cudaEventElapsedTime(&ms, tr[0], tr[9]);
    fprintf(stderr, " | flags raised %.3f ms\n", ms);
    for (auto& e : tr) cudaEventDestroy(e);
    (void)cudaGetLastError();  // unrecorded trace events
  }
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  c->timing.num_evals += 1;
  return SRB_OK;
}

// Phase 2: wait for every rank's contribution, sum this rank's band in fixed rank order and store it,
// with the total cost, into every rank's gradient buffer; then wait until all bands have arrived here.
srb_status srb_peer_gather_dev(srb_ctx* c) {
  if (!c) return SRB_ERR_INVALID;
  if (!c->peer.active) return c->fail(SRB_ERR_STATE, "srb_peer_setup has not been called");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const srb_ctx::Peer& p = c->peer;
  GatherParams G;
  G.world = p.world;
  G.rank = p.rank;
  G.band_begin = p.band_elem[p.rank];
  G.band_len = p.band_elem[p.rank + 1] - p.band_elem[p.rank];
  G.band_cap = p.band_cap;
  G.slots = p.slots[p.rank];
  G.own = c->d_grad + p.band_elem[p.rank];
  for (int r = 0; r < p.world; ++r) G.out[r] = p.out[r];
  const long long flag_base = (long long)c->n_active() + 1 + p.world;
  k_sum_gather<<<c->num_sms * 8, 256, 0, c->stream>>>(G, (long long)c->n_active(), flag_base, p.epoch,
                                                      reinterpret_cast<unsigned int*>(p.d_err + 1), p.d_err);
  // every rank's band (and with it the full gradient) has landed in this rank's buffer
  k_peer_wait<<<1, 32, 0, c->stream>>>(p.out[p.rank], flag_base, 1, p.world, p.epoch, p.d_err);
  c->timing.kernel_launches += 2;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  return SRB_OK;
}

// The flag barriers of the peer path spin a bounded number of times; a rank that never arrives makes them
// record the failure on the device instead of hanging the GPU (and nothing partial is published).  Every
// synchronising entry point reads that record back here: the evaluation is then reported as failed and the
// record (and the last-block counter of k_sum_gather) is cleared so that the next evaluation starts clean.
static srb_status peer_status(srb_ctx* c) {
  srb_ctx::Peer& p = c->peer;
  if (!p.active || !p.d_err) return SRB_OK;
  int e[2] = {0, 0};
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(e, p.d_err, sizeof e, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  if (e[0] == 0) return SRB_OK;
  SRB_CUDA_CHECK(c, cudaMemsetAsync(p.d_err, 0, sizeof e, c->stream));
  return c->fail(SRB_ERR_STATE, "peer barrier timed out: a rank did not deliver its gradient bands; "
                                "the gradient and cost of this evaluation are not valid");
}

srb_status srb_peer_status(srb_ctx* c) {
  if (!c) return SRB_ERR_INVALID;
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  return peer_status(c);
}

srb_status srb_memcpy_d2h(srb_ctx* c, void* dst_host, const void* src_dev, unsigned long long bytes) {
  if (!c || !dst_host || !src_dev) return SRB_ERR_INVALID;
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return peer_status(c);
}

srb_status srb_set_profiling(srb_ctx* c, int on) {
  if (!c) return SRB_ERR_INVALID;
  c->profiling = on != 0;
  return SRB_OK;
}

static srb_status term_host(srb_ctx* c, const double* x_host, double* g_accum, double* cost, bool data) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host) return c->fail(SRB_ERR_INVALID, "null estimate");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const size_t bytes = c->n_active() * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = true;
  if (g_accum) SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_grad, g_accum, bytes, cudaMemcpyHostToDevice, c->stream));
  // single terms always run the reference-order kernels (they ADD into the caller's gradient in
  // the reference's operation order)
  const int saved = c->path;
  c->path = SRB_PATH_REFERENCE_ORDER;
  srb_status st = eval_core(c, c->d_x, g_accum ? c->d_grad : nullptr, nullptr, data, !data, true);
  c->path = saved;
  if (st != SRB_OK) return st;
  if (g_accum) SRB_CUDA_CHECK(c, cudaMemcpyAsync(g_accum, c->d_grad, bytes, cudaMemcpyDeviceToHost, c->stream));
  return fetch_cost(c, data ? 0 : 1, cost);
}

srb_status srb_data_term(srb_ctx* c, const double* x_host, double* g_accum, double* cost) {
  return term_host(c, x_host, g_accum, cost, true);
}
srb_status srb_irls_term(srb_ctx* c, const double* x_host, double* g_accum, double* cost) {
  return term_host(c, x_host, g_accum, cost, false);
}

srb_status srb_reg_apply(srb_ctx* c, const double* x_host, int C, double* values_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host || !values_out) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (c->reg_kind == SRB_REG_NONE) return c->fail(SRB_ERR_STATE, "no regularizer configured");
  if (C < 1 || C > c->g.Ct) return c->fail(SRB_ERR_INVALID, "invalid channel count");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  srb_status st = dev_alloc(c, &c->d_vals, (size_t)c->g.Ct * c->P);
  if (st != SRB_OK) return st;
  const size_t bytes = (size_t)C * c->P * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = false;
  k_reg_values<0><<<grid2d(c->g.W, c->g.H, C), dim3(32, 8), 0, c->stream>>>(make_reg_params(c, C), c->d_x, c->d_vals);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(values_out, c->d_vals, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_reg_apply_diff(srb_ctx* c, const double* x_host, const double* cst_host, int C,
                              double* values_out, double* partials_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!x_host || !cst_host || !values_out || !partials_out) return c->fail(SRB_ERR_INVALID, "null buffer");
  if (c->reg_kind == SRB_REG_NONE) return c->fail(SRB_ERR_STATE, "no regularizer configured");
  if (C < 1 || C > c->g.Ct) return c->fail(SRB_ERR_INVALID, "invalid channel count");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  srb_status st = dev_alloc(c, &c->d_vals, (size_t)c->g.Ct * c->P);
  if (st != SRB_OK) return st;
  if ((st = dev_alloc(c, &c->d_aux, (size_t)c->g.Ct * c->P)) != SRB_OK) return st;
  const size_t bytes = (size_t)C * c->P * sizeof(double);
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
  c->x_resident = false;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_aux, cst_host, bytes, cudaMemcpyHostToDevice, c->stream));
  const RegParams R = make_reg_params(c, C);
  k_reg_values<0><<<grid2d(c->g.W, c->g.H, C), dim3(32, 8), 0, c->stream>>>(R, c->d_x, c->d_vals);
  // partials go to d_grad (scratch here)
  k_reg_partials<0><<<grid2d(c->g.W, c->g.H, C), dim3(32, 8), 0, c->stream>>>(
      R, c->d_x, c->d_vals, c->d_aux, 0.0, 0, c->g.H, c->d_grad, nullptr);
  c->timing.kernel_launches += 2;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(values_out, c->d_vals, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(partials_out, c->d_grad, bytes, cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_forward(srb_ctx* c, int frame, const double* hr_host, int H, int W, double* lr_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!hr_host || !lr_out || H <= 0 || W <= 0) return c->fail(SRB_ERR_INVALID, "bad image");
  if (frame < 0 || frame >= c->g.N) return c->fail(SRB_ERR_INVALID, "frame index out of range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  int h, w;
  lr_size(c->g.s, H, W, &h, &w);
  if (h <= 0 || w <= 0) return c->fail(SRB_ERR_INVALID, "image smaller than the downsampling scale");
  std::vector<int> tab((size_t)h + w + H + 1);
  int* src_r = tab.data();
  int* src_c = src_r + h;
  int* rowY = src_c + w;
  for (int q = 0; q < h; ++q) src_r[q] = nearest_index(q, H, h);
  for (int q = 0; q < w; ++q) src_c[q] = nearest_index(q, W, w);
  const WarpQ wq = quantize_warp(c->shifts_h[2 * frame], c->shifts_h[2 * frame + 1], H, rowY);
  rowY[H] = wq.nX;
  TempBuf t_tab, t_in, t_out;
  SRB_CUDA_CHECK(c, cudaMalloc(&t_tab.p, tab.size() * sizeof(int)));
  SRB_CUDA_CHECK(c, cudaMalloc(&t_in.p, (size_t)H * W * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&t_out.p, (size_t)h * w * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(t_tab.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(t_in.p, hr_host, (size_t)H * W * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  GenericParams P = make_params(c, false);
  P.H = H; P.W = W; P.h = h; P.w = w; P.N = 1; P.Ca = 1; P.Ct = 1; P.c0 = 0;
  P.src_r = t_tab.get<int>(); P.src_c = P.src_r + h; P.rowY = P.src_c + w; P.nX = P.rowY + H;
  k_forward_generic<0><<<grid2d(w, h, 1), dim3(32, 8), 0, c->stream>>>(P, t_in.get<double>(), nullptr,
                                                                     t_out.get<double>(), nullptr);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(lr_out, t_out.p, (size_t)h * w * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_forward_all(srb_ctx* c, const double* hr_host, double* lr_out_host) {
  if (!c) return SRB_ERR_INVALID;
  if (!hr_host || !lr_out_host) return c->fail(SRB_ERR_INVALID, "null buffer");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const Geometry& G = c->g;
  const size_t n_hr = (size_t)G.Ct * c->P, n_lr = (size_t)G.N * G.Ct * c->p;
  srb_status st = dev_alloc(c, &c->d_pooled, n_lr);  // scratch of the reference-order path, same shape
  if (st != SRB_OK) return st;
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(c->d_x, hr_host, n_hr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->x_resident = false;
  GenericParams P = make_params(c, false);
  P.Ca = G.Ct;  // every channel, whatever the active channel range is
  P.c0 = 0;
  k_forward_generic<0><<<grid2d(G.w, G.h, G.N * G.Ct), dim3(32, 8), 0, c->stream>>>(P, c->d_x, nullptr, c->d_pooled, nullptr);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(lr_out_host, c->d_pooled, n_lr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

srb_status srb_transpose(srb_ctx* c, int frame, const double* lr_host, int h, int w, double* hr_out) {
  if (!c) return SRB_ERR_INVALID;
  if (!lr_host || !hr_out || h <= 0 || w <= 0) return c->fail(SRB_ERR_INVALID, "bad image");
  if (frame < 0 || frame >= c->g.N) return c->fail(SRB_ERR_INVALID, "frame index out of range");
  SRB_CUDA_CHECK(c, cudaSetDevice(c->device));
  const int H = h * c->g.s, W = w * c->g.s;
  std::vector<int> tab((size_t)H + 1);
  const WarpQ wq = quantize_warp(-c->shifts_h[2 * frame], -c->shifts_h[2 * frame + 1], H, tab.data());
  tab[H] = wq.nX;
  TempBuf t_tab, t_in, t_out;
  SRB_CUDA_CHECK(c, cudaMalloc(&t_tab.p, tab.size() * sizeof(int)));
  SRB_CUDA_CHECK(c, cudaMalloc(&t_in.p, (size_t)h * w * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMalloc(&t_out.p, (size_t)H * W * sizeof(double)));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(t_tab.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(t_in.p, lr_host, (size_t)h * w * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  GenericParams P = make_params(c, true);
  P.H = H; P.W = W; P.h = h; P.w = w; P.N = 1; P.Ca = 1; P.Ct = 1; P.c0 = 0;
  P.src_r = nullptr; P.src_c = nullptr; P.rowY = t_tab.get<int>(); P.nX = P.rowY + H;
  k_adjoint_generic<<<grid2d(W, H, 1), dim3(32, 8), 0, c->stream>>>(P, t_in.get<double>(), t_out.get<double>(), 1.0, 0);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  SRB_CUDA_CHECK(c, cudaMemcpyAsync(hr_out, t_out.p, (size_t)H * W * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return SRB_OK;
}

// ---- plumbing -----------------------------------------------------------------------------------
srb_status srb_pin_host(void* ptr, unsigned long long bytes) {
  if (!ptr || !bytes) return SRB_ERR_INVALID;
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);  // pinned for every device (srb_multi_*)
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return SRB_ERR_CUDA;
  }
  return SRB_OK;
}
srb_status srb_unpin_host(void* ptr) {
  if (!ptr) return SRB_ERR_INVALID;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return SRB_ERR_CUDA;
  }
  return SRB_OK;
}
void* srb_stream(srb_ctx* c) { return c ? (void*)c->stream : nullptr; }
double* srb_dev_x(srb_ctx* c) { return c ? c->d_x : nullptr; }
double* srb_dev_gradient(srb_ctx* c) { return c ? c->d_grad : nullptr; }
srb_status srb_synchronize(srb_ctx* c) {
  if (!c) return SRB_ERR_INVALID;
  SRB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
  return peer_status(c);
}
srb_status srb_get_timing(srb_ctx* c, srb_timing* out) {
  if (!c || !out) return SRB_ERR_INVALID;
  const double nf = (double)c->g.N / ((double)c->g.s * c->g.s);
  c->timing.algorithmic_bytes_per_eval =
      (unsigned long long)(8.0 * (double)c->n_active() * ((reg_active(c) ? 3.0 : 2.0) + nf));
  if (c->profiling) {
    float ms = 0.f;
    if (cudaEventQuery(c->ev[5]) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess)
      c->timing.last_main_kernel_ms = ms;
    else
      (void)cudaGetLastError();
  }
  *out = c->timing;
  return SRB_OK;
}

}  // extern "C"

#include "srb_multi.cuh"  // single-process multi-GPU form (needs everything above)
#include "srb_multi_solver.cuh"  // the device-resident solver on several devices (row bands)
#include "srb_frontend.cuh"  // data generation, initial estimate, scores, ENVI + spectral PCA
