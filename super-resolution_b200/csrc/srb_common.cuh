// srb_common.cuh -- shared declarations of the B200 MAP-gradient engine (libsrb200.so).
//
// Data layout in HBM (all fp64, planar row-major, identical to the reference's host layout,
// src/util/util.cpp:81-89):
//   x, gradient, IRLS weights : [Ca][H][W]          (Ca = active channel range, c1-c0)
//   observations              : [N][Ct][h][w]       (LR resolution, uploaded once)
//   pooled residuals (scratch): [N][Ca][h][w]       (reference-order path only)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/srb200.h"

namespace srb {

// Quantised translation of one cv::warpAffine call (motion_module.cpp:18-24).  OpenCV walks the
// destination, derives the source coordinate in 1/1024 px fixed point, adds 16 and drops to 1/32
// px.  For destination column x the source is X = 32*x + nX (integer part X>>5, bilinear weight
// (X&31)/32); the row value is rounded per destination row, so it lives in a per-row table
// Y[row] (same packing).  `uniform` says Y[row] == 32*row + nY for every row (true unless a shift
// sits within ~1e-10 of a quantisation boundary) -- the fused kernel needs that.
struct WarpQ {
  int nX;
  int nY;
  bool uniform;
};

struct Geometry {
  int H, W, h, w, s, K, hk;  // HR size, LR size, scale, PSF side, PSF half width
  int N, Ct;                 // frames held by this context, channels per observation
};

// Device-side parameter block of the reference-order kernels.
struct GenericParams {
  int H, W, h, w, s, K, hk;
  int N;           // frames
  int Ca;          // active channels (x / gradient planes)
  int Ct, c0;      // observation channel count and first active channel
  const int* src_r;  // [h] decimation row map (cv::resize INTER_NEAREST)
  const int* src_c;  // [w] decimation column map
  const double* psf;   // [K*K] blur_kernel_
  const int* rowY;     // [N][H] per destination row: quantised source row (fixed point, 1/32 px)
  const int* nX;       // [N]
};

}  // namespace srb

struct srb_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  srb::Geometry g{};
  size_t P = 0, p = 0;  // HR / LR pixels per channel
  bool has_motion = false, has_blur = false;

  std::vector<double> psf_h;           // K*K (K=1, {1.0} when the model has no blur)
  std::vector<double> shifts_h;        // 2N
  std::vector<srb::WarpQ> warp_fwd, warp_tr;
  bool warps_uniform = true;
  bool warps_integer = true;

  // device constants / tables
  double* d_psf = nullptr;
  int *d_src_r = nullptr, *d_src_c = nullptr;
  int *d_rowY_fwd = nullptr, *d_rowY_tr = nullptr, *d_nX_fwd = nullptr, *d_nX_tr = nullptr;

  // observations and state
  double* d_y = nullptr;
  bool have_obs = false;
  bool x_resident = false;  // d_x holds the caller's whole estimate of the active channel range (srb_reweight(NULL))
  int c0 = 0, c1 = 0;
  int reg_kind = SRB_REG_NONE;
  double lambda = 0.0;
  int btv_R = 3;
  double btv_decay = 0.5;
  double* d_decay = nullptr;  // [2R+1] std::pow(decay, i+j) computed on the host
  std::vector<double> decay_h;  // the same table on the host (kernel parameter of k_btv_tile)
  double* d_w = nullptr;      // IRLS weights [Ct][H][W]
  int reg_row0 = 0, reg_row1 = 0;
  int path = SRB_PATH_AUTO;
  bool strict_cost = false;   // reference-order kernels sum their costs sequentially in the reference's order
  double* d_resid = nullptr;  // [N][Ct][h][w] raw residuals (strict cost only)
  void* fused = nullptr;  // srb::FusedState (srb_kernels_fused.cuh)

  // work buffers
  double* d_x = nullptr;      // [Ct*P]
  double* d_grad = nullptr;   // [Ct*P + 1]
  double* d_pooled = nullptr; // [N][Ct][h][w]
  double* d_vals = nullptr;   // [Ct*P] regularizer values
  double* d_aux = nullptr;    // [Ct*P] second scratch plane (constants / partials)
  double* d_partial = nullptr;  // per-block cost partial sums
  size_t partial_capacity = 0;
  double* cg_store = nullptr;   // solver vectors + reduction slots of the device-resident solver, kept between solves
  size_t cg_store_doubles = 0;
  double* cg_h_out = nullptr;   // pinned scalar mirror of the solver
  double* d_cost = nullptr;   // [4] data cost, reg cost, total, spare
  double* h_cost = nullptr;   // pinned mirror of d_cost

  // multi-GPU peer state (srb_peer_*): slot arrays / gradient buffers of all ranks, mapped through
  // CUDA IPC; `token` is the pseudo gradient pointer that selects scatter mode in fused_eval_units
  struct Peer {
    bool active = false;
    int rank = 0, world = 1;
    long long band_cap = 0;        // doubles per slot (largest band)
    int band_unit[9] = {};         // rank o owns units [band_unit[o], band_unit[o+1]) ...
    long long band_elem[9] = {};   // ... = gradient elements [band_elem[o], band_elem[o+1])
    double* slots[8] = {};         // slot array bases (index = owner rank); [rank] is local
    double* out[8] = {};           // gradient(+cost, +flags) buffers; [rank] is local
    unsigned long long epoch = 0;  // evaluation counter, identical on every rank
    int* d_err = nullptr;          // [0] set by the waits on timeout, [1] = last-block counter of k_sum_gather
    cudaStream_t s_copy[2] = {nullptr, nullptr};  // DMA pushes of finished bands to their owners
    cudaEvent_t ev_band[8] = {}, ev_copy[2] = {};
  } peer;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // host<->device pipeline of srb_eval: copy-in / copy-out streams and per-chunk events
  static constexpr int kMaxPipe = 16;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kMaxPipe] = {}, ev_k[kMaxPipe] = {}, ev_pipe[4] = {};
  int pipe_chunks = 16;
  bool timing_valid = false;
  bool profiling = false;  // record events around the dominant kernel (srb_set_profiling)
  srb_timing timing{};
  std::string err;

  srb_status fail(srb_status st, const std::string& msg) {
    err = msg;
    return st;
  }
  int Ca() const { return c1 - c0; }
  size_t n_active() const { return (size_t)(c1 - c0) * P; }
};

#define SRB_CUDA_CHECK(ctx, call)                                                              \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      char buf__[512];                                                                         \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
               __FILE__, __LINE__);                                                            \
      return (ctx)->fail(SRB_ERR_CUDA, buf__);                                                 \
    }                                                                                          \
  } while (0)

namespace srb {

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the thread block; result valid in thread 0.  blockDim.x*blockDim.y*blockDim.z <= 1024.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[32];
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, wid = tid >> 5;
  v = warp_sum(v);
  __syncthreads();  // protects warp_part against a previous call
  if (lane == 0) warp_part[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = (lane < (nthreads + 31) / 32) ? warp_part[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// Deterministic final reduction of per-block partials: out[slot] = sum(partial[0..n)).
__global__ void k_reduce_partials(const double* __restrict__ partial, size_t n,
                                  double* __restrict__ out, int slot) {
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[slot] = acc;
}

}  // namespace srb
