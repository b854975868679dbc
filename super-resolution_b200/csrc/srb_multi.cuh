// srb_multi.cuh -- ONE host thread, several B200s: the multi-GPU form of the hot path behind the C ABI
// (include/srb200.h, srb_multi_*).  Included at the end of srb_api.cu.
//
// The reference is a single process whose solver calls ObjectiveFunction::ComputeAllTerms from one
// thread (irls_map_solver.cpp:192-265, alglib_objective.cpp:142-152), so the drop-in has to drive every
// GPU from that thread: no torch.distributed, no CUDA IPC -- peer access is enabled directly and
// everything is ordered with streams and events.
//
// Partition: the data term is a sum over LR frames (objective_data_term.cpp:104-114);
// device r holds a contiguous block of the frames (one srb_ctx per device, the same kernels as the single-GPU
// path), the estimate and the IRLS weights are replicated, the regularization term is split by HR row
// bands, and the ONE exchange per evaluation is the sum of the partial gradients, done as a
// reduce-scatter over NVLink (copy engines, band by band, behind the tile kernel's work on the next band)
// into the band each device owns.  srb_multi_eval is built around the PCIe boundary, which is what bounds
// an evaluation driven by a host solver: device r fetches only ITS band of x from the host over ITS PCIe
// link, the bands are all-gathered over NVLink (900 GB/s per direction instead of ~55), every device
// evaluates its frames, and device r returns only its summed band of the gradient -- host traffic per
// link drops by the number of devices.
#pragma once
#include "srb_workers.h"

struct srb_multi {
  int G = 0;
  int partition = SRB_PARTITION_FRAMES;  // SRB_PARTITION_FRAMES | SRB_PARTITION_ROWS
  bool rows_ok = false;                  // rows partition: the current configuration can be cut into row bands
  std::vector<int> dev;
  std::vector<srb_ctx*> rank;
  std::vector<int> frame_begin;          // [G + 1]
  srb_model_desc desc{};                 // the whole model (pointers into the vectors below)
  std::vector<double> psf, shifts;
  size_t lr_plane = 0;                   // h * w
  // The active range is cut into `ngroups` groups of whole channels (no dependency crosses a channel
  // boundary for the data term or 2-D TV), each group into G contiguous bands of gradient units: device o
  // owns elements [band_elem[g][o], band_elem[g][o + 1]) of group g.  Groups flow through the stages
  // H2D -> all-gather -> evaluate -> sum -> D2H as a pipeline.
  static constexpr int kMaxGroups = 4;
  bool bands_valid = false;
  bool pipelined = false;                // every rank can evaluate unit ranges (fused path, no border band)
  int ngroups = 1;
  int band_unit[kMaxGroups][SRB_MAX_PEERS + 1] = {};
  long long band_elem[kMaxGroups][SRB_MAX_PEERS + 1] = {};
  std::vector<cudaStream_t> s_gather;
  std::vector<cudaEvent_t> ev_h2d[kMaxGroups], ev_x[kMaxGroups], ev_part[kMaxGroups], ev_t0, ev_t1;
  std::vector<double*> h_cost;           // pinned, [4] per device
  srb::DeviceWorkers* workers = nullptr; // helper threads of the multi-device solver (srb_multi_solver.cuh), made on first use
  std::string err;
  srb_timing timing{};
  double last_ms[6] = {};                // h2d, all-gather, compute + scatter, sum, d2h, total (device 0's clock)

  srb_status fail(srb_status st, const std::string& m) {
    err = m;
    return st;
  }
};

namespace srb {

#define SRB_MULTI_CHECK(m, call)                                                                 \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      char buf__[512];                                                                           \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
               __FILE__, __LINE__);                                                              \
      return (m)->fail(SRB_ERR_CUDA, buf__);                                                     \
    }                                                                                            \
  } while (0)

// srb_multi_* walk over the devices with cudaSetDevice; the caller's current device is restored on return
// (the host solver that owns the thread may be using CUDA itself).
struct DeviceRestore {
  int dev = -1;
  DeviceRestore() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      dev = -1;
      (void)cudaGetLastError();
    }
  }
  ~DeviceRestore() {
    if (dev >= 0) cudaSetDevice(dev);
  }
};

struct MultiPtrs {
  double* p[SRB_MAX_PEERS];
  long long begin[SRB_MAX_PEERS + 1];
};

// All-gather, pull form: this device copies band s of x from device s's replica (NVLink peer loads) into its
// own replica, for every s != rank.  16-byte accesses; bands start on even elements.
__global__ void __launch_bounds__(256)
k_multi_pull_x(MultiPtrs X, int world, int rank) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int s = 0; s < world; ++s) {
    if (s == rank) continue;
    const long long b = X.begin[s], e = X.begin[s + 1];
    if (((b | e) & 1) == 0) {
      const double2* __restrict__ src = reinterpret_cast<const double2*>(X.p[s] + b);
      double2* __restrict__ dst = reinterpret_cast<double2*>(X.p[rank] + b);
      for (long long i = i0; i < ((e - b) >> 1); i += stride) dst[i] = src[i];
    } else {
      for (long long i = i0; i < e - b; i += stride) X.p[rank][b + i] = X.p[s][b + i];
    }
  }
}

// Reduce-scatter, pull form: out[i] = sum over devices s = 0 .. world-1 (fixed order: deterministic) of
// device s's partial gradient, for this device's band [begin, end); the partials of the other devices are
// read over NVLink, the sum replaces this device's own partial.
__global__ void __launch_bounds__(256)
k_multi_sum_pull(MultiPtrs Gp, int world, int rank, long long begin, long long end) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((begin | end) & 1) == 0) {
    for (long long i = i0; i < ((end - begin) >> 1); i += stride) {
      double2 acc = make_double2(0.0, 0.0);
      for (int s = 0; s < world; ++s) {
        const double2 v = reinterpret_cast<const double2*>(Gp.p[s] + begin)[i];
        acc.x += v.x;
        acc.y += v.y;
      }
      reinterpret_cast<double2*>(Gp.p[rank] + begin)[i] = acc;
    }
  } else {
    for (long long i = i0; i < end - begin; i += stride) {
      double acc = 0.0;
      for (int s = 0; s < world; ++s) acc += Gp.p[s][begin + i];
      Gp.p[rank][begin + i] = acc;
    }
  }
}

// The helper threads of a context (G - 1 of them), made on first use; NULL for one device, with SRB_MULTI_THREADS=0,
// or when the system refuses a thread -- the calling thread then issues every device's work itself.
inline DeviceWorkers* multi_workers(srb_multi* m) {
  const char* e = getenv("SRB_MULTI_THREADS");
  if (m->G <= 1 || (e && atoi(e) == 0)) return nullptr;
  if (!m->workers) {
    try {
      m->workers = new DeviceWorkers(m->G - 1);
    } catch (...) {
      m->workers = nullptr;
    }
  }
  return m->workers;
}

inline srb_status multi_status(srb_multi* m, int r, srb_status st) {
  if (st != SRB_OK) m->err = std::string("device ") + std::to_string(m->dev[r]) + ": " + m->rank[r]->err;
  return st;
}

// Groups and band ownership for the current channel range / regularizer: whole-channel groups cut into
// unit-aligned bands when every rank can evaluate unit ranges; one group with an even element split
// otherwise (the evaluation is then one eval_core per device).
inline srb_status multi_plan_bands(srb_multi* m) {
  const int G = m->G;
  srb_ctx* c0 = m->rank[0];
  m->pipelined = true;
  bool ranges_ok = true;
  for (int r = 0; r < G; ++r) {
    m->pipelined = m->pipelined && units_pipelined(m->rank[r]);
    ranges_ok = ranges_ok && unit_ranges_ok(m->rank[r]);
  }
  const long long n = (long long)c0->n_active();
  // rows partition: unit-aligned bands whenever unit ranges can be evaluated (a border band is cut by rows too)
  if (m->partition == SRB_PARTITION_ROWS ? ranges_ok : m->pipelined) {
    const int Ca = c0->Ca();
    const int tr = tile_rows_per_channel(c0), TH = tile_height(c0);
    static const int env_groups = getenv("SRB_MULTI_GROUPS") ? atoi(getenv("SRB_MULTI_GROUPS")) : 0;
    int ng = env_groups > 0 ? env_groups : 3;
    ng = std::max(1, std::min(std::min(ng, Ca), (int)srb_multi::kMaxGroups));
    if (!host_slices_ok(c0)) ng = 1;  // 3-D TV couples the channels: the whole estimate before any evaluation
    m->ngroups = ng;
    auto first_elem = [&](int u) -> long long {
      if (u >= tr * Ca) return n;
      const int ch = u / tr, t = u - ch * tr;
      const int row = t * TH < c0->g.H ? t * TH : c0->g.H;
      return (long long)ch * (long long)c0->P + (long long)row * c0->g.W;
    };
    for (int g = 0; g < ng; ++g) {
      const int u0 = (int)((long long)Ca * g / ng) * tr, u1 = (int)((long long)Ca * (g + 1) / ng) * tr;
      for (int o = 0; o <= G; ++o) {
        const int u = u0 + (int)((long long)(u1 - u0) * o / G);
        m->band_unit[g][o] = u;
        m->band_elem[g][o] = first_elem(u);
      }
    }
  } else {
    m->ngroups = 1;
    for (int o = 0; o <= G; ++o) {
      m->band_unit[0][o] = 0;
      m->band_elem[0][o] = o == G ? n : ((n * o / G) & ~1LL);
    }
  }
  m->rows_ok = ranges_ok && host_slices_ok(c0);
  m->bands_valid = true;
  return SRB_OK;
}

// Rows partition (SRB_PARTITION_ROWS): every device holds every frame and evaluates the WHOLE objective on its
// HR row bands -- the gradient band it produces is final, so there is no exchange between the devices at all:
// device r fetches its bands of x (plus the halo rows the PSF / regularizer stencils reach) from the host over
// its own PCIe link, runs the tile kernel on its units and returns its gradient bands; the host adds G partial
// costs.  Per-device compute, H2D and D2H all drop by the number of devices.
inline srb_status multi_eval_rows(srb_multi* m, const double* x_host, double* g_host, double* cost) {
  const int G = m->G, NG = m->ngroups;
  // The devices do not depend on each other here, so with helper threads (srb_workers.h; one fork/join round per
  // call, the helpers sleep in between -- a host solver computes for milliseconds between two evaluations) every
  // device's ~10 runtime calls per group are issued by its own thread; without them the calling thread issues
  // group by group so that every device's pipeline starts early.  Failures are recorded per device.
  struct DevErr {
    srb_status st = SRB_OK;
    std::string msg;
    srb_status fail(srb_status s, const std::string& t) {
      if (st == SRB_OK) {
        st = s;
        msg = t;
      }
      return s;
    }
  };
  DevErr err[SRB_MAX_PEERS];
  auto begin = [&](int r) -> srb_status {
    SRB_MULTI_CHECK(&err[r], cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(&err[r], cudaEventRecord(m->ev_t0[r], m->rank[r]->s_in));
    return SRB_OK;
  };
  auto issue = [&](int r, int g) -> srb_status {
    DevErr* e = &err[r];
    const long long* be = m->band_elem[g];
    srb_ctx* c = m->rank[r];
    SRB_MULTI_CHECK(e, cudaSetDevice(m->dev[r]));
    const bool have = m->band_unit[g][r + 1] > m->band_unit[g][r];
    if (have) {
      // band + halo, clipped to the channels the band touches (no stencil crosses a channel boundary)
      const long long P = (long long)c->P, W = c->g.W, halo = (long long)stencil_halo_rows(c) * W;
      const long long lo = std::max(be[r] - halo, be[r] / P * P);
      const long long hi = std::min(be[r + 1] + halo, (be[r + 1] + P - 1) / P * P);
      SRB_MULTI_CHECK(e, cudaMemcpyAsync(c->d_x + lo, x_host + lo, (size_t)(hi - lo) * sizeof(double),
                                         cudaMemcpyHostToDevice, c->s_in));
    }
    SRB_MULTI_CHECK(e, cudaEventRecord(m->ev_h2d[g][r], c->s_in));
    SRB_MULTI_CHECK(e, cudaStreamWaitEvent(c->stream, m->ev_h2d[g][r], 0));
    if (g == 0) {
      // only this device's units write their cost slots: the others must read as zero
      const TileLayout L = tile_layout(c);
      const size_t need = 2 * L.nblocks + L.nband;
      if (need > c->partial_capacity) {
        if (c->d_partial) cudaFree(c->d_partial);
        c->d_partial = nullptr;
        c->partial_capacity = 0;
        SRB_MULTI_CHECK(e, cudaMalloc((void**)&c->d_partial, need * sizeof(double)));
        c->partial_capacity = need;
      }
      SRB_MULTI_CHECK(e, cudaMemsetAsync(c->d_partial, 0, need * sizeof(double), c->stream));
    }
    if (have) {
      const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
      bool reg_done = false;
      srb_status st = fused_eval_units(c, c->d_x, g_host ? c->d_grad : nullptr, do_reg, m->band_unit[g][r],
                                       m->band_unit[g][r + 1], &reg_done);
      if (st == SRB_OK)
        st = fused_band_units(c, c->d_x, g_host ? c->d_grad : nullptr, m->band_unit[g][r], m->band_unit[g][r + 1]);
      if (st != SRB_OK) return e->fail(st, std::string("device ") + std::to_string(m->dev[r]) + ": " + c->err);
    }
    if (g == NG - 1) {
      srb_status st = fused_eval_finish(c, c->d_x, g_host ? c->d_grad : nullptr, nullptr, /*run_band=*/false);
      if (st != SRB_OK) return e->fail(st, std::string("device ") + std::to_string(m->dev[r]) + ": " + c->err);
      c->timing.num_evals += 1;
      SRB_MULTI_CHECK(e, cudaMemcpyAsync(m->h_cost[r], c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    SRB_MULTI_CHECK(e, cudaEventRecord(m->ev_part[g][r], c->stream));
    if (g_host && have) {
      SRB_MULTI_CHECK(e, cudaStreamWaitEvent(c->s_out, m->ev_part[g][r], 0));
      SRB_MULTI_CHECK(e, cudaMemcpyAsync(g_host + be[r], c->d_grad + be[r], (size_t)(be[r + 1] - be[r]) * sizeof(double),
                                         cudaMemcpyDeviceToHost, c->s_out));
    }
    c->x_resident = false;  // only this device's bands of x are here
    return SRB_OK;
  };
  auto drain = [&](int r) -> srb_status {
    SRB_MULTI_CHECK(&err[r], cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(&err[r], cudaEventRecord(m->ev_t1[r], m->rank[r]->s_out));
    SRB_MULTI_CHECK(&err[r], cudaStreamSynchronize(m->rank[r]->s_out));
    SRB_MULTI_CHECK(&err[r], cudaStreamSynchronize(m->rank[r]->stream));
    return SRB_OK;
  };
  DeviceWorkers* workers = multi_workers(m);
  if (workers) {
    workers->run(G, [&](int r) {
      bool ok = begin(r) == SRB_OK;
      for (int g = 0; g < NG && ok; ++g) ok = issue(r, g) == SRB_OK;
      (void)drain(r);  // whatever was issued has to finish before the buffers are touched again
    });
  } else {
    bool ok = true;
    for (int r = 0; r < G && ok; ++r) ok = begin(r) == SRB_OK;
    for (int g = 0; g < NG && ok; ++g)
      for (int r = 0; r < G && ok; ++r) ok = issue(r, g) == SRB_OK;
    for (int r = 0; r < G; ++r) (void)drain(r);
  }
  for (int r = 0; r < G; ++r)
    if (err[r].st != SRB_OK) return m->fail(err[r].st, err[r].msg);
  double total = 0.0;
  for (int r = 0; r < G; ++r) total += m->h_cost[r][2];  // fixed device order
  if (cost) *cost = total;
  float ms = 0.f;
  if (cudaSetDevice(m->dev[0]) == cudaSuccess && cudaEventElapsedTime(&ms, m->ev_t0[0], m->ev_t1[0]) == cudaSuccess)
    m->last_ms[5] = ms;
  (void)cudaGetLastError();
  m->timing.num_evals += 1;
  return SRB_OK;
}

}  // namespace srb

extern "C" {

const char* srb_multi_last_error(const srb_multi* m) { return m ? m->err.c_str() : "null context"; }
int srb_multi_num_gpus(const srb_multi* m) { return m ? m->G : 0; }
srb_ctx* srb_multi_rank_ctx(srb_multi* m, int r) { return (m && r >= 0 && r < m->G) ? m->rank[r] : nullptr; }

void srb_multi_destroy(srb_multi* m) {
  if (!m) return;
  srb::DeviceRestore restore_device;
  for (int r = 0; r < (int)m->rank.size(); ++r) {
    cudaSetDevice(m->dev[r]);
    if (m->rank[r] && m->rank[r]->stream) cudaStreamSynchronize(m->rank[r]->stream);
    if (r < (int)m->s_gather.size() && m->s_gather[r]) { cudaStreamSynchronize(m->s_gather[r]); cudaStreamDestroy(m->s_gather[r]); }
    auto kill = [&](std::vector<cudaEvent_t>& v) { if (r < (int)v.size() && v[r]) cudaEventDestroy(v[r]); };
    kill(m->ev_t0); kill(m->ev_t1);
    for (int g = 0; g < srb_multi::kMaxGroups; ++g) { kill(m->ev_h2d[g]); kill(m->ev_x[g]); kill(m->ev_part[g]); }
    if (r < (int)m->h_cost.size() && m->h_cost[r]) cudaFreeHost(m->h_cost[r]);
  }
  delete m->workers;  // joins the helper threads
  for (srb_ctx* c : m->rank) srb_destroy(c);
  delete m;
}

srb_status srb_multi_create(const srb_model_desc* d, int n_gpus, const int* devices, srb_multi** out) {
  int partition = SRB_PARTITION_FRAMES;
  if (const char* e = getenv("SRB_MULTI_PARTITION")) partition = (e[0] == 'r' || e[0] == '1') ? SRB_PARTITION_ROWS : SRB_PARTITION_FRAMES;
  return srb_multi_create_partitioned(d, n_gpus, devices, partition, out);
}

srb_status srb_multi_create_partitioned(const srb_model_desc* d, int n_gpus, const int* devices, int partition,
                                        srb_multi** out) {
  if (!out) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  *out = nullptr;
  srb_multi* m = new (std::nothrow) srb_multi();
  if (!m) return SRB_ERR_NOMEM;
  *out = m;  // returned even on failure so the caller can read srb_multi_last_error, then srb_multi_destroy
  if (!d) return m->fail(SRB_ERR_INVALID, "null model description");
  if (n_gpus < 1 || n_gpus > SRB_MAX_PEERS) return m->fail(SRB_ERR_INVALID, "number of GPUs must be 1..8");
  if (d->num_frames <= 0) return m->fail(SRB_ERR_INVALID, "cannot solve with 0 observations");  // map_solver.cpp:56-57
  if (partition != SRB_PARTITION_FRAMES && partition != SRB_PARTITION_ROWS) return m->fail(SRB_ERR_INVALID, "unknown partition");
  m->partition = partition;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return m->fail(SRB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  m->G = n_gpus;
  m->dev.resize(n_gpus);
  for (int r = 0; r < n_gpus; ++r) {
    m->dev[r] = devices ? devices[r] : r;
    if (m->dev[r] < 0 || m->dev[r] >= ndev) return m->fail(SRB_ERR_INVALID, "invalid CUDA device index");
    // SRB_MULTI_SHARE_DEVICES=1 (tests, small boxes): the same physical device may appear more than once; every entry
    // still gets its own context, streams and buffers, so the whole multi-device logic runs on one GPU
    const bool share = getenv("SRB_MULTI_SHARE_DEVICES") && atoi(getenv("SRB_MULTI_SHARE_DEVICES")) != 0;
    for (int q = 0; q < r && !share; ++q)
      if (m->dev[q] == m->dev[r]) return m->fail(SRB_ERR_INVALID, "a CUDA device is listed twice");
  }
  m->desc = *d;
  const int K = d->psf_size;
  if (d->psf && K > 0) m->psf.assign(d->psf, d->psf + (size_t)K * K);
  if (d->shifts) m->shifts.assign(d->shifts, d->shifts + (size_t)2 * d->num_frames);
  m->desc.psf = m->psf.empty() ? nullptr : m->psf.data();
  m->desc.shifts = m->shifts.empty() ? nullptr : m->shifts.data();
  m->lr_plane = (size_t)d->lr_height * d->lr_width;
  // contiguous frame blocks (any partition is valid, the data term is a plain sum over frames); a device
  // beyond the number of frames holds none and contributes only its regularizer band
  m->frame_begin.resize(n_gpus + 1);
  {
    const int base = d->num_frames / n_gpus, extra = d->num_frames % n_gpus;
    int f = 0;
    for (int r = 0; r < n_gpus; ++r) {
      m->frame_begin[r] = f;
      f += base + (r < extra ? 1 : 0);
    }
    m->frame_begin[n_gpus] = f;
  }
  m->rank.assign(n_gpus, nullptr);
  m->h_cost.assign(n_gpus, nullptr);
  m->s_gather.assign(n_gpus, nullptr);
  m->ev_t0.assign(n_gpus, nullptr);
  m->ev_t1.assign(n_gpus, nullptr);
  for (int g = 0; g < srb_multi::kMaxGroups; ++g) {
    m->ev_h2d[g].assign(n_gpus, nullptr);
    m->ev_x[g].assign(n_gpus, nullptr);
    m->ev_part[g].assign(n_gpus, nullptr);
  }
  for (int r = 0; r < n_gpus; ++r) {
    srb_model_desc dr = m->desc;
    if (m->partition == SRB_PARTITION_FRAMES) {
      dr.num_frames = m->frame_begin[r + 1] - m->frame_begin[r];
      dr.shifts = m->desc.shifts ? m->desc.shifts + (size_t)2 * m->frame_begin[r] : nullptr;
    }  // rows: every device holds every frame and evaluates its HR row bands of the whole objective
    srb_status st = srb_create_shard(&dr, m->dev[r], &m->rank[r]);
    if (st != SRB_OK) {
      m->err = std::string("device ") + std::to_string(m->dev[r]) + ": " + (m->rank[r] ? m->rank[r]->err : "out of memory");
      return st;
    }
    const int H = m->rank[r]->g.H;
    const int band = (H + n_gpus - 1) / n_gpus;
    if (m->partition == SRB_PARTITION_FRAMES)
      srb_set_regularizer_rows(m->rank[r], std::min(H, r * band), std::min(H, (r + 1) * band));
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    for (int q = 0; q < n_gpus && m->partition == SRB_PARTITION_FRAMES; ++q) {  // (rows: no device reads another's memory)
      if (q == r || m->dev[q] == m->dev[r]) continue;
      int can = 0;
      SRB_MULTI_CHECK(m, cudaDeviceCanAccessPeer(&can, m->dev[r], m->dev[q]));
      if (!can) return m->fail(SRB_ERR_CUDA, "the GPUs cannot access each other's memory (NVLink / PCIe peer access)");
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SRB_MULTI_CHECK(m, e);
      (void)cudaGetLastError();
    }
    int lo = 0, hi = 0;
    SRB_MULTI_CHECK(m, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SRB_MULTI_CHECK(m, cudaStreamCreateWithPriority(&m->s_gather[r], cudaStreamNonBlocking, hi));
    for (int g = 0; g < srb_multi::kMaxGroups; ++g) {
      SRB_MULTI_CHECK(m, cudaEventCreateWithFlags(&m->ev_h2d[g][r], cudaEventDisableTiming));
      SRB_MULTI_CHECK(m, cudaEventCreateWithFlags(&m->ev_x[g][r], cudaEventDisableTiming));
      SRB_MULTI_CHECK(m, cudaEventCreateWithFlags(&m->ev_part[g][r], cudaEventDisableTiming));
    }
    SRB_MULTI_CHECK(m, cudaEventCreate(&m->ev_t0[r]));
    SRB_MULTI_CHECK(m, cudaEventCreate(&m->ev_t1[r]));
    SRB_MULTI_CHECK(m, cudaMallocHost((void**)&m->h_cost[r], 4 * sizeof(double)));
  }
  return SRB_OK;
}

srb_status srb_multi_set_observations(srb_multi* m, const double* lr_host) {
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  if (!lr_host) return m->fail(SRB_ERR_INVALID, "null observations");
  const size_t per_frame = (size_t)m->desc.num_channels * m->lr_plane;
  for (int r = 0; r < m->G; ++r) {
    const size_t first = m->partition == SRB_PARTITION_FRAMES ? (size_t)m->frame_begin[r] : 0;
    srb_status st = srb_set_observations(m->rank[r], lr_host + first * per_frame);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_channel_range(srb_multi* m, int c0, int c1) {
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_channel_range(m->rank[r], c0, c1);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_regularizer(srb_multi* m, int kind, double lambda, int btv_range, double btv_decay) {
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_regularizer(m->rank[r], kind, lambda, btv_range, btv_decay);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_irls_weights(srb_multi* m, const double* w) {
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  // replicated: every device reads the same host buffer over its own PCIe link, all copies in flight at once
  for (int r = 0; r < m->G; ++r) {
    srb_ctx* c = m->rank[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    if (!w) {
      srb_status st = srb_set_irls_weights(c, nullptr);
      if (st != SRB_OK) return srb::multi_status(m, r, st);
    } else {
      SRB_MULTI_CHECK(m, cudaMemcpyAsync(c->d_w, w, c->n_active() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
  }
  for (int r = 0; r < m->G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(m->rank[r]->stream));
  }
  return SRB_OK;
}

srb_status srb_multi_set_path(srb_multi* m, int path) {
  if (!m) return SRB_ERR_INVALID;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_path(m->rank[r], path);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

// IRLS re-weighting (irls_map_solver.cpp:128-143): every device needs the full weight image; each computes
// it from its replica of x (the regularizer values are local, one streaming pass), no exchange.
srb_status srb_multi_reweight(srb_multi* m, const double* x_host, double* w_out) {
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  if (!x_host) return m->fail(SRB_ERR_INVALID, "null estimate");
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_reweight(m->rank[r], x_host, r == 0 ? w_out : nullptr);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  return SRB_OK;
}

// ObjectiveFunction::ComputeAllTerms over all devices: host x in, host gradient (may be NULL) and cost out.
srb_status srb_multi_eval(srb_multi* m, const double* x_host, double* g_host, double* cost) {
  using namespace srb;
  if (!m) return SRB_ERR_INVALID;
  srb::DeviceRestore restore_device;
  if (!x_host) return m->fail(SRB_ERR_INVALID, "null estimate");
  const int G = m->G;
  for (int r = 0; r < G; ++r)
    if (!m->rank[r]->have_obs) return m->fail(SRB_ERR_STATE, "srb_multi_set_observations has not been called");
  if (G == 1) {
    srb_status st = srb_eval(m->rank[0], x_host, g_host, cost);
    if (st == SRB_OK) m->timing.num_evals += 1;
    return multi_status(m, 0, st);
  }
  if (!m->bands_valid) {
    srb_status st = multi_plan_bands(m);
    if (st != SRB_OK) return st;
  }
  if (m->partition == SRB_PARTITION_ROWS) {
    if (m->rows_ok) return multi_eval_rows(m, x_host, g_host, cost);
    // a configuration that cannot be cut into row bands (border band of special samples, 3-D TV, a model the
    // tile kernel does not cover): every device holds the whole model, so device 0 evaluates it alone
    srb_status st = srb_eval(m->rank[0], x_host, g_host, cost);
    if (st == SRB_OK) m->timing.num_evals += 1;
    return multi_status(m, 0, st);
  }
  const int NG = m->ngroups;
  MultiPtrs X, Gp;
  for (int r = 0; r < G; ++r) {
    X.p[r] = m->rank[r]->d_x;
    Gp.p[r] = m->rank[r]->d_grad;
  }
  // Streams per device: s_in (H2D of its bands), s_gather (pulls the other bands of x over NVLink), stream
  // (the tile kernel), s_out (sums its band over the devices' partials, then D2H).  Groups are issued one
  // after the other on every stream, so group g + 1 is copied in while group g is evaluated and group g - 1
  // returns to the host.
  for (int r = 0; r < G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_t0[r], m->rank[r]->s_in));
  }
  for (int g = 0; g < NG; ++g) {
    const long long* be = m->band_elem[g];
    for (int o = 0; o <= G; ++o) X.begin[o] = be[o];
    // ---- 1. every device fetches its own band of x over its own PCIe link ------------------------------
    for (int r = 0; r < G; ++r) {
      srb_ctx* c = m->rank[r];
      SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
      if (be[r + 1] > be[r])
        SRB_MULTI_CHECK(m, cudaMemcpyAsync(c->d_x + be[r], x_host + be[r], (size_t)(be[r + 1] - be[r]) * sizeof(double),
                                           cudaMemcpyHostToDevice, c->s_in));
      SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_h2d[g][r], c->s_in));
    }
    // ---- 2. all-gather over NVLink: every device pulls the other devices' bands into its replica ----------
    for (int r = 0; r < G; ++r) {
      srb_ctx* c = m->rank[r];
      SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
      for (int q = 0; q < G; ++q) SRB_MULTI_CHECK(m, cudaStreamWaitEvent(m->s_gather[r], m->ev_h2d[g][q], 0));
      const long long len = be[G] - be[0];
      const int blocks = (int)std::max<long long>(1, std::min<long long>((len / 2 + 255) / 256, (long long)c->num_sms * 4));
      k_multi_pull_x<<<blocks, 256, 0, m->s_gather[r]>>>(X, G, r);
      c->timing.kernel_launches += 1;
      SRB_MULTI_CHECK(m, cudaGetLastError());
      SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_x[g][r], m->s_gather[r]));
    }
    // ---- 3. partial objective of every device's frames over the group ------------------------------------
    for (int r = 0; r < G; ++r) {
      srb_ctx* c = m->rank[r];
      SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
      SRB_MULTI_CHECK(m, cudaStreamWaitEvent(c->stream, m->ev_x[g][r], 0));
      if (m->pipelined) {
        const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
        bool reg_done = false;
        if (m->band_unit[g][G] > m->band_unit[g][0]) {
          srb_status st = fused_eval_units(c, c->d_x, c->d_grad, do_reg, m->band_unit[g][0], m->band_unit[g][G], &reg_done);
          if (st != SRB_OK) return multi_status(m, r, st);
        }
        if (g == NG - 1) {
          srb_status st = fused_eval_finish(c, c->d_x, c->d_grad, nullptr);
          if (st != SRB_OK) return multi_status(m, r, st);
          c->timing.num_evals += 1;
        }
      } else {
        srb_status st = eval_core(c, c->d_x, c->d_grad, nullptr, true, true, false);
        if (st != SRB_OK) return multi_status(m, r, st);
      }
      SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_part[g][r], c->stream));
      if (g == NG - 1)
        SRB_MULTI_CHECK(m, cudaMemcpyAsync(m->h_cost[r], c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    // ---- 4. every owner sums its band over the devices' partials (NVLink peer loads, fixed order) and
    //         returns it over its own PCIe link -----------------------------------------------------------
    if (g_host) {
      for (int o = 0; o < G; ++o) {
        srb_ctx* c = m->rank[o];
        if (be[o + 1] <= be[o]) continue;
        SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[o]));
        for (int q = 0; q < G; ++q) SRB_MULTI_CHECK(m, cudaStreamWaitEvent(c->s_out, m->ev_part[g][q], 0));
        const long long len = be[o + 1] - be[o];
        const int blocks = (int)std::max<long long>(1, std::min<long long>((len / 2 + 255) / 256, (long long)c->num_sms * 4));
        k_multi_sum_pull<<<blocks, 256, 0, c->s_out>>>(Gp, G, o, be[o], be[o + 1]);
        c->timing.kernel_launches += 1;
        SRB_MULTI_CHECK(m, cudaGetLastError());
        SRB_MULTI_CHECK(m, cudaMemcpyAsync(g_host + be[o], c->d_grad + be[o], (size_t)len * sizeof(double),
                                           cudaMemcpyDeviceToHost, c->s_out));
      }
    }
  }
  for (int r = 0; r < G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_t1[r], m->rank[r]->s_out));
  }
  // every stream of every device drains before the buffers are touched again: a device's partial gradient
  // and its replica of x are read by its peers
  for (int r = 0; r < G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(m->rank[r]->s_out));
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(m->s_gather[r]));
  }
  double total = 0.0;
  for (int r = 0; r < G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(m->rank[r]->stream));
    total += m->h_cost[r][2];  // fixed device order
  }
  if (cost) *cost = total;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, m->ev_t0[0], m->ev_t1[0]) == cudaSuccess) m->last_ms[5] = ms;
  (void)cudaGetLastError();
  m->timing.num_evals += 1;
  return SRB_OK;
}

srb_status srb_multi_get_timing(srb_multi* m, srb_timing* out) {
  if (!m || !out) return SRB_ERR_INVALID;
  srb_timing t{};
  for (int r = 0; r < m->G; ++r) {
    srb_timing tr{};
    srb_get_timing(m->rank[r], &tr);
    t.kernel_launches += tr.kernel_launches;
    t.algorithmic_bytes_per_eval += tr.algorithmic_bytes_per_eval;
  }
  t.num_evals = m->timing.num_evals;
  t.last_eval_kernel_ms = m->last_ms[5];
  *out = t;
  return SRB_OK;
}

}  // extern "C"
