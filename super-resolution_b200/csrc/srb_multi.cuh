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

struct srb_multi {
  int G = 0;
  std::vector<int> dev;
  std::vector<srb_ctx*> rank;
  std::vector<int> frame_begin;          // [G + 1]
  srb_model_desc desc{};                 // the whole model (pointers into the vectors below)
  std::vector<double> psf, shifts;
  size_t lr_plane = 0;                   // h * w
  // ownership of the active gradient: device o owns elements [band_elem[o], band_elem[o + 1])
  bool bands_valid = false;
  bool pipelined = false;                // every rank can evaluate unit ranges (fused path, no border band)
  int band_unit[SRB_MAX_PEERS + 1] = {};
  long long band_elem[SRB_MAX_PEERS + 1] = {};
  long long band_cap = 0;
  std::vector<double*> slots;            // per device: [G][band_cap] incoming partial bands
  long long slots_cap = 0;
  std::vector<cudaStream_t> s_gather, s_push[2];
  std::vector<cudaEvent_t> ev_h2d, ev_x, ev_band[SRB_MAX_PEERS], ev_pushed, ev_sum, ev_t0, ev_t1;
  std::vector<double*> h_cost;           // pinned, [4] per device
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

// own += slots[s] for s != rank, in rank order: out = sum_s contribution_s (fixed order: deterministic)
__global__ void __launch_bounds__(256)
k_multi_sum_band(double* __restrict__ own, const double* __restrict__ slots, long long cap, long long len,
                 int world, int rank) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    double acc = 0.0;
    for (int s = 0; s < world; ++s) acc += s == rank ? own[i] : slots[(long long)s * cap + i];
    own[i] = acc;
  }
}

inline srb_status multi_status(srb_multi* m, int r, srb_status st) {
  if (st != SRB_OK) m->err = std::string("device ") + std::to_string(m->dev[r]) + ": " + m->rank[r]->err;
  return st;
}

// Band ownership for the current channel range / regularizer: unit-aligned when every rank can evaluate
// unit ranges (then finished bands leave while the next ones are computed), even element split otherwise.
inline srb_status multi_plan_bands(srb_multi* m) {
  const int G = m->G;
  srb_ctx* c0 = m->rank[0];
  m->pipelined = true;
  for (int r = 0; r < G; ++r) m->pipelined = m->pipelined && units_pipelined(m->rank[r]);
  const long long n = (long long)c0->n_active();
  if (m->pipelined) {
    peer_bands(c0, G, m->band_unit, m->band_elem, &m->band_cap);
  } else {
    m->band_cap = 0;
    for (int o = 0; o <= G; ++o) {
      m->band_unit[o] = 0;
      m->band_elem[o] = o == G ? n : ((n * o / G) & ~1LL);
      if (o > 0) m->band_cap = std::max(m->band_cap, m->band_elem[o] - m->band_elem[o - 1]);
    }
  }
  if (m->band_cap > m->slots_cap) {
    for (int r = 0; r < G; ++r) {
      SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
      if (m->slots[r]) cudaFree(m->slots[r]);
      m->slots[r] = nullptr;
      if (cudaMalloc((void**)&m->slots[r], (size_t)G * (size_t)m->band_cap * sizeof(double)) != cudaSuccess) {
        (void)cudaGetLastError();
        return m->fail(SRB_ERR_NOMEM, "cudaMalloc failed (multi-GPU band slots)");
      }
    }
    m->slots_cap = m->band_cap;
  }
  m->bands_valid = true;
  return SRB_OK;
}

}  // namespace srb

extern "C" {

const char* srb_multi_last_error(const srb_multi* m) { return m ? m->err.c_str() : "null context"; }
int srb_multi_num_gpus(const srb_multi* m) { return m ? m->G : 0; }
srb_ctx* srb_multi_rank_ctx(srb_multi* m, int r) { return (m && r >= 0 && r < m->G) ? m->rank[r] : nullptr; }

void srb_multi_destroy(srb_multi* m) {
  if (!m) return;
  for (int r = 0; r < (int)m->rank.size(); ++r) {
    cudaSetDevice(m->dev[r]);
    if (m->rank[r] && m->rank[r]->stream) cudaStreamSynchronize(m->rank[r]->stream);
    if (r < (int)m->s_gather.size() && m->s_gather[r]) { cudaStreamSynchronize(m->s_gather[r]); cudaStreamDestroy(m->s_gather[r]); }
    for (auto& sp : m->s_push)
      if (r < (int)sp.size() && sp[r]) { cudaStreamSynchronize(sp[r]); cudaStreamDestroy(sp[r]); }
    auto kill = [&](std::vector<cudaEvent_t>& v) { if (r < (int)v.size() && v[r]) cudaEventDestroy(v[r]); };
    kill(m->ev_h2d); kill(m->ev_x); kill(m->ev_pushed); kill(m->ev_sum); kill(m->ev_t0); kill(m->ev_t1);
    for (auto& v : m->ev_band) kill(v);
    if (r < (int)m->slots.size() && m->slots[r]) cudaFree(m->slots[r]);
    if (r < (int)m->h_cost.size() && m->h_cost[r]) cudaFreeHost(m->h_cost[r]);
  }
  for (srb_ctx* c : m->rank) srb_destroy(c);
  delete m;
}

srb_status srb_multi_create(const srb_model_desc* d, int n_gpus, const int* devices, srb_multi** out) {
  if (!out) return SRB_ERR_INVALID;
  *out = nullptr;
  srb_multi* m = new (std::nothrow) srb_multi();
  if (!m) return SRB_ERR_NOMEM;
  *out = m;  // returned even on failure so the caller can read srb_multi_last_error, then srb_multi_destroy
  if (!d) return m->fail(SRB_ERR_INVALID, "null model description");
  if (n_gpus < 1 || n_gpus > SRB_MAX_PEERS) return m->fail(SRB_ERR_INVALID, "number of GPUs must be 1..8");
  if (d->num_frames <= 0) return m->fail(SRB_ERR_INVALID, "cannot solve with 0 observations");  // map_solver.cpp:56-57
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
    for (int q = 0; q < r; ++q)
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
  m->slots.assign(n_gpus, nullptr);
  m->h_cost.assign(n_gpus, nullptr);
  m->s_gather.assign(n_gpus, nullptr);
  for (auto& sp : m->s_push) sp.assign(n_gpus, nullptr);
  for (auto* v : {&m->ev_h2d, &m->ev_x, &m->ev_pushed, &m->ev_sum, &m->ev_t0, &m->ev_t1}) v->assign(n_gpus, nullptr);
  for (auto& v : m->ev_band) v.assign(n_gpus, nullptr);
  for (int r = 0; r < n_gpus; ++r) {
    srb_model_desc dr = m->desc;
    dr.num_frames = m->frame_begin[r + 1] - m->frame_begin[r];
    dr.shifts = m->desc.shifts ? m->desc.shifts + (size_t)2 * m->frame_begin[r] : nullptr;
    srb_status st = srb_create_shard(&dr, m->dev[r], &m->rank[r]);
    if (st != SRB_OK) {
      m->err = std::string("device ") + std::to_string(m->dev[r]) + ": " + (m->rank[r] ? m->rank[r]->err : "out of memory");
      return st;
    }
    const int H = m->rank[r]->g.H;
    const int band = (H + n_gpus - 1) / n_gpus;
    srb_set_regularizer_rows(m->rank[r], std::min(H, r * band), std::min(H, (r + 1) * band));
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    for (int q = 0; q < n_gpus; ++q) {
      if (q == r) continue;
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
    for (auto& sp : m->s_push) SRB_MULTI_CHECK(m, cudaStreamCreateWithPriority(&sp[r], cudaStreamNonBlocking, hi));
    for (auto* v : {&m->ev_h2d, &m->ev_x, &m->ev_pushed, &m->ev_sum})
      SRB_MULTI_CHECK(m, cudaEventCreateWithFlags(&(*v)[r], cudaEventDisableTiming));
    SRB_MULTI_CHECK(m, cudaEventCreate(&m->ev_t0[r]));
    SRB_MULTI_CHECK(m, cudaEventCreate(&m->ev_t1[r]));
    for (auto& v : m->ev_band) SRB_MULTI_CHECK(m, cudaEventCreateWithFlags(&v[r], cudaEventDisableTiming));
    SRB_MULTI_CHECK(m, cudaMallocHost((void**)&m->h_cost[r], 4 * sizeof(double)));
  }
  return SRB_OK;
}

srb_status srb_multi_set_observations(srb_multi* m, const double* lr_host) {
  if (!m) return SRB_ERR_INVALID;
  if (!lr_host) return m->fail(SRB_ERR_INVALID, "null observations");
  const size_t per_frame = (size_t)m->desc.num_channels * m->lr_plane;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_observations(m->rank[r], lr_host + (size_t)m->frame_begin[r] * per_frame);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_channel_range(srb_multi* m, int c0, int c1) {
  if (!m) return SRB_ERR_INVALID;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_channel_range(m->rank[r], c0, c1);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_regularizer(srb_multi* m, int kind, double lambda, int btv_range, double btv_decay) {
  if (!m) return SRB_ERR_INVALID;
  for (int r = 0; r < m->G; ++r) {
    srb_status st = srb_set_regularizer(m->rank[r], kind, lambda, btv_range, btv_decay);
    if (st != SRB_OK) return srb::multi_status(m, r, st);
  }
  m->bands_valid = false;
  return SRB_OK;
}

srb_status srb_multi_set_irls_weights(srb_multi* m, const double* w) {
  if (!m) return SRB_ERR_INVALID;
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
  if (!x_host) return m->fail(SRB_ERR_INVALID, "null estimate");
  const int G = m->G;
  for (int r = 0; r < G; ++r)
    if (!m->rank[r]->have_obs) return m->fail(SRB_ERR_STATE, "srb_multi_set_observations has not been called");
  if (G == 1) {
    srb_status st = srb_eval(m->rank[0], x_host, g_host, cost);
    return multi_status(m, 0, st);
  }
  if (!m->bands_valid) {
    srb_status st = multi_plan_bands(m);
    if (st != SRB_OK) return st;
  }
  const long long* be = m->band_elem;
  // ---- 1. every device fetches its own band of x over its own PCIe link --------------------------------
  for (int r = 0; r < G; ++r) {
    srb_ctx* c = m->rank[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_t0[r], c->s_in));
    if (be[r + 1] > be[r])
      SRB_MULTI_CHECK(m, cudaMemcpyAsync(c->d_x + be[r], x_host + be[r], (size_t)(be[r + 1] - be[r]) * sizeof(double),
                                         cudaMemcpyHostToDevice, c->s_in));
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_h2d[r], c->s_in));
  }
  // ---- 2. all-gather of the bands over NVLink: device r pushes its band into every replica ---------------
  for (int r = 0; r < G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    SRB_MULTI_CHECK(m, cudaStreamWaitEvent(m->s_gather[r], m->ev_h2d[r], 0));
    if (be[r + 1] > be[r])
      for (int i = 1; i < G; ++i) {
        const int q = (r + i) % G;   // staggered destinations: every link carries one band at a time
        SRB_MULTI_CHECK(m, cudaMemcpyPeerAsync(m->rank[q]->d_x + be[r], m->dev[q], m->rank[r]->d_x + be[r], m->dev[r],
                                               (size_t)(be[r + 1] - be[r]) * sizeof(double), m->s_gather[r]));
      }
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_x[r], m->s_gather[r]));
  }
  // ---- 3. partial objective of every device's frames, band by band (other owners first); a finished
  //         band goes to its owner's slot by copy engine while the SMs compute the next one ---------------
  for (int r = 0; r < G; ++r) {
    srb_ctx* c = m->rank[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    for (int q = 0; q < G; ++q) SRB_MULTI_CHECK(m, cudaStreamWaitEvent(c->stream, m->ev_x[q], 0));
    // the slots of this device may still be read by the previous evaluation's sum (same stream: ordered)
    const bool do_reg = reg_active(c) && c->reg_row1 > c->reg_row0;
    double* d_g = c->d_grad;
    if (m->pipelined) {
      for (int i = 0; i < G; ++i) {
        const int o = (r + 1 + i) % G;  // i == G - 1  <=>  o == r: the own band last, it stays local
        if (m->band_unit[o + 1] > m->band_unit[o]) {
          bool reg_done = false;
          srb_status st = fused_eval_units(c, c->d_x, d_g, do_reg, m->band_unit[o], m->band_unit[o + 1], &reg_done);
          if (st != SRB_OK) return multi_status(m, r, st);
        }
        if (o == r) break;
        SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_band[i][r], c->stream));
        if (be[o + 1] > be[o]) {
          cudaStream_t sp = m->s_push[i & 1][r];
          SRB_MULTI_CHECK(m, cudaStreamWaitEvent(sp, m->ev_band[i][r], 0));
          SRB_MULTI_CHECK(m, cudaMemcpyPeerAsync(m->slots[o] + (long long)r * m->band_cap, m->dev[o], d_g + be[o], m->dev[r],
                                                 (size_t)(be[o + 1] - be[o]) * sizeof(double), sp));
        }
      }
      srb_status st = fused_eval_finish(c, c->d_x, d_g, nullptr);
      if (st != SRB_OK) return multi_status(m, r, st);
      c->timing.num_evals += 1;
    } else {
      srb_status st = eval_core(c, c->d_x, d_g, nullptr, true, true, false);
      if (st != SRB_OK) return multi_status(m, r, st);
      SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_band[0][r], c->stream));
      for (int i = 0; i + 1 < G; ++i) {
        const int o = (r + 1 + i) % G;
        if (be[o + 1] <= be[o]) continue;
        cudaStream_t sp = m->s_push[i & 1][r];
        SRB_MULTI_CHECK(m, cudaStreamWaitEvent(sp, m->ev_band[0][r], 0));
        SRB_MULTI_CHECK(m, cudaMemcpyPeerAsync(m->slots[o] + (long long)r * m->band_cap, m->dev[o], d_g + be[o], m->dev[r],
                                               (size_t)(be[o + 1] - be[o]) * sizeof(double), sp));
      }
    }
    SRB_MULTI_CHECK(m, cudaMemcpyAsync(m->h_cost[r], c->d_cost, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    // "all pushes of device r delivered": one event per push stream, joined on the first
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_sum[r], m->s_push[1][r]));
    SRB_MULTI_CHECK(m, cudaStreamWaitEvent(m->s_push[0][r], m->ev_sum[r], 0));
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_pushed[r], m->s_push[0][r]));
  }
  // ---- 4. every owner sums its band in fixed device order and returns it over its own PCIe link -------
  for (int o = 0; o < G; ++o) {
    srb_ctx* c = m->rank[o];
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[o]));
    if (g_host && be[o + 1] > be[o]) {
      for (int q = 0; q < G; ++q)
        if (q != o) SRB_MULTI_CHECK(m, cudaStreamWaitEvent(c->stream, m->ev_pushed[q], 0));
      const long long len = be[o + 1] - be[o];
      const int blocks = (int)std::min<long long>((len + 255) / 256, (long long)c->num_sms * 8);
      k_multi_sum_band<<<blocks, 256, 0, c->stream>>>(c->d_grad + be[o], m->slots[o], m->band_cap, len, G, o);
      c->timing.kernel_launches += 1;
      SRB_MULTI_CHECK(m, cudaGetLastError());
      SRB_MULTI_CHECK(m, cudaMemcpyAsync(g_host + be[o], c->d_grad + be[o], (size_t)len * sizeof(double),
                                         cudaMemcpyDeviceToHost, c->stream));
    } else {
      // nothing to sum here, but the next evaluation must not overwrite slots that are still being filled
      for (int q = 0; q < G; ++q)
        if (q != o) SRB_MULTI_CHECK(m, cudaStreamWaitEvent(c->stream, m->ev_pushed[q], 0));
    }
    SRB_MULTI_CHECK(m, cudaEventRecord(m->ev_t1[o], c->stream));
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
