// srb_multi_solver.cuh -- the device-resident solver (srb_cg.h: ALGLIB's mincg / minlbfgs and the IRLS loop,
// SURVEY.md 8f row N1) on SEVERAL B200s, driven by ONE host thread like everything behind srb_multi_*
// (irls_map_solver.cpp:192-265 is one process and one thread).  Included at the end of srb_api.cu, after
// srb_multi.cuh.
//
// Partition: the row bands of SRB_PARTITION_ROWS.  Every device holds every frame; the (channel, tile row) units
// of the active range are cut into G contiguous bands, and a contiguous band of units is a contiguous range
// [begin, end) of EVERY solver vector (x, g, d, ... are all [c][row][col]).  Device r
//   * keeps its range of every solver vector (plus, for the estimate, the few halo rows either side of it),
//   * runs the streaming kernels of srb_cg_device.cuh on its range only -- 1/G of every vector pass,
//   * evaluates the WHOLE objective on its units (srb_eval_unit_range_dev): its gradient range is final, there is
//     no exchange of the gradient at all.
// What crosses the devices per line-search step: the halo rows of the trial point (2 KH + 1 rows, + R for BTV,
// each way, pulled from the neighbours over NVLink by copy engines: cudaMemcpyPeerAsync behind an event of the
// neighbour's stream) and eight scalars per device, summed on the host in fixed device order (deterministic).
// The estimate enters and leaves over G PCIe links at once, each device moving only its range.
//
// The scalar logic of the backend (the cached sums, the unit direction formed on the fly) is DeviceCgBackend's,
// with every sum taken over all devices; the kernels are the same.
//
// Issue rate.  A line-search step is about a dozen runtime calls per device; issued device after device from the
// calling thread they would cost more than the kernels they start.  Every backend operation is therefore one or
// two fork/join ROUNDS of srb::DeviceWorkers (srb_workers.h): the calling thread issues device 0's work, G - 1 helper
// threads the other devices', each ending with the copy of its eight scalars and the synchronisation of its own
// stream; the calling thread then adds the scalars.  The join between two rounds is what orders "every device has
// recorded the event that marks its rows of x" before "every device waits on its neighbours' events".
// SRB_MULTI_THREADS=0: no helpers, the calling thread walks over the devices (A/B runs).
#pragma once
#include <atomic>
#include <mutex>

#include "srb_row_bands.h"

namespace srb {

struct MultiCgBackend {
  using Vec = int;  // slot: 0 = the estimate (every device's d_x), 1 + i = scratch vector i
  static constexpr int kNone = -1;

  struct Pull {  // elements [begin, end) of the estimate are owned by device `from`
    int from;
    long long begin, end;
  };
  struct Part {
    srb_ctx* c = nullptr;
    int dev = 0;
    long long off = 0, nl = 0;  // this device's range of every vector: [off, off + nl)
    int u0 = 0, u1 = 0;         // = units [u0, u1)
    int nblk = 1;
    double *d_part = nullptr, *d_out = nullptr, *h_out = nullptr;
    std::vector<double*> slot;  // base pointers (whole-vector indexing) of the solver vectors on this device
    std::vector<Pull> pulls;    // halo rows of the estimate, by owner
    cudaEvent_t ev_slice = nullptr;
    DeviceCgWorkspace ws;
  };

  srb_multi* m = nullptr;
  int G = 0;
  long long n = 0;
  Part part[SRB_MAX_PEERS];
  double out[8] = {};  // the scalars of the last operation, combined over the devices
  std::atomic<int> status{SRB_OK};  // first failure of any device (helpers report concurrently)
  std::mutex err_mu;
  DeviceWorkers* workers = nullptr;
  long long evals = 0;

  // (see DeviceCgBackend) the unit direction d = (dk * s1) * s2 is never stored
  Vec unit_d = kNone, unit_src = kNone, unit_g0 = kNone;
  double unit_s1 = 1.0, unit_s2 = 1.0;
  Vec dir_dk = kNone, dir_g = kNone;
  double dir_sumsq = 0.0, dir_gdk = 0.0;
  Vec trial_g = kNone;
  double trial_dy = 0.0, trial_gg = 0.0, trial_gy = 0.0;

  long long size() const { return n; }
  bool ok() const { return status.load(std::memory_order_relaxed) == SRB_OK; }
  srb_status result() const { return (srb_status)status.load(); }
  void fail_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return;
    std::lock_guard<std::mutex> lk(err_mu);
    if (ok()) status = m->fail(SRB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  void fail_rank(int r, srb_status st) {
    if (st == SRB_OK) return;
    std::lock_guard<std::mutex> lk(err_mu);
    if (ok()) status = multi_status(m, r, st);
  }
  // One round: f(r, part) for every device that owns a non-empty range, with that device current -- device 0 on
  // the calling thread, the others on the helper threads; returns when all have returned.
  template <class F>
  void each(F f) {
    auto job = [&](int r) {
      Part& p = part[r];
      if (p.nl <= 0) return;
      fail_cuda(cudaSetDevice(p.dev), "cudaSetDevice");
      f(r, p);
    };
    if (workers && G > 1) {
      workers->run(G, job);
    } else {
      for (int r = 0; r < G; ++r) job(r);
    }
  }
  static double* at(const Part& p, Vec v) { return p.slot[v] + p.off; }
  void launched(Part& p, int k = 1) { p.c->timing.kernel_launches += k; }
  void finish(Part& p, int nsums, int max_mask = 0, int offset = 0) {
    k_cg_finish<<<1, CG_NT, 0, p.c->stream>>>(p.d_part, p.nblk, nsums, max_mask, p.d_out + offset);
    launched(p, 2);
  }
  // end of a device's share of a round: its eight scalars -> its pinned mirror, its stream drained
  void pull(Part& p) {
    fail_cuda(cudaMemcpyAsync(p.h_out, p.d_out, 8 * sizeof(double), cudaMemcpyDeviceToHost, p.c->stream), "solver fetch");
    fail_cuda(cudaStreamSynchronize(p.c->stream), "solver synchronize");
  }
  // after the round: out[0..8) = sums (maxima where max_mask says so) over the devices, fixed order
  void combine(int max_mask = 0) {
    for (int k = 0; k < 8; ++k) {
      double v = 0.0;
      for (int r = 0; r < G; ++r) {
        if (part[r].nl <= 0) continue;
        const double t = part[r].h_out[k];
        v = ((max_mask >> k) & 1) ? std::fmax(v, t) : v + t;
      }
      out[k] = ok() ? v : NAN;
    }
  }
  void forget(Vec v) {
    if (v == unit_d || v == unit_src) unit_d = unit_src = kNone;
    if (v == dir_dk || v == dir_g) dir_dk = dir_g = kNone;
    if (v == trial_g) trial_g = kNone;
    if (v == unit_g0) trial_g = unit_g0 = kNone;
  }

  // Halo exchange, two halves in two rounds: mark() -- this device's rows of x are complete on its stream (the
  // last call of a round) -- and halo() -- pull the halo rows from their owners behind THEIR marks (the first call
  // of the next round; the join in between guarantees that every mark has been recorded).
  void mark(Part& p) {
    if (G > 1) fail_cuda(cudaEventRecord(p.ev_slice, p.c->stream), "cudaEventRecord");
  }
  void halo(Part& p, Vec x) {
    {
      for (const Pull& h : p.pulls) {
        const Part& q = part[h.from];
        fail_cuda(cudaStreamWaitEvent(p.c->stream, q.ev_slice, 0), "cudaStreamWaitEvent");
        const size_t bytes = (size_t)(h.end - h.begin) * sizeof(double);
        if (q.dev == p.dev)  // SRB_MULTI_SHARE_DEVICES: two contexts on one GPU
          fail_cuda(cudaMemcpyAsync(p.slot[x] + h.begin, q.slot[x] + h.begin, bytes, cudaMemcpyDeviceToDevice, p.c->stream), "halo copy");
        else
          fail_cuda(cudaMemcpyPeerAsync(p.slot[x] + h.begin, p.dev, q.slot[x] + h.begin, q.dev, bytes, p.c->stream), "halo copy");
      }
    }
  }
  // objective + gradient of this device's units at x -> its range of g, its share of the cost -> d_out[5]
  void evaluate(int r, Part& p, Vec x, Vec g) {
    halo(p, x);
    if (ok()) fail_rank(r, srb_eval_unit_range_dev(p.c, p.slot[x], p.slot[g], p.u0, p.u1, p.d_out + 5));
  }

  void eval(Vec x, Vec g, double* f) {
    forget(g);
    if (G > 1) each([&](int, Part& p) { mark(p); });
    each([&](int r, Part& p) {
      evaluate(r, p, x, g);
      pull(p);
    });
    ++evals;
    combine();
    *f = out[5];
  }
  void copy(Vec dst, Vec src) {
    forget(dst);
    each([&](int, Part& p) {
      fail_cuda(cudaMemcpyAsync(at(p, dst), at(p, src), (size_t)p.nl * sizeof(double), cudaMemcpyDeviceToDevice, p.c->stream), "solver copy");
    });
  }
  void neg_copy(Vec dst, Vec src) {
    forget(dst);
    each([&](int, Part& p) {
      k_cg_neg_copy<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, dst), at(p, src), p.nl);
      launched(p);
    });
  }
  void zero(Vec v) {
    forget(v);
    each([&](int, Part& p) { fail_cuda(cudaMemsetAsync(at(p, v), 0, (size_t)p.nl * sizeof(double), p.c->stream), "solver zero"); });
  }
  double dot(Vec a, Vec b) {
    each([&](int, Part& p) {
      if (b == unit_d && unit_src != kNone)
        k_cg_dot_scaled<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, a), at(p, unit_src), unit_s1, unit_s2, p.nl, p.d_part);
      else if (a == unit_d && unit_src != kNone)
        k_cg_dot_scaled<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, b), at(p, unit_src), unit_s1, unit_s2, p.nl, p.d_part);
      else
        k_cg_reduce<0><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, a), at(p, b), nullptr, p.nl, p.d_part);
      finish(p, 1);
      pull(p);
    });
    combine();
    return out[0];
  }
  double sum_sq(Vec a) {
    each([&](int, Part& p) {
      k_cg_reduce<1><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, a), nullptr, nullptr, p.nl, p.d_part);
      finish(p, 1);
      pull(p);
    });
    combine();
    return out[0];
  }
  double max_abs(Vec a) {
    each([&](int, Part& p) {
      k_cg_max_abs<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, a), p.nl, p.d_part);
      finish(p, 1, 1);
      pull(p);
    });
    combine(1);
    return out[0];
  }
  void normalize_to(Vec d, Vec dk, double mx, Vec g0, double* stp, double* slope, double* dd) {
    forget(d);
    if (mx == 0.0) {
      copy(d, dk);
      *slope = 0.0;
      *dd = 0.0;
      return;
    }
    const double s1 = 1 / mx;
    double sumsq_scaled = 0.0, gdk = 0.0;
    bool have = dir_dk == dk && dir_g == g0;
    if (have) {
      sumsq_scaled = dir_sumsq * s1 * s1;
      gdk = dir_gdk;
      have = std::isfinite(dir_sumsq) && dir_sumsq > 1e-280 && std::isfinite(sumsq_scaled) && sumsq_scaled > 0.0;
    }
    if (!have) {
      each([&](int, Part& p) {
        k_cg_scaled_sumsq<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, dk), s1, p.nl, p.d_part);
        finish(p, 1);
        k_cg_reduce<0><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, g0), at(p, dk), nullptr, p.nl, p.d_part);
        finish(p, 1, 0, 1);
        pull(p);
      });
      combine();
      sumsq_scaled = out[0];
      gdk = out[1];
    }
    const double s2 = 1 / std::sqrt(sumsq_scaled);
    unit_d = d; unit_src = dk; unit_g0 = g0; unit_s1 = s1; unit_s2 = s2;
    trial_g = kNone;
    *stp = *stp / s1;
    *stp = *stp / s2;
    *slope = gdk * s1 * s2;
    *dd = sumsq_scaled * s2 * s2;
  }
  void trial(Vec x, Vec x0, double stp, Vec d, Vec g, double* f, double* dg, double* moved) {
    forget(x);
    const bool unit = d == unit_d && unit_src != kNone;
    each([&](int, Part& p) {
      if (unit) k_cg_step_scaled<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, x), at(p, x0), stp, at(p, unit_src), unit_s1, unit_s2, p.nl, p.d_part);
      else k_cg_step<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, x), at(p, x0), stp, at(p, d), p.nl, p.d_part);
      finish(p, 1, 0, 4);  // moved -> d_out[4]
      mark(p);
    });
    if (g == trial_g) trial_g = kNone;
    const bool sums = unit && unit_g0 != kNone && g != unit_g0;
    each([&](int r, Part& p) {
      evaluate(r, p, x, g);
      if (sums) {
        k_cg_trial_sums<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, g), at(p, unit_g0), at(p, unit_src), unit_s1, unit_s2, p.nl, p.d_part);
        finish(p, 4);
      } else if (unit) {
        k_cg_dot_scaled<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, g), at(p, unit_src), unit_s1, unit_s2, p.nl, p.d_part);
        finish(p, 1);
      } else {
        k_cg_reduce<0><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, g), at(p, d), nullptr, p.nl, p.d_part);
        finish(p, 1);
      }
      pull(p);
    });
    ++evals;
    combine();
    *f = out[5];
    *dg = out[0];
    *moved = out[4];
    if (sums) {
      trial_g = g;
      trial_dy = out[1]; trial_gg = out[2]; trial_gy = out[3];
    }
  }
  void beta_terms(Vec gn, Vec go, Vec dk, double* dy, double* gg, double* gy) {
    if (gn == trial_g && go == unit_g0 && dk == unit_src) {
      *dy = trial_dy; *gg = trial_gg; *gy = trial_gy;
      return;
    }
    each([&](int, Part& p) {
      k_cg_reduce<3><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, gn), at(p, go), at(p, dk), p.nl, p.d_part);
      finish(p, 3);
      pull(p);
    });
    combine();
    *dy = out[0]; *gg = out[1]; *gy = out[2];
  }
  void direction(Vec dk, Vec g, double beta, double* gg, double* mx) {
    forget(dk);
    each([&](int, Part& p) {
      k_cg_direction<<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, dk), at(p, g), beta, p.nl, p.d_part);
      finish(p, 4, 2);
      pull(p);
    });
    combine(2);
    *gg = out[0]; *mx = out[1];
    dir_dk = dk; dir_g = g; dir_sumsq = out[2]; dir_gdk = out[3];
  }
  template <int MODE>
  void update(Vec dst, double a, Vec src) {
    forget(dst);
    each([&](int, Part& p) {
      k_cg_update<MODE><<<p.nblk, CG_NT, 0, p.c->stream>>>(at(p, dst), a, src == kNone ? nullptr : at(p, src), p.nl);
      launched(p);
    });
  }
  void add(Vec dst, Vec src) { update<0>(dst, 0.0, src); }
  void add_scaled(Vec dst, double a, Vec src) { update<1>(dst, a, src); }
  void sub_scaled(Vec dst, double a, Vec src) { update<2>(dst, a, src); }
  void scale(Vec v, double a) { update<3>(v, a, kNone); }
  // w = 1 / max(1e-5, reg(x)) (irls_map_solver.cpp:128-143): regularizer values are local, so every device
  // re-weights from its own rows of x plus the halo -- the weights its units read are exact, the rest unused
  void reweight(Vec x) {
    if (G > 1) each([&](int, Part& p) { mark(p); });
    each([&](int r, Part& p) {
      halo(p, x);
      if (ok()) fail_rank(r, reweight_dev(p.c, p.slot[x]));
    });
  }
};

// Can the current configuration be solved on row bands?  (fused tile kernel on every device, a regularizer it covers
// that does not couple the channels; a frame-sharded context holds only some frames per device.)
inline bool multi_solver_rows_ok(srb_multi* m) {
  if (m->G > 1 && m->partition != SRB_PARTITION_ROWS) return false;
  for (int r = 0; r < m->G; ++r)
    if (!unit_ranges_ok(m->rank[r])) return false;
  return host_slices_ok(m->rank[0]);
}

// The helper threads spin between rounds while a solve is running and sleep otherwise.
struct MultiSolveScope {
  DeviceWorkers* w = nullptr;
  MultiSolveScope(srb_multi* m, MultiCgBackend* be) {
    w = multi_workers(m);
    be->workers = w;
    if (w) w->set_hot(true);
  }
  ~MultiSolveScope() {
    if (w) w->set_hot(false);
  }
};

// Devices, ranges, halos, workspaces of one solve; uploads every device's range (+ halo) of x.
inline srb_status multi_solver_begin(srb_multi* m, MultiCgBackend* be, const double* x_host, int num_vectors) {
  const int G = m->G;
  srb_ctx* c0 = m->rank[0];
  be->m = m;
  be->G = G;
  be->n = (long long)c0->n_active();
  // the partition is host arithmetic (srb_row_bands.h; tests/test_row_bands.py); the halo is the same on every device
  const std::vector<RowBand> bands = plan_row_bands(G, c0->Ca(), c0->g.H, c0->g.W, tile_height(c0), stencil_halo_rows(c0));
  for (int r = 0; r < G; ++r) {
    MultiCgBackend::Part& p = be->part[r];
    p.c = m->rank[r];
    p.dev = m->dev[r];
    p.u0 = bands[r].u0;
    p.u1 = bands[r].u1;
    p.off = bands[r].begin;
    p.nl = bands[r].end - bands[r].begin;
    p.ev_slice = m->ev_x[0][r];
  }
  for (int r = 0; r < G; ++r) {
    MultiCgBackend::Part& p = be->part[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(p.dev));
    DeviceCgBackend single;
    srb_status st = p.ws.init(p.c, &single, num_vectors);
    if (st != SRB_OK) return multi_status(m, r, st);
    p.d_part = single.d_part;
    p.d_out = single.d_out;
    p.h_out = single.h_out;
    p.nblk = (int)std::max<long long>(1, std::min<long long>((p.nl + CG_NT - 1) / CG_NT, single.nblk));
    p.slot.assign(1, p.c->d_x);
    p.slot.insert(p.slot.end(), p.ws.scratch.begin(), p.ws.scratch.end());
    SRB_MULTI_CHECK(m, cudaMemsetAsync(p.d_out, 0, 8 * sizeof(double), p.c->stream));
    for (int k = 0; k < 8; ++k) p.h_out[k] = 0.0;
    p.pulls.clear();
    if (p.nl <= 0) continue;
    const long long lo = bands[r].halo_begin, hi = bands[r].halo_end;
    for (const RowBandPull& h : bands[r].pulls) p.pulls.push_back({h.from, h.begin, h.end});
    // direct NVLink copies where the devices can reach each other (a staged copy otherwise: still correct)
    for (const MultiCgBackend::Pull& h : p.pulls) {
      const int qd = be->part[h.from].dev;
      if (qd == p.dev) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, p.dev, qd) == cudaSuccess && can) (void)cudaDeviceEnablePeerAccess(qd, 0);
      (void)cudaGetLastError();
    }
    // the two vectors that are ever evaluated (the estimate and the trial point of the line search, which swap
    // roles) hold defined values outside this device's rows too: the tile kernel's TMA boxes reach past the halo
    // (values it never uses arithmetically, but they must not be signalling garbage from an earlier allocation)
    SRB_MULTI_CHECK(m, cudaMemsetAsync(p.slot[0], 0, (size_t)be->n * sizeof(double), p.c->stream));
    SRB_MULTI_CHECK(m, cudaMemsetAsync(p.slot[1], 0, (size_t)be->n * sizeof(double), p.c->stream));
    SRB_MULTI_CHECK(m, cudaMemcpyAsync(p.c->d_x + lo, x_host + lo, (size_t)(hi - lo) * sizeof(double),
                                       cudaMemcpyHostToDevice, p.c->stream));
    p.c->x_resident = false;  // only this device's rows of x are here
  }
  return SRB_OK;
}

// Every device returns its range of the solution; all streams drain.
inline srb_status multi_solver_end(srb_multi* m, MultiCgBackend* be, double* x_host) {
  for (int r = 0; r < be->G; ++r) {
    MultiCgBackend::Part& p = be->part[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(p.dev));
    if (p.nl > 0 && be->ok())
      SRB_MULTI_CHECK(m, cudaMemcpyAsync(x_host + p.off, p.c->d_x + p.off, (size_t)p.nl * sizeof(double),
                                         cudaMemcpyDeviceToHost, p.c->stream));
  }
  for (int r = 0; r < be->G; ++r) {
    MultiCgBackend::Part& p = be->part[r];
    SRB_MULTI_CHECK(m, cudaSetDevice(p.dev));
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(p.c->stream));
    SRB_MULTI_CHECK(m, cudaGetLastError());
  }
  return be->result();
}

inline std::vector<int> multi_solver_scratch(int num_vectors) {
  std::vector<int> s(num_vectors);
  for (int i = 0; i < num_vectors; ++i) s[i] = 1 + i;
  return s;
}

inline srb_status multi_solver_check(srb_multi* m, const double* x_host, const srb_cg_options* options) {
  if (!x_host) return m->fail(SRB_ERR_INVALID, "null estimate");
  if (!cg_options_valid(options)) return m->fail(SRB_ERR_INVALID, "invalid solver thresholds");
  for (int r = 0; r < m->G; ++r)
    if (!m->rank[r]->have_obs) return m->fail(SRB_ERR_STATE, "srb_multi_set_observations has not been called");
  if (m->G > 1 && m->partition != SRB_PARTITION_ROWS)
    return m->fail(SRB_ERR_STATE, "the multi-GPU solver needs the row-band partition (srb_multi_create_partitioned with "
                                  "SRB_PARTITION_ROWS): a frame shard cannot evaluate the whole objective on its rows");
  return SRB_OK;
}

}  // namespace srb

extern "C" {

// RunCGSolverAnalyticalDiff (alglib_objective.cpp:47-75) on all devices.
srb_status srb_multi_cg_minimize(srb_multi* m, double* x_host, const srb_cg_options* options, srb_cg_report* report) {
  using namespace srb;
  if (!m) return SRB_ERR_INVALID;
  DeviceRestore restore_device;
  srb_status st = multi_solver_check(m, x_host, options);
  if (st != SRB_OK) return st;
  if (!multi_solver_rows_ok(m))  // e.g. 3-D TV: every device holds the whole model, device 0 solves alone
    return multi_status(m, 0, srb_cg_minimize(m->rank[0], x_host, options, report));
  MultiCgBackend be;
  MultiSolveScope scope(m, &be);
  if ((st = multi_solver_begin(m, &be, x_host, kCgScratchVectors)) != SRB_OK) return st;
  std::vector<int> scratch = multi_solver_scratch(kCgScratchVectors);
  const CgReport rep = cg_minimize(be, 0, scratch.data(), cg_options_from(options));
  if ((st = multi_solver_end(m, &be, x_host)) != SRB_OK) return st;
  cg_report_to(rep, report);
  m->timing.num_evals += be.evals;
  return SRB_OK;
}

// RunLBFGSSolverAnalyticalDiff (alglib_objective.cpp:111-140) on all devices.
srb_status srb_multi_lbfgs_minimize(srb_multi* m, double* x_host, const srb_cg_options* options, srb_cg_report* report) {
  using namespace srb;
  if (!m) return SRB_ERR_INVALID;
  DeviceRestore restore_device;
  srb_status st = multi_solver_check(m, x_host, options);
  if (st != SRB_OK) return st;
  if (!options || options->num_lbfgs_hessian_corrections < 1)
    return m->fail(SRB_ERR_INVALID, "invalid solver options (L-BFGS needs 1..64 correction pairs)");
  if ((long long)options->num_lbfgs_hessian_corrections > (long long)m->rank[0]->n_active())
    return m->fail(SRB_ERR_INVALID, "more correction pairs than parameters");
  if (!multi_solver_rows_ok(m)) return multi_status(m, 0, srb_lbfgs_minimize(m->rank[0], x_host, options, report));
  const int mm = options->num_lbfgs_hessian_corrections, nvec = lbfgs_scratch_vectors(mm);
  MultiCgBackend be;
  MultiSolveScope scope(m, &be);
  if ((st = multi_solver_begin(m, &be, x_host, nvec)) != SRB_OK) return st;
  std::vector<int> scratch = multi_solver_scratch(nvec);
  const CgReport rep = lbfgs_minimize(be, 0, scratch.data(), mm, cg_options_from(options));
  if ((st = multi_solver_end(m, &be, x_host)) != SRB_OK) return st;
  cg_report_to(rep, report);
  m->timing.num_evals += be.evals;
  return SRB_OK;
}

// IRLSMapSolver::RunIRLSLoop (irls_map_solver.cpp:45-157) on all devices.
srb_status srb_multi_solve_irls(srb_multi* m, double* x_host, const srb_cg_options* options, int max_num_irls_iterations,
                                double irls_cost_difference_threshold, srb_irls_report* report) {
  using namespace srb;
  if (!m) return SRB_ERR_INVALID;
  DeviceRestore restore_device;
  srb_status st = multi_solver_check(m, x_host, options);
  if (st != SRB_OK) return st;
  if (max_num_irls_iterations < 0 || !(irls_cost_difference_threshold >= 0)) return m->fail(SRB_ERR_INVALID, "invalid solver options");
  const bool has_reg = reg_active(m->rank[0]);
  if (max_num_irls_iterations == 0 && irls_cost_difference_threshold == 0.0 && has_reg)
    return m->fail(SRB_ERR_INVALID, "unlimited IRLS iterations with a zero cost-difference threshold never terminate "
                                    "(the reference's defaults are 20 and 1e-5, irls_map_solver.h:27,35)");
  if (!multi_solver_rows_ok(m))
    return multi_status(m, 0, srb_solve_irls(m->rank[0], x_host, options, max_num_irls_iterations,
                                             irls_cost_difference_threshold, report));
  for (int r = 0; r < m->G; ++r) {  // irls_map_solver.cpp:66-74: all weights 1
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    if ((st = reset_weights(m->rank[r])) != SRB_OK) return multi_status(m, r, st);
  }
  const int mm = options ? options->num_lbfgs_hessian_corrections : 0;
  const int nvec = mm > 0 ? lbfgs_scratch_vectors(mm) : kCgScratchVectors;
  MultiCgBackend be;
  MultiSolveScope scope(m, &be);
  if ((st = multi_solver_begin(m, &be, x_host, nvec)) != SRB_OK) return st;
  std::vector<int> scratch = multi_solver_scratch(nvec);
  const IrlsReport rep = irls_solve(be, 0, scratch.data(), cg_options_from(options), max_num_irls_iterations,
                                    irls_cost_difference_threshold, has_reg, mm);
  if ((st = multi_solver_end(m, &be, x_host)) != SRB_OK) return st;
  // the weights are internal to the loop (a local vector of RunIRLSLoop, irls_map_solver.cpp:66-74); every device
  // has re-weighted only the rows it owns, so they are not left behind: back to 1 everywhere
  for (int r = 0; r < m->G; ++r) {
    SRB_MULTI_CHECK(m, cudaSetDevice(m->dev[r]));
    if ((st = reset_weights(m->rank[r])) != SRB_OK) return multi_status(m, r, st);
    SRB_MULTI_CHECK(m, cudaStreamSynchronize(m->rank[r]->stream));
  }
  if (report) {
    report->num_irls_iterations = rep.irls_iterations;
    report->num_solver_iterations = rep.solver_iterations;
    report->num_evaluations = rep.nfev;
    report->last_termination_type = rep.last_termination;
    report->final_cost = rep.f;
  }
  m->timing.num_evals += be.evals;
  return SRB_OK;
}

}  // extern "C"
