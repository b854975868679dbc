// cg_host_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Instantiates the product's CG restatement (super-resolution_b200/csrc/srb_cg.h) over plain host
// arrays with ALGLIB's summation orders (ap.cpp:4667-4692 ae_v_dotproduct: groups of four; the
// hand-written loops of mincgiteration: left to right), so that tests/test_cg_restatement.py can
// compare it bit for bit with the reference's own ALGLIB (oracle/_ref, ref_mincg).  The product
// instantiates the same template over device vectors (srb_cg_device.cuh).
#include <cmath>
#include <cstring>
#include <vector>

#include "../super-resolution_b200/csrc/srb_cg.h"

extern "C" {
typedef void (*srbcg_fg_cb)(long long n, const double* x, double* f, double* g, void* user);
typedef void (*srbcg_reweight_cb)(long long n, const double* x, void* user);
}

namespace {

struct HostBackend {
  using Vec = double*;
  long long n;
  srbcg_fg_cb cb;
  void* user;
  srbcg_reweight_cb rw = nullptr;
  long long evals = 0;

  long long size() const { return n; }
  void eval(Vec x, Vec g, double* f) { cb(n, x, f, g, user); ++evals; }
  void copy(Vec d, Vec s) { std::memcpy(d, s, (size_t)n * sizeof(double)); }
  void neg_copy(Vec d, Vec s) { for (long long i = 0; i < n; ++i) d[i] = -s[i]; }
  void scale_to(Vec d, Vec s, double a) { for (long long i = 0; i < n; ++i) d[i] = s[i] * a; }
  void scale(Vec v, double a) { for (long long i = 0; i < n; ++i) v[i] *= a; }
  void step_to(Vec d, Vec b, double a, Vec dir) { for (long long i = 0; i < n; ++i) d[i] = b[i] + a * dir[i]; }
  void zero(Vec v) { for (long long i = 0; i < n; ++i) v[i] = 0.0; }
  void add(Vec d, Vec s) { for (long long i = 0; i < n; ++i) d[i] += s[i]; }               // ae_v_add
  void add_scaled(Vec d, double a, Vec s) { for (long long i = 0; i < n; ++i) d[i] += a * s[i]; }  // ae_v_addd
  void sub_scaled(Vec d, double a, Vec s) { for (long long i = 0; i < n; ++i) d[i] -= a * s[i]; }  // ae_v_subd
  // ae_v_dotproduct: groups of four, then the remainder
  template <class FA, class FB>
  double dot4(FA a, FB b) {
    double r = 0;
    const long long n4 = n / 4;
    long long i = 0;
    for (long long k = 0; k < n4; ++k, i += 4)
      r += a(i) * b(i) + a(i + 1) * b(i + 1) + a(i + 2) * b(i + 2) + a(i + 3) * b(i + 3);
    for (; i < n; ++i) r += a(i) * b(i);
    return r;
  }
  double dot(Vec a, Vec b) {
    return dot4([a](long long i) { return a[i]; }, [b](long long i) { return b[i]; });
  }
  double sum_sq(Vec a) {
    double r = 0;
    for (long long i = 0; i < n; ++i) r = r + a[i] * a[i];
    return r;
  }
  double sum_sq_diff(Vec a, Vec b) {
    double r = 0;
    for (long long i = 0; i < n; ++i) r = r + (a[i] - b[i]) * (a[i] - b[i]);
    return r;
  }
  double max_abs(Vec a) {
    double m = 0;
    for (long long i = 0; i < n; ++i) m = std::fabs(a[i]) > m ? std::fabs(a[i]) : m;
    return m;
  }
  void beta_terms(Vec gn, Vec go, Vec dk, double* dy, double* gg, double* gy) {
    auto y = [gn, go](long long i) { return -go[i] + gn[i]; };  // yk = -g_old, then yk += g_new
    *dy = dot4(y, [dk](long long i) { return dk[i]; });
    *gg = dot(gn, gn);
    *gy = dot4([gn](long long i) { return gn[i]; }, y);
  }
  void reweight(Vec x) { rw(n, x, user); }
  void direction(Vec dk, Vec g, double beta, double* gg, double* mx) {
    for (long long i = 0; i < n; ++i) dk[i] = -g[i] + beta * dk[i];
    *gg = sum_sq(g);
    *mx = max_abs(dk);
  }
  // linminnormalized (alglibinternal.cpp:12165-12195) + the slope mcsrch computes first + the
  // squared length mincgiteration computes after the line search
  void normalize_to(Vec d, Vec dk, double mx, Vec g0, double* stp, double* slope, double* dd) {
    if (mx == 0.0) {
      copy(d, dk);
    } else {
      double s = 1 / mx;
      scale_to(d, dk, s);
      *stp = *stp / s;
      s = 1 / std::sqrt(dot(d, d));
      scale(d, s);
      *stp = *stp / s;
    }
    *slope = dot(g0, d);
    *dd = sum_sq(d);
  }
  void trial(Vec x, Vec x0, double stp, Vec d, Vec g, double* f, double* dg, double* moved) {
    step_to(x, x0, stp, d);
    eval(x, g, f);
    *dg = dot(g, d);
    *moved = sum_sq_diff(x0, x);
  }
};

}  // namespace

extern "C" {
// report: [iterations, nfev, termination type, final f, restarts, objective evaluations]
int srbcg_host_minimize(long long n, double* x_inout, double epsg, double epsf, double epsx, int maxits,
                        srbcg_fg_cb cb, void* user, double* report) {
  HostBackend be{n, cb, user};
  std::vector<double> store((size_t)srb::kCgScratchVectors * n);
  double* scratch[srb::kCgScratchVectors];
  for (int i = 0; i < srb::kCgScratchVectors; ++i) scratch[i] = store.data() + (size_t)i * n;
  srb::CgOptions opt;
  opt.epsg = epsg; opt.epsf = epsf; opt.epsx = epsx; opt.maxits = maxits;
  const srb::CgReport rep = srb::cg_minimize(be, x_inout, scratch, opt);
  report[0] = rep.iterations;
  report[1] = rep.nfev;
  report[2] = rep.termination;
  report[3] = rep.f;
  report[4] = rep.restarts;
  report[5] = (double)be.evals;
  return 0;
}

// IRLSMapSolver::RunIRLSLoop over host arrays; `rw` installs the new IRLS weights in the objective.
// report: [IRLS iterations, CG iterations (summed), nfev (summed), last termination type, final f]
int srbcg_host_irls(long long n, double* x_inout, double epsg, double epsf, double epsx, int maxits,
                    int max_irls_iterations, double cost_difference_threshold, int has_regularizer,
                    int lbfgs_corrections, srbcg_fg_cb cb, srbcg_reweight_cb rw, void* user, double* report) {
  HostBackend be{n, cb, user};
  be.rw = rw;
  const int nv = lbfgs_corrections > 0 ? srb::lbfgs_scratch_vectors(lbfgs_corrections) : srb::kCgScratchVectors;
  std::vector<double> store((size_t)nv * n);
  std::vector<double*> scratch_v(nv);
  for (int i = 0; i < nv; ++i) scratch_v[i] = store.data() + (size_t)i * n;
  double** scratch = scratch_v.data();
  srb::CgOptions opt;
  opt.epsg = epsg; opt.epsf = epsf; opt.epsx = epsx; opt.maxits = maxits;
  const srb::IrlsReport rep = srb::irls_solve(be, x_inout, scratch, opt, max_irls_iterations,
                                              cost_difference_threshold, has_regularizer != 0, lbfgs_corrections);
  report[0] = rep.irls_iterations;
  report[1] = rep.solver_iterations;
  report[2] = rep.nfev;
  report[3] = rep.last_termination;
  report[4] = rep.f;
  return 0;
}

// ALGLIB minlbfgs restated (srb_cg.h: lbfgs_minimize) over host arrays.
// report: [iterations, nfev, termination type, final f, restarts, objective evaluations]
int srbcg_host_lbfgs(long long n, double* x_inout, int m, double epsg, double epsf, double epsx, int maxits,
                     srbcg_fg_cb cb, void* user, double* report) {
  HostBackend be{n, cb, user};
  const int nv = srb::lbfgs_scratch_vectors(m);
  std::vector<double> store((size_t)nv * n);
  std::vector<double*> scratch(nv);
  for (int i = 0; i < nv; ++i) scratch[i] = store.data() + (size_t)i * n;
  srb::CgOptions opt;
  opt.epsg = epsg; opt.epsf = epsf; opt.epsx = epsx; opt.maxits = maxits;
  const srb::CgReport rep = srb::lbfgs_minimize(be, x_inout, scratch.data(), m, opt);
  report[0] = rep.iterations;
  report[1] = rep.nfev;
  report[2] = rep.termination;
  report[3] = rep.f;
  report[4] = rep.restarts;
  report[5] = (double)be.evals;
  return 0;
}
}
