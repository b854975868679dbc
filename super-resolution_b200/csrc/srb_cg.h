// srb_cg.h -- nonlinear conjugate gradients with the solver vectors resident wherever the backend
// keeps them.
//
// The reference minimises the MAP objective with ALGLIB's `mincg` (alglib_objective.cpp:47-75:
// mincgcreate / mincgsetcond / mincgoptimize / mincgresults; ALGLIB 3.x vendored under libs/alglib,
// optimization.cpp:17137-17850 `mincgiteration`, alglibinternal.cpp:12165-12195 `linminnormalized`,
// :12313-12640 `mcsrch`, :12972-13232 `mcstep`).  Every evaluation there crosses PCIe twice (x in,
// g out).  This header restates that algorithm -- the same direction update, line search, safeguards,
// restart rules and stopping tests, in the same order -- over an abstract vector backend, so that
// x, g, d, ... never leave the device and only scalars reach the host.  With the host backend of
// tests/ (sequential sums in ALGLIB's order) the iterates are bit-identical to ALGLIB's; the CUDA
// backend (srb_cg_device.cuh) differs only in the summation order of its reductions.
//
// ALGLIB's defaults as the reference uses them: no preconditioner, unit scales, cgtype = -1 (the
// min(DY, HS) hybrid clipped at 0), no step bound, analytic gradient, first trial step 1.
//
// Backend concept (all vectors have the problem's length n; "sum" orders are ALGLIB's on the host
// backend of the tests, free on the device).  The operations are the fused groups the CUDA backend
// runs as one or two streaming kernels each:
//   using Vec = ...;                       cheap handle
//   long long size();
//   void   eval(Vec x, Vec g, double* f);                      objective + gradient at x
//   void   neg_copy(Vec dst, Vec src);                         dst = -src
//   void   copy(Vec dst, Vec src);
//   void   zero(Vec v);
//   double dot(Vec a, Vec b);                                  ae_v_dotproduct (ap.cpp:4667-4692)
//   double sum_sq(Vec a);                                      sum a_i^2 (plain loop in ALGLIB)
//   double max_abs(Vec a);
//   void   normalize_to(Vec d, Vec dk, double mx, Vec g0, double* stp, double* slope, double* dd);
//            linminnormalized with mx = max |dk_i|: d = (dk * s1) * s2, s1 = 1 / mx,
//            s2 = 1 / sqrt(<dk * s1, dk * s1>), *stp = *stp / s1 / s2  (d = dk, *stp unchanged when
//            mx == 0);  *slope = <g0, d>;  *dd = sum d_i^2
//   void   trial(Vec x, Vec x0, double stp, Vec d, Vec g, double* f, double* dg, double* moved);
//            x = x0 + stp * d;  f, g = objective at x;  dg = <g, d>;  moved = sum (x0_i - x_i)^2
//   void   beta_terms(Vec g_new, Vec g_old, Vec dk, double* dy, double* gg, double* gy);
//            with y = g_new - g_old:  dy = <y, dk>, gg = <g_new, g_new>, gy = <g_new, y>
//   void   direction(Vec dk, Vec g, double beta, double* gg, double* mx);
//            dk = -g + beta * dk (in place);  gg = sum g_i^2;  mx = max |dk_i| of the new direction
//   void   reweight(Vec x);                                    irls_solve only: new IRLS weights from x
//   lbfgs_minimize only:
//   void   add(Vec dst, Vec src);                              dst += src
//   void   add_scaled(Vec dst, double a, Vec src);             dst += a * src   (ae_v_addd)
//   void   sub_scaled(Vec dst, double a, Vec src);             dst -= a * src   (ae_v_subd)
//   void   scale(Vec v, double a);                             v *= a
#pragma once
#include <cmath>

namespace srb {

struct CgOptions {  // mincgsetcond (alglib_objective.cpp:57-62)
  double epsg = 0.0;   // stop when |g| <= epsg
  double epsf = 0.0;   // stop when f_k - f_{k+1} <= epsf * max(|f_k|, |f_{k+1}|, 1)
  double epsx = 0.0;   // stop when the step length <= epsx
  int maxits = 0;      // 0: unlimited
};

struct CgReport {      // mincgreport + the final cost RunCGSolverAnalyticalDiff returns
  int iterations = 0;
  int nfev = 0;
  int termination = 0;  // 1 epsf, 2 epsx, 4 epsg, 5 maxits, 7 repeated restarts, -8 inf / nan
  int restarts = 0;     // line searches that did not end on the Wolfe conditions
  double f = 0.0;       // objective at the last evaluated point (mincgstate.f)
};

namespace cg_detail {

constexpr double kFtol = 1e-3;                     // sufficient decrease
constexpr double kGtol = 0.3;                      // curvature condition (mincg's)
constexpr double kGtolLbfgs = 0.4;                 // minlbfgs's (optimization.cpp:8941)
constexpr double kXtol = 100 * 5e-16;               // ALGLIB's ae_machineepsilon is 5e-16 (ap.h:855)
constexpr int kMaxFev = 20;
constexpr double kStpMin = 1e-50;
constexpr double kStpMaxDefault = 1e50;
constexpr int kRestartCountdown = 10;

inline double max2(double a, double b) { return a > b ? a : b; }
inline double min2(double a, double b) { return a > b ? b : a; }

// Interval of uncertainty of the More-Thuente search: best step so far (x), other end (y).
struct Bracket {
  double stx, fx, dx;
  double sty, fy, dy;
  bool bracketed;
};

// One safeguarded trial-step update from the trial (stp, fp, dp); returns which of the four cases
// applied (0: inputs inconsistent, nothing changed).
inline int trial_step(Bracket& b, double& stp, double fp, double dp, double stmin, double stmax) {
  if ((b.bracketed && (stp <= min2(b.stx, b.sty) || stp >= max2(b.stx, b.sty))) ||
      b.dx * (stp - b.stx) >= 0.0 || stmax < stmin)
    return 0;
  const double sgnd = dp * (b.dx / std::fabs(b.dx));
  int which;
  bool bound;
  double stpf;
  if (fp > b.fx) {
    // higher value: minimum bracketed; cubic step if closer to stx, else midway to the quadratic one
    which = 1;
    bound = true;
    const double theta = 3 * (b.fx - fp) / (stp - b.stx) + b.dx + dp;
    const double s = max2(std::fabs(theta), max2(std::fabs(b.dx), std::fabs(dp)));
    double gamma = s * std::sqrt((theta / s) * (theta / s) - b.dx / s * (dp / s));
    if (stp < b.stx) gamma = -gamma;
    const double p = gamma - b.dx + theta;
    const double q = gamma - b.dx + gamma + dp;
    const double r = p / q;
    const double stpc = b.stx + r * (stp - b.stx);
    const double stpq = b.stx + b.dx / ((b.fx - fp) / (stp - b.stx) + b.dx) / 2 * (stp - b.stx);
    stpf = std::fabs(stpc - b.stx) < std::fabs(stpq - b.stx) ? stpc : stpc + (stpq - stpc) / 2;
    b.bracketed = true;
  } else if (sgnd < 0.0) {
    // lower value, derivatives of opposite sign: bracketed; the farther of cubic and secant
    which = 2;
    bound = false;
    const double theta = 3 * (b.fx - fp) / (stp - b.stx) + b.dx + dp;
    const double s = max2(std::fabs(theta), max2(std::fabs(b.dx), std::fabs(dp)));
    double gamma = s * std::sqrt((theta / s) * (theta / s) - b.dx / s * (dp / s));
    if (stp > b.stx) gamma = -gamma;
    const double p = gamma - dp + theta;
    const double q = gamma - dp + gamma + b.dx;
    const double r = p / q;
    const double stpc = stp + r * (b.stx - stp);
    const double stpq = stp + dp / (dp - b.dx) * (b.stx - stp);
    stpf = std::fabs(stpc - stp) > std::fabs(stpq - stp) ? stpc : stpq;
    b.bracketed = true;
  } else if (std::fabs(dp) < std::fabs(b.dx)) {
    // lower value, same sign, derivative shrinking: cubic only if it tends to infinity in the step
    // direction or its minimum lies beyond stp
    which = 3;
    bound = true;
    const double theta = 3 * (b.fx - fp) / (stp - b.stx) + b.dx + dp;
    const double s = max2(std::fabs(theta), max2(std::fabs(b.dx), std::fabs(dp)));
    double gamma = s * std::sqrt(max2(0.0, (theta / s) * (theta / s) - b.dx / s * (dp / s)));
    if (stp > b.stx) gamma = -gamma;
    const double p = gamma - dp + theta;
    const double q = gamma + (b.dx - dp) + gamma;
    const double r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (b.stx - stp);
    else stpc = stp > b.stx ? stmax : stmin;
    const double stpq = stp + dp / (dp - b.dx) * (b.stx - stp);
    if (b.bracketed) stpf = std::fabs(stp - stpc) < std::fabs(stp - stpq) ? stpc : stpq;
    else stpf = std::fabs(stp - stpc) > std::fabs(stp - stpq) ? stpc : stpq;
  } else {
    // lower value, same sign, derivative not shrinking
    which = 4;
    bound = false;
    if (b.bracketed) {
      const double theta = 3 * (fp - b.fy) / (b.sty - stp) + b.dy + dp;
      const double s = max2(std::fabs(theta), max2(std::fabs(b.dy), std::fabs(dp)));
      double gamma = s * std::sqrt((theta / s) * (theta / s) - b.dy / s * (dp / s));
      if (stp > b.sty) gamma = -gamma;
      const double p = gamma - dp + theta;
      const double q = gamma - dp + gamma + b.dy;
      const double r = p / q;
      stpf = stp + r * (b.sty - stp);
    } else {
      stpf = stp > b.stx ? stmax : stmin;
    }
  }
  // the interval update does not depend on the case
  if (fp > b.fx) {
    b.sty = stp; b.fy = fp; b.dy = dp;
  } else {
    if (sgnd < 0.0) { b.sty = b.stx; b.fy = b.fx; b.dy = b.dx; }
    b.stx = stp; b.fx = fp; b.dx = dp;
  }
  stpf = min2(stmax, stpf);
  stpf = max2(stmin, stpf);
  stp = stpf;
  if (b.bracketed && bound) {
    const double lim = b.stx + 0.66 * (b.sty - b.stx);
    stp = b.sty > b.stx ? min2(lim, stp) : max2(lim, stp);
  }
  return which;
}

// Line search along the unit direction d from x0 (More-Thuente, ALGLIB's mcsrch).  On entry f and
// dginit = <g(x0), d> are the objective and its slope at x0 and stp the first trial step; on exit
// x, f, g belong to the last point evaluated and stp is its step.  info: 1 Wolfe conditions hold,
// 2 interval below xtol, 3 evaluation budget, 4 step at the lower bound, 5 step at the upper bound,
// 6 rounding / no progress, 0 d is not a descent direction (nothing evaluated, nfev left as it was).
template <class B>
void line_search(B& be, typename B::Vec x0, double dginit, typename B::Vec x, typename B::Vec g, double& f,
                 typename B::Vec d, double& stp, double stpmax, double gtol, double trim_threshold, int& info,
                 int& nfev) {
  if (stpmax == 0.0) stpmax = kStpMaxDefault;
  if (stp < kStpMin) stp = kStpMin;
  if (stp > stpmax) stp = stpmax;
  info = 0;
  if (stpmax < kStpMin && stpmax > 0.0) {
    info = 5;
    stp = stpmax;
    return;
  }
  if (be.size() <= 0 || stp <= 0.0 || stpmax < kStpMin) return;
  if (dginit >= 0.0) return;
  Bracket br;
  br.bracketed = false;
  bool stage1 = true;
  int infoc = 1;
  nfev = 0;
  const double finit = f;
  const double dgtest = kFtol * dginit;
  double width = stpmax - kStpMin;
  double width1 = width / 0.5;
  br.stx = 0.0; br.fx = finit; br.dx = dginit;
  br.sty = 0.0; br.fy = finit; br.dy = dginit;
  for (;;) {
    double stmin, stmax;
    if (br.bracketed) {
      if (br.stx < br.sty) { stmin = br.stx; stmax = br.sty; }
      else { stmin = br.sty; stmax = br.stx; }
    } else {
      stmin = br.stx;
      stmax = stp + 4.0 * (stp - br.stx);
    }
    if (stp > stpmax) stp = stpmax;
    if (stp < kStpMin) stp = kStpMin;
    // unusual termination ahead: fall back to the best step so far
    if ((br.bracketed && (stp <= stmin || stp >= stmax)) || nfev >= kMaxFev - 1 || infoc == 0 ||
        (br.bracketed && stmax - stmin <= kXtol * stmax))
      stp = br.stx;
    double dg = 0.0, moved = 0.0;
    be.trial(x, x0, stp, d, g, &f, &dg, &moved);
    if (f >= trim_threshold) {  // trimfunction: bounded from above near singularities
      f = trim_threshold;
      be.zero(g);
      dg = be.dot(g, d);
    }
    info = 0;
    nfev += 1;
    const double ftest1 = finit + stp * dgtest;
    if ((br.bracketed && (stp <= stmin || stp >= stmax)) || infoc == 0) info = 6;
    if (stp == stpmax && f < finit && f <= ftest1 && dg <= dgtest) info = 5;
    if (stp == kStpMin && (f >= finit || f > ftest1 || dg >= dgtest)) info = 4;
    if (nfev >= kMaxFev) info = 3;
    if (br.bracketed && stmax - stmin <= kXtol * stmax) info = 2;
    if (f < finit && f <= ftest1 && std::fabs(dg) <= -gtol * dginit) info = 1;
    if (info != 0) {
      if ((info == 1 || info == 5) && (f >= finit || moved == 0.0)) info = 6;
      return;
    }
    if (stage1 && f <= ftest1 && dg >= min2(kFtol, gtol) * dginit) stage1 = false;
    if (stage1 && f <= br.fx && f > ftest1) {
      // not enough decrease yet: work on f(x0 + t d) - f(x0) - ftol * t * dginit
      Bracket m = br;
      m.fx = br.fx - br.stx * dgtest; m.fy = br.fy - br.sty * dgtest;
      m.dx = br.dx - dgtest;          m.dy = br.dy - dgtest;
      infoc = trial_step(m, stp, f - stp * dgtest, dg - dgtest, stmin, stmax);
      br.stx = m.stx; br.sty = m.sty; br.bracketed = m.bracketed;
      br.fx = m.fx + m.stx * dgtest; br.fy = m.fy + m.sty * dgtest;
      br.dx = m.dx + dgtest;         br.dy = m.dy + dgtest;
    } else {
      infoc = trial_step(br, stp, f, dg, stmin, stmax);
    }
    if (br.bracketed) {  // force a sufficient decrease of the interval
      if (std::fabs(br.sty - br.stx) >= 0.66 * width1) stp = br.stx + 0.5 * (br.sty - br.stx);
      width1 = width;
      width = std::fabs(br.sty - br.stx);
    }
  }
}

}  // namespace cg_detail

// Minimises the backend's objective from x (overwritten with the result).  Scratch: five vectors
// (the second point of the x rotation, two gradients, the CG direction, the unit search direction);
// ALGLIB keeps eleven (xk, xn, dk, dn, x, d, g, yk, work0, work1, s): the copies between them are
// pointer rotations here, y_k and d_{k+1} are formed on the fly.
constexpr int kCgScratchVectors = 5;
template <class B>
CgReport cg_minimize(B& be, typename B::Vec x_inout, typename B::Vec* scratch, CgOptions opt) {
  using namespace cg_detail;
  using Vec = typename B::Vec;
  if (opt.epsg == 0.0 && opt.epsf == 0.0 && opt.epsx == 0.0 && opt.maxits == 0) opt.epsx = 1e-6;
  CgReport rep;
  Vec xk = x_inout, xt = scratch[0];  // current point, trial point of the line search
  Vec gk = scratch[1], gt = scratch[2];
  Vec dk = scratch[3], d = scratch[4];
  const long long n = be.size();
  double f = 0.0;
  be.eval(xk, gk, &f);
  const double trim_threshold = 10 * (std::fabs(f) + 1);
  be.neg_copy(dk, gk);
  rep.f = f;
  if (std::sqrt(be.sum_sq(gk)) <= opt.epsg) {
    rep.termination = 4;
    return rep;
  }
  rep.nfev = 1;
  double fold = f;
  double last_good_step = 1.0;
  int restart_timer = kRestartCountdown;
  int nfev = 0;  // of the last line search that evaluated anything (mincgstate.nfev)
  double mx = be.max_abs(dk);
  for (;;) {
    // unit direction; the first trial step is the length of the previous accepted step
    double stp = 1.0, dginit = 0.0, dd = 0.0;
    be.normalize_to(d, dk, mx, gk, &stp, &dginit, &dd);
    if (last_good_step != 0.0) stp = last_good_step;
    int info = 0;
    line_search(be, xk, dginit, xt, gt, f, d, stp, 0.0, kGtol, trim_threshold, info, nfev);
    if (info == 0) {  // nothing was evaluated: the "new" point is the current one
      be.copy(xt, xk);
      be.copy(gt, gk);
    }
    double beta = 0.0;
    if (info == 1) {
      double dy, gg, gy;
      be.beta_terms(gt, gk, dk, &dy, &gg, &gy);
      beta = max2(0.0, min2(gg / dy, gy / dy));  // min(Dai-Yuan, Hestenes-Stiefel), clipped at 0
    } else {
      rep.restarts += 1;
    }
    if (rep.iterations > 0 && rep.iterations % (3 + n) == 0) beta = 0.0;
    if (info == 1 || info == 5) restart_timer = kRestartCountdown;
    else restart_timer -= 1;
    double gg;
    be.direction(dk, gt, beta, &gg, &mx);
    const double step_len = stp * std::sqrt(dd);
    if (info == 1) last_good_step = step_len;
    rep.f = f;
    Vec t = xk; xk = xt; xt = t;  // the accepted point becomes the current one
    t = gk; gk = gt; gt = t;
    if (!std::isfinite(gg) || !std::isfinite(f)) { rep.termination = -8; break; }
    rep.nfev += nfev;
    rep.iterations += 1;
    if (rep.iterations >= opt.maxits && opt.maxits > 0) { rep.termination = 5; break; }
    if (std::sqrt(gg) <= opt.epsg) { rep.termination = 4; break; }
    if (fold - f <= opt.epsf * max2(std::fabs(fold), max2(std::fabs(f), 1.0))) { rep.termination = 1; break; }
    if (step_len <= opt.epsx) { rep.termination = 2; break; }
    if (restart_timer <= 0) { rep.termination = 7; break; }
    fold = f;
  }
  if (xk != x_inout) be.copy(x_inout, xk);  // mincgresults: the last point of the last line search
  return rep;
}

// ALGLIB's minlbfgs as the reference configures it (RunLBFGSSolverAnalyticalDiff,
// alglib_objective.cpp:111-140; optimization.cpp:21640-22330): m correction pairs, default
// preconditioner (gamma_k scaling), no step bound, the same line search with gtol = 0.4.  First
// step min(1/|g|, 1), later ones 1; a line search that does not end on the Wolfe conditions
// restarts from the antigradient WITHOUT advancing the pair counter.  Scratch: 5 + 2 m vectors.
// Termination types as in CgReport plus -2 (s'y or y'y rounded to zero).
inline int lbfgs_scratch_vectors(int m) { return 5 + 2 * m; }

template <class B>
CgReport lbfgs_minimize(B& be, typename B::Vec x_inout, typename B::Vec* scratch, int m, CgOptions opt) {
  using namespace cg_detail;
  using Vec = typename B::Vec;
  if (opt.epsg == 0.0 && opt.epsf == 0.0 && opt.epsx == 0.0 && opt.maxits == 0) opt.epsx = 1e-6;
  CgReport rep;
  Vec xk = x_inout, xt = scratch[0];   // current point (x0 of the line search), trial point
  Vec g = scratch[1];
  Vec d = scratch[2], du = scratch[3];  // search direction and its normalised copy
  Vec work = scratch[4];
  Vec* sk = scratch + 5;
  Vec* yk = scratch + 5 + m;
  double rho[64], theta[64];
  if (m < 1 || m > 64) { rep.termination = -1; return rep; }
  double f = 0.0;
  be.eval(xk, g, &f);
  const double trim_threshold = 10 * (std::fabs(f) + 1);
  rep.nfev = 1;
  rep.f = f;
  double fold = f;
  if (std::sqrt(be.sum_sq(g)) <= opt.epsg) {
    rep.termination = 4;
    return rep;
  }
  be.neg_copy(d, g);
  double stp = min2(1.0 / std::sqrt(be.dot(g, g)), 1.0);
  int k = 0;
  int nfev = 0;
  for (;;) {
    const int p = k % m;
    const int q = k < m - 1 ? k : m - 1;
    be.neg_copy(sk[p], xk);
    be.neg_copy(yk[p], g);
    if (k != 0) stp = 1.0;
    double dginit = 0.0, dd = 0.0;
    be.normalize_to(du, d, be.max_abs(d), g, &stp, &dginit, &dd);
    int info = 0;
    // g is overwritten by the line search: its slope at x0 was taken above
    line_search(be, xk, dginit, xt, g, f, du, stp, 0.0, kGtolLbfgs, trim_threshold, info, nfev);
    if (info == 0) be.copy(xt, xk);  // nothing evaluated: the new point is the current one
    rep.nfev += nfev;
    rep.iterations += 1;
    rep.f = f;
    be.add(sk[p], xt);
    be.add(yk[p], g);
    { Vec t = xk; xk = xt; xt = t; }
    const double gg = be.sum_sq(g);
    if (!std::isfinite(gg) || !std::isfinite(f)) { rep.termination = -8; break; }
    if (rep.iterations >= opt.maxits && opt.maxits > 0) { rep.termination = 5; break; }
    if (std::sqrt(gg) <= opt.epsg) { rep.termination = 4; break; }
    if (fold - f <= opt.epsf * max2(std::fabs(fold), max2(std::fabs(f), 1.0))) { rep.termination = 1; break; }
    if (std::sqrt(be.sum_sq(sk[p])) <= opt.epsx) { rep.termination = 2; break; }
    if (info != 1) {
      // no Wolfe point: skip the update, restart from the antigradient
      rep.restarts += 1;
      fold = f;
      be.neg_copy(d, g);
      continue;
    }
    const double sy = be.dot(yk[p], sk[p]);
    const double yy = be.dot(yk[p], yk[p]);
    if (sy == 0.0 || yy == 0.0) { rep.termination = -2; break; }
    rho[p] = 1 / sy;
    const double gamma = sy / yy;
    // two-loop recursion: work = H_{k+1} g
    be.copy(work, g);
    for (int i = k; i >= k - q; --i) {
      const int ic = i % m;
      const double v = be.dot(sk[ic], work);
      theta[ic] = v;
      be.sub_scaled(work, v * rho[ic], yk[ic]);
    }
    be.scale(work, gamma);
    for (int i = k - q; i <= k; ++i) {
      const int ic = i % m;
      const double v = be.dot(yk[ic], work);
      be.add_scaled(work, rho[ic] * (-v + theta[ic]), sk[ic]);
    }
    be.neg_copy(d, work);
    fold = f;
    k += 1;
  }
  if (xk != x_inout) be.copy(x_inout, xk);
  return rep;
}

// IRLSMapSolver::RunIRLSLoop (irls_map_solver.cpp:45-157): conjugate-gradient solves of the
// re-weighted least-squares problem until the cost of two consecutive solves differs by less than
// the threshold.  The backend's objective must use the weights that be.reweight(x) installs
// (w = 1 / max(1e-5, reg(x)), :128-143); the caller resets them to 1 beforehand (:66-74).
struct IrlsReport {
  int irls_iterations = 0;
  int solver_iterations = 0;  // summed over the outer iterations
  int nfev = 0;
  int last_termination = 0;
  double f = 0.0;             // cost the last CG solve returned
};

// lbfgs_corrections = 0: conjugate gradients (CG_SOLVER, the default); m > 0: L-BFGS with m pairs
// (LBFGS_SOLVER; scratch must then hold lbfgs_scratch_vectors(m) vectors).
template <class B>
IrlsReport irls_solve(B& be, typename B::Vec x_inout, typename B::Vec* scratch, const CgOptions& opt,
                      int max_irls_iterations, double cost_difference_threshold, bool has_regularizer,
                      int lbfgs_corrections = 0) {
  IrlsReport out;
  double previous_cost = INFINITY;
  double cost_difference = cost_difference_threshold + 1.0;
  while (std::fabs(cost_difference) >= cost_difference_threshold) {
    const CgReport rep = lbfgs_corrections > 0 ? lbfgs_minimize(be, x_inout, scratch, lbfgs_corrections, opt)
                                               : cg_minimize(be, x_inout, scratch, opt);
    out.solver_iterations += rep.iterations;
    out.nfev += rep.nfev;
    out.last_termination = rep.termination;
    out.f = rep.f;
    if (!has_regularizer) break;  // nothing to re-weight: one solve (:118-121)
    be.reweight(x_inout);
    cost_difference = previous_cost - rep.f;
    previous_cost = rep.f;
    out.irls_iterations += 1;
    if (max_irls_iterations > 0 && out.irls_iterations >= max_irls_iterations) break;
  }
  return out;
}

}  // namespace srb
