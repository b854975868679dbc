// ref_shim.cpp -- C entry points over the REFERENCE'S OWN translation units, compiled unmodified
// from /root/reference into oracle/_ref/libsr_ref.so.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// What is the real reference here (compiled where it lies, see oracle/Makefile):
//   libs/alglib/src/{ap,alglibinternal,alglibmisc,linalg,solvers,optimization}.cpp  (ALGLIB 3.10.0)
//   src/optimization/{tv_regularizer,btv_regularizer,objective_function,
//                     objective_irls_regularization_term,alglib_objective,map_solver,
//                     irls_map_solver}.cpp
// What is NOT (needs OpenCV C++, absent from this image): ImageData / ImageModel (stubs/ holds
// minimal containers with the members the files above use) and ObjectiveDataTerm, whose
// constructor/Compute are defined below on top of the oracle restatement (sr_oracle.c) -- or on
// top of a caller-supplied callback, which is how the GPU tests drive the UNMODIFIED reference
// IRLS + ALGLIB loop with the CUDA engine plugged in behind the ObjectiveTerm / Regularizer seams.
#include <chrono>
#include <cstring>
#include <limits>
#include <cmath>
#include <memory>
#include <utility>
#include <vector>

#include "image/image_data.h"
#include "image_model/image_model.h"
#include "optimization/btv_regularizer.h"
#include "optimization/irls_map_solver.h"
#include "optimization/objective_data_term.h"
#include "optimization/objective_function.h"
#include "optimization/objective_irls_regularization_term.h"
#include "optimization/tv_regularizer.h"
#include "sr_oracle.h"
#include "optimization/alglib_objective.h"

#include "glog/logging.h"

#include "ref_shim.h"

namespace {
struct SolveContext {
  sro_model model;
  const double* obs_hr = nullptr;  // [N][C_total][H][W]
  int C_total = 0;
  int num_threads = 1;
  const ref_callbacks* cbs = nullptr;
  ref_stats stats{};
};
SolveContext* g_ctx = nullptr;
}  // namespace

namespace super_resolution {

// objective_data_term.cpp:77-96 (constructor) and :98-116 (Compute), with the per-observation
// body (:15-75) delegated to the oracle restatement or to the plugged-in engine.
ObjectiveDataTerm::ObjectiveDataTerm(const ImageModel& image_model,
                                     const std::vector<ImageData>& observations,
                                     const int channel_start, const int channel_end,
                                     const cv::Size& image_size)
    : image_model_(image_model), observations_(observations), channel_start_(channel_start),
      channel_end_(channel_end), image_size_(image_size) {
  CHECK_GT(observations.size(), 0) << "Cannot solve with 0 observations.";
  CHECK_GE(channel_start, 0);
  CHECK_LE(channel_end, observations[0].GetNumChannels());
  CHECK_GT(channel_end, channel_start);
}

double ObjectiveDataTerm::Compute(const double* estimated_image_data, double* gradient) const {
  CHECK_NOTNULL(estimated_image_data);
  CHECK(g_ctx != nullptr);
  const auto t0 = std::chrono::steady_clock::now();
  double cost;
  if (g_ctx->cbs && g_ctx->cbs->data_term) {
    cost = g_ctx->cbs->data_term(estimated_image_data, gradient, channel_start_, channel_end_,
                                 g_ctx->cbs->user);
  } else {
    cost = sro_data_term(&g_ctx->model, estimated_image_data, image_size_.height,
                         image_size_.width, channel_end_ - channel_start_, g_ctx->obs_hr,
                         g_ctx->C_total, channel_start_, gradient, g_ctx->num_threads);
  }
  g_ctx->stats.num_data_term_evals++;
  g_ctx->stats.seconds_in_data_term +=
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return cost;
}

// A Regularizer whose two virtuals forward to caller-supplied functions: the seam a maintainer
// would subclass to put the CUDA engine behind ObjectiveIRLSRegularizationTerm (regularizer.h:26-45).
class CallbackRegularizer : public Regularizer {
 public:
  CallbackRegularizer(const cv::Size& size, const ref_callbacks* cbs)
      : Regularizer(size), cbs_(cbs) {}
  std::vector<double> ApplyToImage(const double* image_data, const int num_channels) const override {
    std::vector<double> values((size_t)image_size_.area() * num_channels);
    cbs_->reg_apply(image_data, num_channels, values.data(), cbs_->user);
    return values;
  }
  std::pair<std::vector<double>, std::vector<double>> ApplyToImageWithDifferentiation(
      const double* image_data, const std::vector<double>& gradient_constants,
      const int num_channels) const override {
    const size_t n = (size_t)image_size_.area() * num_channels;
    std::vector<double> values(n), partials(n);
    cbs_->reg_apply_diff(image_data, gradient_constants.data(), num_channels, values.data(),
                         partials.data(), cbs_->user);
    return std::make_pair(std::move(values), std::move(partials));
  }
 private:
  const ref_callbacks* cbs_;
};

}  // namespace super_resolution

using namespace super_resolution;  // NOLINT

static std::shared_ptr<Regularizer> MakeRegularizer(int kind, int R, double decay, int H, int W) {
  const cv::Size size(W, H);
  if (kind == 2) return std::make_shared<BilateralTotalVariationRegularizer>(size, R, decay);
  auto tv = std::make_shared<TotalVariationRegularizer>(size);
  tv->SetUse3dTotalVariation(kind == 1);
  return tv;
}

extern "C" {

// Regularizer::ApplyToImage of the reference's own classes.
void ref_reg_apply(int kind, int R, double decay, const double* x, int H, int W, int C,
                   double* values) {
  const std::vector<double> v = MakeRegularizer(kind, R, decay, H, W)->ApplyToImage(x, C);
  std::memcpy(values, v.data(), v.size() * sizeof(double));
}

// Regularizer::ApplyToImageWithDifferentiation of the reference's own classes.
void ref_reg_apply_diff(int kind, int R, double decay, const double* x, const double* constants,
                        int H, int W, int C, double* values, double* partials) {
  const std::vector<double> cst(constants, constants + (size_t)H * W * C);
  const auto vp = MakeRegularizer(kind, R, decay, H, W)->ApplyToImageWithDifferentiation(x, cst, C);
  std::memcpy(values, vp.first.data(), vp.first.size() * sizeof(double));
  std::memcpy(partials, vp.second.data(), vp.second.size() * sizeof(double));
}

// ObjectiveIRLSRegularizationTerm::Compute of the reference (gradient accumulated, may be NULL).
double ref_irls_term(int kind, int R, double decay, double lambda, const double* weights,
                     const double* x, int H, int W, int C, double* gradient) {
  const std::vector<double> w(weights, weights + (size_t)H * W * C);
  const cv::Size size(W, H);
  ObjectiveIRLSRegularizationTerm term(MakeRegularizer(kind, R, decay, H, W), lambda, w, C, size);
  return term.Compute(x, gradient);
}

// ObjectiveFunction::ComputeAllTerms of the reference: oracle data term + reference IRLS term.
double ref_compute_all_terms(const sro_model* m, const double* x, int H, int W, int C,
                             const double* obs_hr, int reg_kind, int R, double decay,
                             double lambda, const double* weights, double* gradient,
                             int num_threads) {
  SolveContext ctx;
  ctx.model = *m;
  ctx.obs_hr = obs_hr;
  ctx.C_total = C;
  ctx.num_threads = num_threads;
  g_ctx = &ctx;
  ImageModel image_model(m->scale);
  std::vector<ImageData> observations(1, ImageData(obs_hr, cv::Size(W, H), C));
  const cv::Size size(W, H);
  ObjectiveFunction objective(H * W * C);
  objective.AddTerm(std::make_shared<ObjectiveDataTerm>(image_model, observations, 0, C, size));
  std::vector<double> w;
  if (weights && lambda > 0) {
    w.assign(weights, weights + (size_t)H * W * C);
    objective.AddTerm(std::make_shared<ObjectiveIRLSRegularizationTerm>(
        MakeRegularizer(reg_kind, R, decay, H, W), lambda, w, C, size));
  }
  const double f = objective.ComputeAllTerms(x, gradient);
  g_ctx = nullptr;
  return f;
}

void ref_default_options(ref_options* o) {
  const IRLSMapSolverOptions d;
  o->solver = d.least_squares_solver == LBFGS_SOLVER ? 1 : 0;
  o->max_num_solver_iterations = d.max_num_solver_iterations;
  o->max_num_irls_iterations = d.max_num_irls_iterations;
  o->gradient_norm_threshold = d.gradient_norm_threshold;
  o->cost_decrease_threshold = d.cost_decrease_threshold;
  o->parameter_variation_threshold = d.parameter_variation_threshold;
  o->irls_cost_difference_threshold = d.irls_cost_difference_threshold;
  o->split_channels = d.split_channels;
  o->num_lbfgs_hessian_corrections = d.num_lbfgs_hessian_corrections;
  o->use_numerical_differentiation = d.use_numerical_differentiation;
  o->numerical_differentiation_step = d.numerical_differentiation_step;
  o->num_threads = 1;
}

// IRLSMapSolver::Solve of the reference (irls_map_solver.cpp:192-265; ALGLIB inner loop
// alglib_objective.cpp:47-139), driven exactly like SetupAndRunSolver (super_resolution.cpp:126-199).
//   lr  : [N][C][h][w] low-resolution observations;  x0 : [C][H][W] initial estimate
//   reg_kind < 0 or lambda <= 0 => no regulariser is added
//   cbs : NULL => oracle data term + reference regularisers; otherwise the plugged-in engine
int ref_solve(const sro_model* m, const double* lr, int N, int C, int h, int w, const double* x0,
              int reg_kind, int R, double decay, double lambda, const ref_options* opt,
              const ref_callbacks* cbs, double* out, ref_stats* stats) {
  return ref_solve_with_regularizer(m, lr, N, C, h, w, x0, reg_kind, R, decay, lambda, opt, cbs, nullptr, out, stats);
}

}  // extern "C"

// The same with a caller-built Regularizer object handed to IRLSMapSolver::AddRegularizer
// (map_solver.h:85-93) -- how oracle/ref_fused.cpp passes the product's CudaRegularizer adapter to the
// reference's unmodified solver.  C++ linkage: both sides are built by the same compiler.
int ref_solve_with_regularizer(const sro_model* m, const double* lr, int N, int C, int h, int w, const double* x0,
                               int reg_kind, int R, double decay, double lambda, const ref_options* opt,
                               const ref_callbacks* cbs, std::shared_ptr<super_resolution::Regularizer> reg_override,
                               double* out, ref_stats* stats) {
  const auto t0 = std::chrono::steady_clock::now();
  const int s = m->scale;
  const int H = h * s, W = w * s;
  const size_t P = (size_t)H * W, p = (size_t)h * w;

  IRLSMapSolverOptions options;
  options.least_squares_solver = opt->solver == 1 ? LBFGS_SOLVER : CG_SOLVER;
  options.max_num_solver_iterations = opt->max_num_solver_iterations;
  options.max_num_irls_iterations = opt->max_num_irls_iterations;
  options.gradient_norm_threshold = opt->gradient_norm_threshold;
  options.cost_decrease_threshold = opt->cost_decrease_threshold;
  options.parameter_variation_threshold = opt->parameter_variation_threshold;
  options.irls_cost_difference_threshold = opt->irls_cost_difference_threshold;
  options.split_channels = opt->split_channels != 0;
  options.num_lbfgs_hessian_corrections = opt->num_lbfgs_hessian_corrections;
  options.use_numerical_differentiation = opt->use_numerical_differentiation != 0;
  options.numerical_differentiation_step = opt->numerical_differentiation_step;

  ImageModel image_model(s);
  std::vector<ImageData> low_res_images;
  for (int k = 0; k < N; ++k)
    low_res_images.push_back(ImageData(lr + (size_t)k * C * p, cv::Size(w, h), C));

  // Contiguous nearest-upsampled observations for the oracle data term (map_solver.cpp:81-85).
  std::vector<double> obs_hr;
  if (!(cbs && cbs->data_term)) {
    obs_hr.resize((size_t)N * C * P);
    for (int k = 0; k < N; ++k)
      for (int c = 0; c < C; ++c)
        sro_resize_nearest(lr + ((size_t)k * C + c) * p, h, w, obs_hr.data() + ((size_t)k * C + c) * P,
                           H, W);
  }

  SolveContext ctx;
  ctx.model = *m;
  ctx.obs_hr = obs_hr.empty() ? nullptr : obs_hr.data();
  ctx.C_total = C;
  ctx.num_threads = opt->num_threads;
  ctx.cbs = cbs;
  g_ctx = &ctx;

  IRLSMapSolver solver(options, image_model, low_res_images, /*print_solver_output=*/false);
  if (reg_kind >= 0 && lambda > 0) {
    std::shared_ptr<Regularizer> reg;
    if (reg_override)
      reg = reg_override;
    else if (cbs && cbs->reg_apply && cbs->reg_apply_diff)
      reg = std::make_shared<CallbackRegularizer>(cv::Size(W, H), cbs);
    else
      reg = MakeRegularizer(reg_kind, R, decay, H, W);
    solver.AddRegularizer(reg, lambda);
  }
  const ImageData initial_estimate(x0, cv::Size(W, H), C);
  const ImageData result = solver.Solve(initial_estimate);
  for (int c = 0; c < C; ++c) std::memcpy(out + (size_t)c * P, result.GetChannelData(c), P * sizeof(double));

  ctx.stats.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = ctx.stats;
  g_ctx = nullptr;
  return 0;
}

extern "C" {

// ---- ALGLIB's mincg on an arbitrary objective (C callback), configured exactly as
// RunCGSolverAnalyticalDiff does (alglib_objective.cpp:47-75).  The checker of the device-resident
// CG restatement (super-resolution_b200/csrc/srb_cg.h): tests/test_cg_restatement.py runs both on the
// same objectives and compares iterates bit for bit.
typedef void (*ref_fg_cb)(long long n, const double* x, double* f, double* g, void* user);
namespace {
struct FgClosure { ref_fg_cb cb; void* user; };
void FgThunk(const alglib::real_1d_array& x, double& func, alglib::real_1d_array& grad, void* ptr) {
  const FgClosure* c = reinterpret_cast<const FgClosure*>(ptr);
  c->cb((long long)x.length(), x.getcontent(), &func, grad.getcontent(), c->user);
}
void NoReport(const alglib::real_1d_array&, double, void*) {}
}  // namespace

// report: [iterations, nfev, termination type, final f (mincgstate.f)]
int ref_mincg(long long n, double* x_inout, double epsg, double epsf, double epsx, int maxits, ref_fg_cb cb,
              void* user, double* report) {
  alglib::real_1d_array x;
  x.setcontent((alglib::ae_int_t)n, x_inout);
  alglib::mincgstate state;
  alglib::mincgreport rep;
  alglib::mincgcreate(x, state);
  alglib::mincgsetcond(state, epsg, epsf, epsx, maxits);
  alglib::mincgsetxrep(state, true);
  FgClosure closure{cb, user};
  alglib::mincgoptimize(state, FgThunk, NoReport, &closure);
  alglib::mincgresults(state, x, rep);
  std::memcpy(x_inout, x.getcontent(), (size_t)n * sizeof(double));
  report[0] = (double)rep.iterationscount;
  report[1] = (double)rep.nfev;
  report[2] = (double)rep.terminationtype;
  report[3] = state.f;
  return 0;
}

// ALGLIB's minlbfgs configured as RunLBFGSSolverAnalyticalDiff does (alglib_objective.cpp:111-140).
// report: [iterations, nfev, termination type, final f (minlbfgsstate.f)]
int ref_minlbfgs(long long n, double* x_inout, int m, double epsg, double epsf, double epsx, int maxits,
                 ref_fg_cb cb, void* user, double* report) {
  alglib::real_1d_array x;
  x.setcontent((alglib::ae_int_t)n, x_inout);
  alglib::minlbfgsstate state;
  alglib::minlbfgsreport rep;
  alglib::minlbfgscreate(m, x, state);
  alglib::minlbfgssetcond(state, epsg, epsf, epsx, maxits);
  alglib::minlbfgssetxrep(state, true);
  FgClosure closure{cb, user};
  alglib::minlbfgsoptimize(state, FgThunk, NoReport, &closure);
  alglib::minlbfgsresults(state, x, rep);
  std::memcpy(x_inout, x.getcontent(), (size_t)n * sizeof(double));
  report[0] = (double)rep.iterationscount;
  report[1] = (double)rep.nfev;
  report[2] = (double)rep.terminationtype;
  report[3] = state.f;
  return 0;
}

}  // extern "C"

