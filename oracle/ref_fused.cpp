// ref_fused.cpp -- the reference's own solver (oracle/_ref/libsr_ref.so: ALGLIB 3.10.0, IRLSMapSolver,
// ObjectiveFunction, ... compiled unmodified from /root/reference) with the PRODUCT plugged in behind its
// seams through the C++ adapters of include/srb200_adapters.hpp.  Built into its own library
// (oracle/_ref/libsr_ref_fused.so) so that the CPU reference arm (libsr_ref.so) never maps libsrb200.so.
// TEST INFRASTRUCTURE, NOT PRODUCT CODE: this is the reference-side code INTEGRATION.md describes, kept here
// so that the tests can run it.
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "image/image_data.h"
#include "image_model/image_model.h"
#include "optimization/irls_map_solver.h"
#include "optimization/objective_function.h"
#include "optimization/alglib_objective.h"
#include "ref_shim.h"
#include "srb200_adapters.hpp"  // include/: the adapters INTEGRATION.md tells a maintainer to add

#include "glog/logging.h"

using namespace super_resolution;  // NOLINT

// IRLSMapSolver::Solve with the B200 engine behind the reference's seams, the way INTEGRATION.md
// wires it: ONE CudaObjectiveTerm (srb_eval = data term + IRLS regularization term, fused) replaces
// the ObjectiveDataTerm built at irls_map_solver.cpp:243-246 and the ObjectiveIRLSRegularizationTerm
// added per outer iteration at :83-93; the weight vector the reference keeps on the host
// (:66-74, :128-143) lives on the device (srb_set_irls_weights / srb_reweight).  The loop below
// restates irls_map_solver.cpp:45-157 and :192-265 around those three substitutions; the inner solve
// is the reference's own RunCGSolverAnalyticalDiff / RunLBFGSSolverAnalyticalDiff + ALGLIB,
// unmodified.  `ctx` already holds the model, the observations and the regularizer
// (srb_create / srb_set_observations / srb_set_regularizer).
namespace {
// The engine behind the loop: one device (srb_ctx) or several driven by this thread (srb_multi).
struct SingleOps {
  srb_ctx* ctx;
  void set_channel_range(int c0, int c1) const { CHECK(srb_set_channel_range(ctx, c0, c1) == SRB_OK) << srb_last_error(ctx); }
  void reweight(const double* x) const { CHECK(srb_reweight(ctx, x, nullptr) == SRB_OK) << srb_last_error(ctx); }
  std::shared_ptr<ObjectiveTerm> term() const { return std::make_shared<CudaObjectiveTerm>(ctx); }
  long evals() const {
    srb_timing tm;
    return srb_get_timing(ctx, &tm) == SRB_OK ? (long)tm.num_evals : 0;
  }
};
struct MultiOps {
  srb_multi* m;
  void set_channel_range(int c0, int c1) const { CHECK(srb_multi_set_channel_range(m, c0, c1) == SRB_OK) << srb_multi_last_error(m); }
  void reweight(const double* x) const { CHECK(srb_multi_reweight(m, x, nullptr) == SRB_OK) << srb_multi_last_error(m); }
  std::shared_ptr<ObjectiveTerm> term() const { return std::make_shared<CudaMultiObjectiveTerm>(m); }
  long evals() const {
    srb_timing tm;
    return srb_multi_get_timing(m, &tm) == SRB_OK ? (long)tm.num_evals : 0;
  }
};

template <class Ops>
int solve_fused_impl(const Ops& eng, int C, int H, int W, const double* x0, int has_regularizer, double lambda_sum,
                     const ref_options* opt, double* out, ref_stats* stats) {
  const auto t0 = std::chrono::steady_clock::now();
  const size_t P = (size_t)H * W;
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
  options.use_numerical_differentiation = false;

  // irls_map_solver.cpp:200-216
  const int per_split = options.split_channels ? 1 : C;
  const int rounds = C / per_split;
  const int num_data_points = per_split * (int)P;
  options.AdjustThresholdsAdaptively(num_data_points, has_regularizer ? lambda_sum : 0.0);
  ref_stats st{};
  for (int i = 0; i < rounds; ++i) {
    const int c0 = i * per_split, c1 = c0 + per_split;
    eng.set_channel_range(c0, c1);  // resets weights to 1 (:66-74)
    alglib::real_1d_array solver_data;  // :232-239
    solver_data.setlength(num_data_points);
    std::memcpy(solver_data.getcontent(), x0 + (size_t)c0 * P, (size_t)num_data_points * sizeof(double));
    // :45-157
    double previous_cost = std::numeric_limits<double>::infinity();
    double cost_difference = options.irls_cost_difference_threshold + 1.0;
    int num_iterations_ran = 0;
    while (std::abs(cost_difference) >= options.irls_cost_difference_threshold) {
      ObjectiveFunction objective_function(num_data_points);
      objective_function.AddTerm(eng.term());
      const double final_cost = options.least_squares_solver == CG_SOLVER
                                    ? RunCGSolverAnalyticalDiff(options, objective_function, &solver_data)
                                    : RunLBFGSSolverAnalyticalDiff(options, objective_function, &solver_data);
      if (!has_regularizer) break;  // :118-121
      // :128-143, on the device: w = 1 / max(1e-5, reg(x))
      eng.reweight(solver_data.getcontent());
      cost_difference = previous_cost - final_cost;
      previous_cost = final_cost;
      num_iterations_ran++;
      if (options.max_num_irls_iterations > 0 && num_iterations_ran >= options.max_num_irls_iterations) break;
    }
    std::memcpy(out + (size_t)c0 * P, solver_data.getcontent(), (size_t)num_data_points * sizeof(double));
  }
  st.num_data_term_evals = eng.evals();
  st.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = st;
  return 0;
}
}  // namespace

extern "C" {

int ref_solve_fused(srb_ctx* ctx, int C, int H, int W, const double* x0, int has_regularizer,
                    double lambda_sum, const ref_options* opt, double* out, ref_stats* stats) {
  return solve_fused_impl(SingleOps{ctx}, C, H, W, x0, has_regularizer, lambda_sum, opt, out, stats);
}

// The same loop with ONE host thread driving several GPUs through CudaMultiObjectiveTerm (srb_multi_eval):
// `multi` already holds the model, the observations and the regularizer.
int ref_solve_fused_multi(srb_multi* multi, int C, int H, int W, const double* x0, int has_regularizer,
                          double lambda_sum, const ref_options* opt, double* out, ref_stats* stats) {
  return solve_fused_impl(MultiOps{multi}, C, H, W, x0, has_regularizer, lambda_sum, opt, out, stats);
}


// IRLSMapSolver::Solve of the reference, UNMODIFIED (irls_map_solver.cpp:45-157, 192-265), with the two
// finer-grained adapters instantiated in C++: CudaRegularizer goes in through AddRegularizer
// (map_solver.h:85-93) and is what ObjectiveIRLSRegularizationTerm calls (values + partials on the device,
// bit-identical); CudaObjectiveDataTerm stands behind the ObjectiveDataTerm the solver constructs itself
// (:243-246; oracle/ref_shim.cpp delegates its Compute to the data_term hook).  With
// srb_set_strict_cost(ctx, 1) cost AND gradient of both terms are bit-identical to the CPU terms, so the
// whole solve must reproduce ref_solve bit for bit (tests/test_gpu_solver.py).
namespace {
struct AdapterHook {
  srb_ctx* ctx;
  CudaObjectiveDataTerm term;
  int c0 = -1, c1 = -1;
  explicit AdapterHook(srb_ctx* c) : ctx(c), term(c) {}
};
double AdapterDataTerm(const double* x, double* grad, int channel_start, int channel_end, void* user) {
  AdapterHook* a = static_cast<AdapterHook*>(user);
  if (a->c0 != channel_start || a->c1 != channel_end) {  // the channel range of the ObjectiveDataTerm (:243-246)
    CHECK(srb_set_channel_range(a->ctx, channel_start, channel_end) == SRB_OK) << srb_last_error(a->ctx);
    a->c0 = channel_start;
    a->c1 = channel_end;
  }
  return a->term.Compute(x, grad);
}
}  // namespace

int ref_solve_adapters(srb_ctx* ctx, const sro_model* m, const double* lr, int N, int C, int h, int w,
                       const double* x0, int has_regularizer, double lambda, const ref_options* opt, double* out,
                       ref_stats* stats) {
  AdapterHook hook(ctx);
  ref_callbacks cbs{};
  cbs.data_term = AdapterDataTerm;
  cbs.user = &hook;
  std::shared_ptr<Regularizer> reg;
  if (has_regularizer) reg = std::make_shared<CudaRegularizer>(cv::Size(w * m->scale, h * m->scale), ctx);
  return ref_solve_with_regularizer(m, lr, N, C, h, w, x0, has_regularizer ? 0 : -1, 0, 0.0, lambda, opt, &cbs,
                                    reg, out, stats);
}

}  // extern "C"
