// ref_shim.h -- declarations shared by oracle/ref_shim.cpp (libsr_ref.so: the reference's own sources, no
// dependency on the product) and oracle/ref_fused.cpp (libsr_ref_fused.so: the reference's solver with the
// product's adapters plugged in).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
#pragma once
#include <memory>

#include "optimization/regularizer.h"
#include "sr_oracle.h"

extern "C" {
typedef double (*ref_data_term_cb)(const double* x, double* grad_accum_or_null, int channel_start,
                                   int channel_end, void* user);
typedef void (*ref_reg_apply_cb)(const double* x, int num_channels, double* values, void* user);
typedef void (*ref_reg_apply_diff_cb)(const double* x, const double* constants, int num_channels,
                                      double* values, double* partials, void* user);
typedef struct {
  ref_data_term_cb data_term;          // NULL => oracle restatement
  ref_reg_apply_cb reg_apply;          // NULL => reference regularizer classes
  ref_reg_apply_diff_cb reg_apply_diff;
  void* user;
} ref_callbacks;

typedef struct {
  int solver;  // 0 = CG_SOLVER, 1 = LBFGS_SOLVER (map_solver.h:20-23)
  int max_num_solver_iterations;
  int max_num_irls_iterations;
  double gradient_norm_threshold, cost_decrease_threshold, parameter_variation_threshold;
  double irls_cost_difference_threshold;
  int split_channels;
  int num_lbfgs_hessian_corrections;
  int use_numerical_differentiation;
  double numerical_differentiation_step;
  int num_threads;  // oracle data term threads
} ref_options;

typedef struct {
  long num_data_term_evals;
  double seconds_in_data_term;
  double seconds_total;
} ref_stats;
}


int ref_solve_with_regularizer(const sro_model* m, const double* lr, int N, int C, int h, int w, const double* x0,
                               int reg_kind, int R, double decay, double lambda, const ref_options* opt,
                               const ref_callbacks* cbs, std::shared_ptr<super_resolution::Regularizer> reg_override,
                               double* out, ref_stats* stats);
