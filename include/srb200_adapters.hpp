// srb200_adapters.hpp -- the C++ adapters that put libsrb200 (include/srb200.h) behind the
// reference's own virtual interfaces.  This is the reference-side code INTEGRATION.md describes:
// it is written against the reference's headers (src/optimization/objective_function.h,
// src/optimization/regularizer.h) and is compiled against the REAL ones in oracle/_ref, where the
// reference's unmodified ALGLIB / solver translation units run on top of it (tests/test_gpu_solver.py).
//
// Error behaviour: the reference aborts through glog CHECK; the adapters CHECK on the C-ABI status.
#ifndef SRB200_ADAPTERS_HPP_
#define SRB200_ADAPTERS_HPP_

#include <utility>
#include <vector>

#include "glog/logging.h"
#include "optimization/objective_function.h"  // super_resolution::ObjectiveTerm
#include "optimization/regularizer.h"         // super_resolution::Regularizer
#include "srb200.h"

namespace super_resolution {

#define SRB_CHECK_OK(ctx, call) \
  CHECK((call) == SRB_OK) << "libsrb200: " << srb_last_error(ctx)

// The WHOLE objective -- data term over this context's frames plus the IRLS-weighted
// regularization term -- as one ObjectiveTerm backed by one fused device evaluation (srb_eval).
// Replaces ObjectiveDataTerm (constructed at irls_map_solver.cpp:243-246) together with the
// ObjectiveIRLSRegularizationTerm the IRLS loop adds (irls_map_solver.cpp:83-93).
//
// ObjectiveFunction::ComputeAllTerms zeroes the gradient and lets terms ADD into it
// (objective_function.cpp:9-19).  This term must be the FIRST term: it writes the gradient
// (equivalent to adding into the zeroed array, without a second host pass over C*H*W doubles).
class CudaObjectiveTerm : public ObjectiveTerm {
 public:
  explicit CudaObjectiveTerm(srb_ctx* ctx) : ctx_(ctx) { CHECK_NOTNULL(ctx); }
  double Compute(const double* estimated_image_data, double* gradient) const override {
    CHECK_NOTNULL(estimated_image_data);
    double cost = 0.0;
    SRB_CHECK_OK(ctx_, srb_eval(ctx_, estimated_image_data, gradient, &cost));
    return cost;
  }

 private:
  srb_ctx* ctx_;
};

// The same term with ONE host thread driving several GPUs (srb_multi_eval): what the reference's single-threaded
// solver holds when more than one B200 is available.  Same contract as CudaObjectiveTerm.
class CudaMultiObjectiveTerm : public ObjectiveTerm {
 public:
  explicit CudaMultiObjectiveTerm(srb_multi* multi) : multi_(multi) { CHECK_NOTNULL(multi); }
  double Compute(const double* estimated_image_data, double* gradient) const override {
    CHECK_NOTNULL(estimated_image_data);
    double cost = 0.0;
    CHECK(srb_multi_eval(multi_, estimated_image_data, gradient, &cost) == SRB_OK)
        << "libsrb200: " << srb_multi_last_error(multi_);
    return cost;
  }

 private:
  srb_multi* multi_;
};

// Only the data term (ObjectiveDataTerm::Compute, objective_data_term.cpp:98-116), ADDING into the
// gradient like the reference; evaluated in the reference's operation order (bit-identical).
class CudaObjectiveDataTerm : public ObjectiveTerm {
 public:
  explicit CudaObjectiveDataTerm(srb_ctx* ctx) : ctx_(ctx) { CHECK_NOTNULL(ctx); }
  double Compute(const double* estimated_image_data, double* gradient) const override {
    double cost = 0.0;
    SRB_CHECK_OK(ctx_, srb_data_term(ctx_, estimated_image_data, gradient, &cost));
    return cost;
  }

 private:
  srb_ctx* ctx_;
};

// A Regularizer (regularizer.h:13-50) whose two virtuals run on the device; the kind / parameters
// are the ones configured on the context with srb_set_regularizer.
class CudaRegularizer : public Regularizer {
 public:
  CudaRegularizer(const cv::Size& image_size, srb_ctx* ctx) : Regularizer(image_size), ctx_(ctx) {}
  std::vector<double> ApplyToImage(const double* image_data, const int num_channels) const override {
    std::vector<double> values((size_t)image_size_.area() * num_channels);
    SRB_CHECK_OK(ctx_, srb_reg_apply(ctx_, image_data, num_channels, values.data()));
    return values;
  }
  std::pair<std::vector<double>, std::vector<double>> ApplyToImageWithDifferentiation(
      const double* image_data, const std::vector<double>& gradient_constants,
      const int num_channels) const override {
    const size_t n = (size_t)image_size_.area() * num_channels;
    CHECK_EQ(gradient_constants.size(), n);
    std::vector<double> values(n), partials(n);
    SRB_CHECK_OK(ctx_, srb_reg_apply_diff(ctx_, image_data, gradient_constants.data(), num_channels,
                                          values.data(), partials.data()));
    return std::make_pair(std::move(values), std::move(partials));
  }

 private:
  srb_ctx* ctx_;
};

}  // namespace super_resolution
#endif  // SRB200_ADAPTERS_HPP_
