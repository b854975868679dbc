// srb_kernels_fused.cuh -- fused tile kernel of the MAP objective (fast path).  PLACEHOLDER: the
// first milestone runs everything through the reference-order kernels.
#pragma once
#include "srb_common.cuh"

namespace srb {
inline bool fused_supported(const srb_ctx*) { return false; }
inline srb_status fused_setup(srb_ctx*) { return SRB_OK; }
inline void fused_teardown(srb_ctx*) {}
inline void fused_reg_changed(srb_ctx*) {}
inline srb_status fused_eval(srb_ctx* c, const double*, double*, bool) {
  return c->fail(SRB_ERR_STATE, "fused path not built");
}
}  // namespace srb
