/*
 * sr_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, fp64 restatement of the reference's MAP-objective hot path
 * (rteammco/super-resolution).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (super-resolution_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED -- see tests/test_oracle_golden.py: every golden vector
 * the reference's own tests hold for this path plus
 * fixtures produced in the build container by the very OpenCV entry points the
 * reference calls (cv2 4.13: warpAffine / filter2D / resize / getGaussianKernel,
 * tests/golden/make_golden.py) and by the reference's own tv/btv regularizer
 * sources compiled unmodified (oracle/_ref).
 *
 * All images are planar row-major fp64, index c*H*W + r*W + col
 * (reference: src/util/util.cpp:81-89).
 */
#ifndef SR_ORACLE_H_
#define SR_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Image formation model A_k = D * B * M_k (reference: src/image_model/image_model.cpp:17-61). */
typedef struct {
  int scale;             /* downsampling scale s >= 1 */
  int psf_size;          /* K (odd) or 0 when the model has no blur operator */
  const double* psf;     /* K*K row-major correlation kernel (blur_kernel_) */
  int num_frames;        /* N observations */
  const double* shifts;  /* 2*N doubles dx_0, dy_0, dx_1, ...; NULL => no motion operator */
} sro_model;

/* cv::getGaussianKernel(n, sigma, CV_64F), sigma > 0 (reference call: blur_module.cpp:20-21). */
void sro_gaussian_kernel(int n, double sigma, double* out);
/* blur_kernel_ = kernel_x * kernel_y.t() (blur_module.cpp:22). out is n*n. */
void sro_gaussian_psf(int n, double sigma, double* out);

/* cv::warpAffine(src, dst, [1 0 dx; 0 1 dy], size): INTER_LINEAR, BORDER_CONSTANT(0)
 * (reference call: motion_module.cpp:18-24). src and dst must not alias. */
void sro_warp_shift(const double* src, int H, int W, double dx, double dy, double* dst);
/* The fixed-point coordinate OpenCV derives for a shift d: returns n such that the source
 * coordinate of destination pixel p is p + n/32 (integer part n>>5, bilinear weight (n&31)/32). */
int sro_warp_quantize(double d);

/* cv::filter2D(src, dst, -1, kernel, Point(-1,-1), 0, BORDER_CONSTANT): correlation, centred
 * anchor, zero border (reference call: matrix_util.cpp:20-27).  kernel is kh x kw. */
void sro_filter2d(const double* src, int H, int W, const double* kernel, int kh, int kw, double* dst);

/* cv::resize(..., INTER_NEAREST) (reference call: image_data.cpp:341-347). */
void sro_resize_nearest(const double* src, int H, int W, double* dst, int H2, int W2);
/* The index map itself: src index used for destination index q (must be bit-exact). */
int sro_nearest_index(int q, int n_src, int n_dst);
/* ResizeAdditiveInterpolation (image_data.cpp:80-134): zero-insert upsample or sum-pool downsample. */
void sro_resize_additive(const double* src, int H, int W, double* dst, int H2, int W2);

/* ImageModel::ApplyToImage(ImageData*, k) for one channel: M_k, B, D in that order
 * (image_model.cpp:86-91).  Output size is int(H*(1/s)) x int(W*(1/s)) (image_data.cpp:353-364). */
void sro_forward(const sro_model* m, int k, const double* hr, int H, int W, double* lr_out);
/* ImageModel::ApplyTransposeToImage for one channel: D^T (zero insert), B^T (filter2D with
 * kernel.t()), M_k^T (warp by -shift) (image_model.cpp:93-101).  Input h x w, output (h*s) x (w*s). */
void sro_transpose(const sro_model* m, int k, const double* lr, int h, int w, double* hr_out);

/* ObjectiveDataTerm::Compute (objective_data_term.cpp:15-116).
 *   x        : C*H*W estimate (channels [channel_start, channel_end) of the image)
 *   obs_hr   : N images of C_total*H*W doubles, already nearest-upsampled to HR
 *              (map_solver.cpp:81-85), frame-major: obs_hr + k*C_total*H*W
 *   gradient : C*H*W, ACCUMULATED into (may be NULL)
 * returns the term's cost.  num_threads > 1 parallelises frames x channels. */
double sro_data_term(const sro_model* m, const double* x, int H, int W, int C,
                     const double* obs_hr, int C_total, int channel_start,
                     double* gradient, int num_threads);

/* Regularizers.  kind: 0 = TV, 1 = 3-D TV, 2 = BTV (btv_range, btv_decay).
 * sro_reg_apply      = Regularizer::ApplyToImage (tv_regularizer.cpp:110-132, btv_regularizer.cpp:67-90)
 * sro_reg_apply_diff = Regularizer::ApplyToImageWithDifferentiation (tv:134-227, btv:92-170)
 * Both write C*H*W outputs; partials is overwritten (starts from zero like the reference). */
void sro_reg_apply(int kind, int btv_range, double btv_decay, const double* x, int H, int W, int C,
                   double* values);
void sro_reg_apply_diff(int kind, int btv_range, double btv_decay, const double* x,
                        const double* constants, int H, int W, int C, double* values,
                        double* partials);

/* ObjectiveIRLSRegularizationTerm::Compute (objective_irls_regularization_term.cpp:10-58). */
double sro_irls_term(int kind, int btv_range, double btv_decay, double lambda, const double* weights,
                     const double* x, int H, int W, int C, double* gradient);
/* IRLS re-weighting (irls_map_solver.cpp:128-143): w = 1 / max(1e-5, reg(x)). */
void sro_reweight(int kind, int btv_range, double btv_decay, const double* x, int H, int W, int C,
                  double* weights);

/* ObjectiveFunction::ComputeAllTerms (objective_function.cpp:5-20) = zero g, data term, then
 * (if lambda > 0 and weights != NULL) one IRLS regularisation term. */
double sro_eval(const sro_model* m, const double* x, int H, int W, int C, const double* obs_hr,
                int reg_kind, int btv_range, double btv_decay, double lambda, const double* weights,
                double* gradient, int num_threads);

#ifdef __cplusplus
}
#endif
#endif  /* SR_ORACLE_H_ */
