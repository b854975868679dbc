/*
 * srb200.h -- C ABI of the B200-native MAP super-resolution gradient engine (libsrb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of rteammco/super-resolution: one evaluation
 * of cost + gradient of the MAP objective.  The reference has no FFI layer; its seams are C++
 * virtual interfaces.  Each entry point below names the reference interface it stands behind
 * (paths relative to the reference root); INTEGRATION.md shows the adapter subclasses a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns an srb_status (0 = SRB_OK) and never
 *     throws; srb_last_error() gives the message of the last failure on a context.  The
 *     reference aborts through glog CHECK on the same conditions (e.g. objective_data_term.cpp:
 *     91-95, image_model.cpp:66-67); adapters CHECK on the status.
 *   - images are planar row-major fp64, index c*H*W + row*W + col (src/util/util.cpp:81-89),
 *     exactly the layout ALGLIB's real_1d_array holds (irls_map_solver.cpp:232-239).
 *   - "host" pointers are caller-owned CPU memory; "dev" pointers are CUDA device memory on the
 *     context's device.  Host buffers may be pinned with srb_pin_host for full PCIe rate.
 *   - a context is bound to one CUDA device and one stream; it is not thread-safe (the reference
 *     calls every interface from one thread).
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device and fails
 *     with SRB_ERR_CUDA otherwise.
 */
#ifndef SRB200_H_
#define SRB200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct srb_ctx srb_ctx;

typedef enum {
  SRB_OK = 0,
  SRB_ERR_INVALID = 1,   /* bad argument (the reference would CHECK-fail) */
  SRB_ERR_CUDA = 2,      /* CUDA runtime error / no device */
  SRB_ERR_GEOMETRY = 3,  /* (size, scale) for which cv::resize's index map is not block-regular */
  SRB_ERR_STATE = 4,     /* call order (e.g. evaluation before srb_set_observations) */
  SRB_ERR_NOMEM = 5
} srb_status;

/* Regularizer kinds (src/optimization/tv_regularizer.h, btv_regularizer.h). */
enum { SRB_REG_NONE = -1, SRB_REG_TV = 0, SRB_REG_TV3D = 1, SRB_REG_BTV = 2 };

/* Kernel path selection.  AUTO takes the fused tile kernel whenever the model qualifies and the
 * reference-order kernels otherwise; both run on the GPU. */
enum { SRB_PATH_AUTO = 0, SRB_PATH_REFERENCE_ORDER = 1, SRB_PATH_FUSED = 2 };

/*
 * The image formation model A_k = D * B * M_k, what ImageModel::CreateImageModel builds
 * (src/image_model/image_model.cpp:17-61; ImageModelParameters image_model.h:26-44) plus the
 * shape of the observations MapSolver is constructed with (src/optimization/map_solver.cpp:52-86:
 * HR size = LR size * scale).
 */
typedef struct {
  int lr_height, lr_width; /* size of every low-resolution observation */
  int num_channels;        /* channels per observation */
  int num_frames;          /* observations held by THIS context (a rank's frame shard) */
  int scale;               /* downsampling scale s >= 1 (DownsamplingModule) */
  int psf_size;            /* K, odd: blur_kernel_ is K x K; 0 = no BlurModule */
  const double* psf;       /* K*K row-major correlation kernel (blur_module.cpp:20-22), host */
  const double* shifts;    /* 2*num_frames doubles dx_0,dy_0,dx_1,... (motion_shift.h:14-18);
                              NULL = no MotionModule */
} srb_model_desc;

/* Library version string. */
const char* srb_version(void);
/* Number of CUDA devices visible (0 if none / no driver). */
int srb_device_count(void);

/* ---- planning (host only: works without a CUDA device) ---------------------------------------
 * What srb_create would decide for this model: whether the fused tile kernel covers it (and why
 * not), how the (frame, tap) entries spread over the sub-pixel phases, and the band of "special" LR
 * samples near the image border that are evaluated in the reference's operation order instead
 * (regular samples: band_lo_r <= row < band_hi_r and band_lo_c <= column < band_hi_c). */
typedef struct {
  int hr_height, hr_width;
  int warps_uniform, warps_integer;
  int fused;                 /* 1: the fused tile kernel covers this model */
  int fractional;            /* bilinear forward / transpose taps */
  int psf_half;
  int num_entries, min_entries_per_phase, max_entries_per_phase;
  int band_lo_r, band_hi_r, band_lo_c, band_hi_c;
  int table_driven;          /* interior tiles use the precomputed residual table: frames per
                                sub-pixel phase it is specialised for (1, 2 or 4), 0 = generic pass */
  int zlayout;               /* Z layout (k_tile_zt): 1 = qualifies (integer shifts, 3x3 .. 9x9 PSF, one
                                frame on every sub-pixel phase), 2 = qualifies with empty phases (at most one
                                frame per phase: frame shards, cfg2), 0 = no.  Both are the DEFAULT kernel for
                                such models; SRB_ZLAYOUT=0 in the environment of srb_create disables it */
  int zt_frames;             /* transposed Z layout (k_tile_zt): frames per non-empty sub-pixel phase, all with
                                the same shift and averaged at upload (1: cfg1-3 and frame shards, 2: cfg4,
                                4: cfg5); 0 = the model does not qualify (fractional shifts, PSF outside
                                3x3 .. 9x9, different shifts on one phase, unequal frame counts) */
  char why[160];             /* reason when fused == 0, or the validation error */
} srb_plan_info;
srb_status srb_plan(const srb_model_desc* desc, srb_plan_info* out);
/* cv::warpAffine's fixed-point translation (motion_module.cpp:18-24): the source coordinate of
 * destination pixel p is p + n/32, n = srb_quantize_shift(d). */
int srb_quantize_shift(double shift);
/* One dimension of the special-sample test: LR sample q of an image of hr_size pixels. */
int srb_sample_is_special(int q, int hr_size, int psf_half, int scale, double shift);

/* ---- life cycle ------------------------------------------------------------------------- */
/* Builds a context on CUDA device `device`.  Replaces: ImageModel::CreateImageModel +
 * MapSolver::MapSolver (map_solver.cpp:52-86).  Fails with SRB_ERR_GEOMETRY when cv::resize
 * (INTER_NEAREST) would not map LR<->HR block-regularly for this size/scale. */
srb_status srb_create(const srb_model_desc* desc, int device, srb_ctx** out);
/* Same for one rank's frame shard of a multi-GPU run: num_frames may be 0 (more devices than frames); such a
 * context contributes only its row band of the regularization term (srb_set_regularizer_rows). */
srb_status srb_create_shard(const srb_model_desc* desc, int device, srb_ctx** out);
void srb_destroy(srb_ctx* ctx);
const char* srb_last_error(const srb_ctx* ctx);

/* Uploads the observations once, at LR resolution: lr is [num_frames][num_channels][h][w].
 * Replaces the observation copies MapSolver keeps (map_solver.cpp:81-85; the nearest-neighbour
 * upsampling done there is a pure index map and is folded into the kernels). */
srb_status srb_set_observations(srb_ctx* ctx, const double* lr_host);
/* Same, from device memory (layout identical). */
srb_status srb_set_observations_dev(srb_ctx* ctx, const double* lr_dev);

/* Channel sub-range [c0, c1) the following evaluations work on -- IRLSMapSolver's
 * split_channels (irls_map_solver.cpp:200-206; ObjectiveDataTerm channel_start/channel_end,
 * objective_data_term.h:24-31).  Default: all channels.  Resets the IRLS weights to 1. */
srb_status srb_set_channel_range(srb_ctx* ctx, int c0, int c1);

/* Regularizer + regularization parameter: MapSolver::AddRegularizer (map_solver.h:92-93) with a
 * TotalVariationRegularizer (SetUse3dTotalVariation for SRB_REG_TV3D) or a
 * BilateralTotalVariationRegularizer(scale_range, spatial_decay) (btv_regularizer.cpp:48-65).
 * kind = SRB_REG_NONE or lambda <= 0 removes it.  Resets the IRLS weights to 1
 * (irls_map_solver.cpp:66-74). */
srb_status srb_set_regularizer(srb_ctx* ctx, int kind, double lambda, int btv_range,
                               double btv_decay);

/* IRLS weights for the active channel range, (c1-c0)*H*W doubles; NULL = all ones.  This is the
 * upload done where the reference constructs ObjectiveIRLSRegularizationTerm per outer iteration
 * (irls_map_solver.cpp:83-93). */
srb_status srb_set_irls_weights(srb_ctx* ctx, const double* weights_host);
/* IRLS re-weighting on device (irls_map_solver.cpp:128-143): w = 1 / max(1e-5, reg(x)).
 * x_host = NULL re-uses the estimate the last HOST-buffer evaluation or solve on the current channel range
 * left in the context (srb_eval, srb_data_term, srb_irls_term, srb_cg_minimize, srb_lbfgs_minimize,
 * srb_solve_irls); after anything else -- the *_dev and srb_peer_* forms work on the caller's own device
 * buffer -- it fails with SRB_ERR_STATE.  weights_out_host may be NULL. */
srb_status srb_reweight(srb_ctx* ctx, const double* x_host, double* weights_out_host);
/* Same for a device-resident estimate ((c1-c0)*H*W doubles on the context's device), stream-ordered, no
 * host synchronisation: what follows srb_cg_minimize_dev in a device-resident IRLS loop. */
srb_status srb_reweight_dev(srb_ctx* ctx, const double* x_dev);

/* ---- device-resident solver ------------------------------------
 * The reference minimises with ALGLIB's mincg on host arrays (RunCGSolverAnalyticalDiff,
 * alglib_objective.cpp:47-75), so every evaluation moves x and g across PCIe.  These entry points
 * run the same algorithm (csrc/srb_cg.h: a restatement of mincgiteration / mcsrch / mcstep, pinned
 * bit for bit against the reference's ALGLIB on the CPU) with all solver vectors in HBM; only
 * scalars reach the host.  Thresholds as in MapSolverOptions (map_solver.h:25-60) after
 * AdjustThresholdsAdaptively; all zero selects ALGLIB's automatic EpsX = 1e-6. */
typedef struct srb_cg_options {
  double gradient_norm_threshold;        /* mincgsetcond EpsG */
  double cost_decrease_threshold;        /* EpsF */
  double parameter_variation_threshold;  /* EpsX */
  int max_num_solver_iterations;         /* MaxIts, 0 = unlimited */
  int num_lbfgs_hessian_corrections;     /* srb_solve_irls: 0 = CG_SOLVER (the reference's default), m > 0 =
                                            LBFGS_SOLVER with m pairs (map_solver.h:20-23, :51);
                                            srb_lbfgs_minimize: m, 1..64; srb_cg_minimize ignores it */
} srb_cg_options;
typedef struct srb_cg_report {   /* alglib::mincgreport + the cost RunCGSolverAnalyticalDiff returns */
  int iterations;
  int num_evaluations;
  int termination_type;          /* 1 EpsF, 2 EpsX, 4 EpsG, 5 MaxIts, 7 repeated restarts, -8 inf / nan */
  int num_restarts;
  double final_cost;
} srb_cg_report;
typedef struct srb_irls_report {
  int num_irls_iterations;
  int num_solver_iterations;     /* summed over the outer iterations */
  int num_evaluations;
  int last_termination_type;
  double final_cost;
} srb_irls_report;
/* RunCGSolverAnalyticalDiff for the active channel range with the current regularizer and IRLS
 * weights: x (n = (c1-c0)*H*W doubles) is the initial estimate on entry and the solution on return.
 * options = NULL: all thresholds zero. */
srb_status srb_cg_minimize(srb_ctx* ctx, double* x_host_inout, const srb_cg_options* options,
                           srb_cg_report* report);
srb_status srb_cg_minimize_dev(srb_ctx* ctx, double* x_dev_inout, const srb_cg_options* options,
                               srb_cg_report* report);
/* RunLBFGSSolverAnalyticalDiff (alglib_objective.cpp:111-140): ALGLIB's minlbfgs
 * (optimization.cpp:21640-22330) restated the same way (csrc/srb_cg.h: lbfgs_minimize, bit-identical to
 * ALGLIB on the CPU); termination type -2 = rounding errors prevent further progress. */
srb_status srb_lbfgs_minimize(srb_ctx* ctx, double* x_host_inout, const srb_cg_options* options,
                              srb_cg_report* report);
srb_status srb_lbfgs_minimize_dev(srb_ctx* ctx, double* x_dev_inout, const srb_cg_options* options,
                                  srb_cg_report* report);
/* IRLSMapSolver::RunIRLSLoop (irls_map_solver.cpp:45-157) for the active channel range: weights reset
 * to 1, then { CG solve; w = 1 / max(1e-5, reg(x)) } until the cost of two consecutive outer
 * iterations differs by less than irls_cost_difference_threshold or max_num_irls_iterations (0 =
 * unlimited) is reached; a single CG solve when no regularizer is configured. */
srb_status srb_solve_irls(srb_ctx* ctx, double* x_host_inout, const srb_cg_options* options,
                          int max_num_irls_iterations, double irls_cost_difference_threshold,
                          srb_irls_report* report);

/* Which kernel path evaluations take (default SRB_PATH_AUTO); srb_active_path reports the one
 * the current configuration resolves to. */
srb_status srb_set_path(srb_ctx* ctx, int path);
int srb_active_path(const srb_ctx* ctx);
/* Parity device for the reference-order kernels (SRB_PATH_REFERENCE_ORDER, srb_data_term, srb_irls_term): sum
 * the data cost and the regularization cost sequentially in the reference's own order
 * (objective_data_term.cpp:36-50 and :104-114, objective_irls_regularization_term.cpp:44-55) instead of by a
 * parallel tree, so that the cost -- like the gradient -- is BIT-IDENTICAL to the CPU reference's and a whole
 * ALGLIB solve on top of it reproduces the CPU solve bit for bit.  One thread per frame: for small problems
 * (cfg1, cfg2); off by default. */
srb_status srb_set_strict_cost(srb_ctx* ctx, int on);
/* 1 when the fused path evaluates its tiles from the observations re-laid out on the HR grid ("Z layout":
 * default for models with integer shifts and at most one frame per sub-pixel phase; SRB_ZLAYOUT=0 in the
 * environment of srb_create disables it), else 0.  The layout is a second, padded copy of the observations
 * ([Ct][W+2*KH rounded to tiles][H+halo] doubles: the size of x for N = s^2) built on the device by
 * srb_set_observations.  No reference counterpart. */
int srb_zlayout_active(const srb_ctx* ctx);
/* Multi-GPU frame sharding: this context holds the frames of one rank.  The data
 * term covers the context's frames; the regularizer term is computed only for HR rows
 * [row_begin, row_end) so that the sum over ranks is the full objective.  Default: all rows. */
srb_status srb_set_regularizer_rows(srb_ctx* ctx, int row_begin, int row_end);

/* ---- the hot path ----------------------------------------------------------------------- */
/* ObjectiveFunction::ComputeAllTerms (objective_function.cpp:5-20), the body of
 * AlglibObjectiveFunction (alglib_objective.cpp:142-152): cost = data term + IRLS regularization
 * term, gradient OVERWRITTEN with the full sum (gradient_host may be NULL: cost only).
 * x and gradient: (c1-c0)*H*W doubles on the host.  On the fused path the call is pipelined: x is
 * copied in slices, each slice's gradient rows return while the next slices are computed, so H2D,
 * kernel and D2H overlap (pin the buffers with srb_pin_host for full effect; SRB_PIPE_CHUNKS sets
 * the slice count, default 16, 1 = serial). */
srb_status srb_eval(srb_ctx* ctx, const double* x_host, double* gradient_host, double* cost);
/* Same with device-resident x / gradient (no PCIe traffic except the cost scalar). */
srb_status srb_eval_dev(srb_ctx* ctx, const double* x_dev, double* gradient_dev, double* cost);
/* Multi-GPU form: writes this rank's partial gradient to gradient_cost_dev[0..n) and its partial
 * cost to gradient_cost_dev[n] (n = (c1-c0)*H*W), all on the context's stream, without any host
 * synchronisation -- ready for ONE allreduce(sum) over n+1 doubles. */
srb_status srb_eval_partial_dev(srb_ctx* ctx, const double* x_dev, double* gradient_cost_dev);

/* Pipelined multi-GPU form.  The active gradient is cut, in memory order, into "units" of
 * rows_per_unit HR rows of one channel (the last unit of a channel may be shorter).
 * srb_eval_units_dev evaluates units [unit_begin, unit_end): it writes exactly those gradient rows
 * of gradient_cost_dev (a contiguous range, see srb_unit_range) so the caller can start the
 * allreduce of that slice on another stream while the next units are computed.  After the last
 * unit, srb_eval_finish_dev adds what is not tiled (border-band samples, non-fused regularizers)
 * and writes the rank's partial cost to gradient_cost_dev[n].  Needs the fused path
 * (srb_active_path == SRB_PATH_FUSED); srb_num_units reports 1 unit when pipelining is not
 * possible for the current configuration (then [0,1) is the whole evaluation). */
srb_status srb_num_units(srb_ctx* ctx, int* num_units, int* rows_per_unit);
/* Element range [*begin, *end) of the gradient covered by units [unit_begin, unit_end). */
srb_status srb_unit_range(srb_ctx* ctx, int unit_begin, int unit_end, unsigned long long* begin,
                          unsigned long long* end);
srb_status srb_eval_units_dev(srb_ctx* ctx, const double* x_dev, double* gradient_cost_dev,
                              int unit_begin, int unit_end);
srb_status srb_eval_finish_dev(srb_ctx* ctx, const double* x_dev, double* gradient_cost_dev);

/* Row-band partition with one process per GPU (the alternative to frame sharding): every
 * rank's context holds EVERY frame; a rank evaluates units [unit_begin, unit_end) of the whole objective -- the
 * gradient rows it writes are final, no cross-rank sum of the gradient exists -- and *cost_dev (device, may be
 * NULL) receives the cost of exactly those units, ready for a scalar allreduce.  x_dev must be current on the
 * rank's rows plus srb_halo_rows() rows either side (what the stencils reach: the PSF twice, + 1 for TV or R for BTV,
 * + twice the largest shift for models with a border band), which neighbouring ranks exchange
 * (sharding.RowBandObjective).  Units are (channel, 32-row tile) in memory order, tile_rows = ceil(H / 32) per channel.
 * Needs the fused path and a regularizer it covers (SRB_ERR_STATE otherwise). */
srb_status srb_eval_unit_range_dev(srb_ctx* ctx, const double* x_dev, double* gradient_dev, int unit_begin,
                                   int unit_end, double* cost_dev);
int srb_halo_rows(const srb_ctx* ctx);

/* ---- multi-GPU peer path: reduce-scatter / all-gather over NVLink peer memory ------------------
 * One process per GPU.  Every rank owns a contiguous band of gradient units.  The tile kernel
 * evaluates the gradient band by band; a finished band that belongs to another rank is pushed by a
 * copy engine into slot [rank] of its owner (a CUDA-IPC peer mapping over NVLink) while the SMs
 * compute the next band, so the reduce-scatter traffic hides behind the math.  After a flag barrier
 * in peer memory the owner sums its band in fixed rank order and stores the result into the
 * gradient buffer of every rank (srb_peer_gather_dev), which is the all-gather half.
 * Buffers are allocated with srb_dev_alloc (plain cudaMalloc, exportable), exchanged as 64-byte IPC
 * handles by the host layer (sharding.py uses torch.distributed for that), opened with
 * srb_ipc_open and registered with srb_peer_setup.  Needs the fused path without border band and
 * with a fused or absent regularizer (srb_num_units > 1); otherwise SRB_ERR_STATE. */
srb_status srb_peer_sizes(srb_ctx* ctx, int world, unsigned long long* slots_bytes,
                          unsigned long long* out_bytes);
srb_status srb_dev_alloc(void** dev_ptr, unsigned long long bytes);
srb_status srb_dev_free(void* dev_ptr);
srb_status srb_ipc_export(const void* dev_ptr, unsigned char handle[64]);
srb_status srb_ipc_open(const unsigned char handle[64], void** dev_ptr);
srb_status srb_ipc_close(void* dev_ptr);
/* slot_bases[o] / out_bases[o]: slot array / gradient buffer of rank o (local pointer for o == rank).
 * out buffers hold n gradient doubles, the total cost at [n], and `world` partial-cost slots. */
srb_status srb_peer_setup(srb_ctx* ctx, int rank, int world, double* const* slot_bases,
                          double* const* out_bases);
/* Phase 1: evaluate this rank's partial objective band by band, pushing every finished band to
 * its owner, then post the partial cost and raise this rank's "scattered" flag on every rank. */
srb_status srb_peer_scatter_dev(srb_ctx* ctx, const double* x_dev);
/* Phase 2: wait for all ranks' flags, sum this rank's band over the slots and store it, with the
 * total cost, into every rank's gradient buffer; wait until all bands have arrived here. */
srb_status srb_peer_gather_dev(srb_ctx* ctx);
/* The flag barriers spin a bounded number of times.  When a rank never arrives they record the failure on the
 * device, skip the sum and the stores (nothing partial is published; the other ranks time out in turn) and
 * return.  srb_peer_status, srb_synchronize and srb_memcpy_d2h read that record: SRB_ERR_STATE means the
 * gradient and cost of the evaluation are not valid (the record is cleared by the call). */
srb_status srb_peer_status(srb_ctx* ctx);
srb_status srb_memcpy_d2h(srb_ctx* ctx, void* dst_host, const void* src_dev, unsigned long long bytes);

/* ---- single-process multi-GPU form ------------------------------------------------------------------
 * The reference's solver is one process and one thread (IRLSMapSolver::Solve, irls_map_solver.cpp:192-265;
 * AlglibObjectiveFunction, alglib_objective.cpp:142-152): srb_multi_* lets that thread drive up to 8 GPUs
 * through the same call shape as srb_eval.  desc describes the WHOLE model (all frames); the frames are
 * sharded in contiguous blocks over the devices (objective_data_term.cpp:104-114 is the loop being split),
 * x and the IRLS weights are replicated, the regularization term is split by HR row bands.  Per evaluation
 * device r copies only its 1/G band of x from the host over its own PCIe link, the bands are all-gathered
 * over NVLink peer memory, every device evaluates its frames band by band while copy engines push finished
 * bands to their owners (the reduce-scatter), and device r returns its summed band of the gradient to the
 * host: host traffic per PCIe link is 1/G of srb_eval's.  devices = NULL means devices 0 .. n_gpus-1.  Pin
 * the host buffers with srb_pin_host.  Results equal srb_eval's up to fp64 re-association of the
 * cross-device sum (fixed device order: deterministic).
 *
 * Partition.  SRB_PARTITION_FRAMES is the one described above (the contract partition of SURVEY 8e: the
 * LR-frame axis).  Its exchange moves the whole gradient (C*H*W doubles) over NVLink per evaluation whatever the
 * number of devices, and since the fused kernels apply the PSF once per evaluation, not once per frame, a
 * device's kernel time barely drops with fewer frames.  SRB_PARTITION_ROWS cuts the HR image instead: every
 * device holds every frame and evaluates the WHOLE objective on its HR row bands; device r fetches only its
 * bands of x (plus the few halo rows the PSF and regularizer stencils reach) from the host and returns its
 * bands of the gradient, which are final -- no exchange between the devices at all, and compute, H2D and D2H
 * per device all drop by n_gpus.  Costs n_gpus copies of the observations in HBM (cfg3 100 MB, cfg5 1.6 GB
 * per device).  Applies when the fused tile kernel covers the model, with a regularizer other than 3-D TV (which
 * couples the channels); otherwise device 0 evaluates alone.  srb_multi_create takes the partition
 * from SRB_MULTI_PARTITION=frames|rows in the environment (default frames).
 *
 * Threads.  The entry points are called from one thread and are not re-entrant per context.  For the row-band
 * evaluation and the multi-device solver below the library keeps n_gpus - 1 helper threads per context (made on
 * first use, joined by srb_multi_destroy) that issue the other devices' runtime calls beside the calling thread;
 * they sleep between calls and spin only while a srb_multi_*_minimize / srb_multi_solve_irls call is running.
 * SRB_MULTI_THREADS=0 in the environment keeps everything on the calling thread. */
enum { SRB_PARTITION_FRAMES = 0, SRB_PARTITION_ROWS = 1 };
typedef struct srb_multi srb_multi;
srb_status srb_multi_create(const srb_model_desc* desc, int n_gpus, const int* devices, srb_multi** out);
srb_status srb_multi_create_partitioned(const srb_model_desc* desc, int n_gpus, const int* devices, int partition,
                                        srb_multi** out);
void srb_multi_destroy(srb_multi* m);
const char* srb_multi_last_error(const srb_multi* m);
int srb_multi_num_gpus(const srb_multi* m);
srb_ctx* srb_multi_rank_ctx(srb_multi* m, int rank);            /* the per-device context (diagnostics) */
srb_status srb_multi_set_observations(srb_multi* m, const double* lr_host);  /* [num_frames][C][h][w] */
srb_status srb_multi_set_channel_range(srb_multi* m, int c0, int c1);
srb_status srb_multi_set_regularizer(srb_multi* m, int kind, double lambda, int btv_range, double btv_decay);
srb_status srb_multi_set_irls_weights(srb_multi* m, const double* weights_host);
srb_status srb_multi_reweight(srb_multi* m, const double* x_host, double* weights_out_host);
srb_status srb_multi_set_path(srb_multi* m, int path);
/* ObjectiveFunction::ComputeAllTerms (objective_function.cpp:5-20) on all devices; gradient_host may be NULL. */
srb_status srb_multi_eval(srb_multi* m, const double* x_host, double* gradient_host, double* cost);
/* The device-resident solver (srb_cg_minimize / srb_lbfgs_minimize / srb_solve_irls above: RunCGSolverAnalyticalDiff,
 * RunLBFGSSolverAnalyticalDiff, alglib_objective.cpp:47-140; IRLSMapSolver::RunIRLSLoop, irls_map_solver.cpp:45-157)
 * on all devices of a SRB_PARTITION_ROWS context, still one host thread.  The (channel, tile row) units of the
 * active range are cut into n_gpus contiguous bands; a band of units is a contiguous range of every solver vector,
 * and device r keeps only that range of x, g, d, ...: every vector pass and every evaluation costs 1/n_gpus per
 * device, the gradient is never exchanged, and per line-search step only the halo rows of the trial point
 * (srb_halo_rows() rows each way, pulled from the neighbouring devices over NVLink by copy engines) and eight
 * scalars per device (summed on the host in fixed device order: deterministic) cross the devices.  x enters and
 * leaves over n_gpus PCIe links at once.  Same iterates as the single-device solver up to the re-association of
 * the reductions.  A frame-sharded context (n_gpus > 1) is refused with SRB_ERR_STATE; a configuration the row
 * bands do not cover (3-D TV, a model outside the fused tile kernel) is solved by device 0 alone.
 * SRB_MULTI_SHARE_DEVICES=1 in the environment of srb_multi_create lets `devices` name the same GPU more than once
 * (every entry still gets its own context and streams): the whole multi-device logic then runs on one GPU. */
srb_status srb_multi_cg_minimize(srb_multi* m, double* x_host_inout, const srb_cg_options* options,
                                 srb_cg_report* report);
srb_status srb_multi_lbfgs_minimize(srb_multi* m, double* x_host_inout, const srb_cg_options* options,
                                    srb_cg_report* report);
/* The IRLS weights are internal to the loop (a local vector of RunIRLSLoop, irls_map_solver.cpp:66-74): they
 * start at 1 and are reset to 1 on return (every device has re-weighted only the rows it owns). */
srb_status srb_multi_solve_irls(srb_multi* m, double* x_host_inout, const srb_cg_options* options,
                                int max_num_irls_iterations, double irls_cost_difference_threshold,
                                srb_irls_report* report);

/* ObjectiveDataTerm::Compute (objective_data_term.cpp:98-116): returns the data cost and ADDS the
 * data gradient into gradient_host (may be NULL). */
srb_status srb_data_term(srb_ctx* ctx, const double* x_host, double* gradient_host_accum,
                         double* cost);
/* ObjectiveIRLSRegularizationTerm::Compute (objective_irls_regularization_term.cpp:10-58):
 * cost = sum lambda*w*r^2, gradient ADDED into gradient_host (may be NULL). */
srb_status srb_irls_term(srb_ctx* ctx, const double* x_host, double* gradient_host_accum,
                         double* cost);

/* Regularizer::ApplyToImage (regularizer.h:26-28; tv_regularizer.cpp:110-132,
 * btv_regularizer.cpp:67-90) for the configured regularizer: values_out[num_channels*H*W]. */
srb_status srb_reg_apply(srb_ctx* ctx, const double* x_host, int num_channels,
                         double* values_out);
/* Regularizer::ApplyToImageWithDifferentiation (regularizer.h:41-45; tv_regularizer.cpp:134-227,
 * btv_regularizer.cpp:92-170): values and partial derivatives of sum_j c_j r_j^2. */
srb_status srb_reg_apply_diff(srb_ctx* ctx, const double* x_host, const double* constants_host,
                              int num_channels, double* values_out, double* partials_out);

/* ImageModel::ApplyToImage(ImageData*, index) for one channel (image_model.cpp:86-91): M_k, B, D
 * applied to an H x W image (any size); lr_out must hold int(H*(1/s)) * int(W*(1/s)) doubles
 * (image_data.cpp:353-364); the decimation index map is cv::resize's, bit-exact. */
srb_status srb_forward(srb_ctx* ctx, int frame, const double* hr_host, int H, int W,
                       double* lr_out_host);
/* ImageModel::ApplyToImage for EVERY frame of the model and every channel of one HR image of the
 * solver's size (image_model.cpp:76-84 looped over frames, what generate_data.cpp:83-127 and
 * super_resolution.cpp:286-311 do to synthesise an LR stack, without the additive noise):
 * hr is [num_channels][H][W], lr_out is [num_frames][num_channels][h][w]; one kernel launch. */
srb_status srb_forward_all(srb_ctx* ctx, const double* hr_host, double* lr_out_host);
/* ImageModel::ApplyTransposeToImage for one channel (image_model.cpp:93-101): D^T (zero insert),
 * B^T (correlation with blur_kernel_.t()), M_k^T (warp by the negated shift); input h x w,
 * output (h*s) x (w*s). */
srb_status srb_transpose(srb_ctx* ctx, int frame, const double* lr_host, int h, int w,
                         double* hr_out_host);

/* ---- the steps either side of the hot path (N2 data generation, N3 initial estimate +
 * scores, N4 hyperspectral front end) -- on the device, behind the same boundary ----------------------------- */

/* ImageData::ResizeImage(size, INTERPOLATE_LINEAR) (image_data.cpp:310-350: cv::resize INTER_LINEAR per channel)
 * for a planar image src [C][h][w] -> dst [C][H][W].  OpenCV's sampling geometry (half-pixel centres, clamped at the
 * borders), rows interpolated horizontally then vertically, every product and sum rounded separately.  OpenCV's
 * own build evaluates its coefficient tables partly in float; results agree with cv2 to 1 ulp for power-of-two
 * scale factors and to ~3e-7 otherwise (tests/test_frontend_oracle.py pins the oracle against cv2 fixtures). */
srb_status srb_resize_linear(srb_ctx* ctx, const double* src_host, int num_channels, int h, int w, int H, int W,
                             double* dst_host);
/* The solver's initial estimate (super_resolution.cpp:368-373): LR observation `frame` (0 in the reference) of
 * the active channel range, upsampled bilinearly to the HR size; needs srb_set_observations.  The _dev form
 * writes (c1-c0)*H*W doubles on the device, stream-ordered: srb_cg_minimize_dev can follow directly. */
srb_status srb_initial_estimate(srb_ctx* ctx, int frame, double* x_host_out);
srb_status srb_initial_estimate_dev(srb_ctx* ctx, int frame, double* x_dev_out);
/* PeakSignalToNoiseRatioEvaluator::Evaluate (peak_signal_to_noise_ratio.cpp:11-54; peak value 1.0, +inf for
 * identical images) and StructuralSimilarityEvaluator::Evaluate (structural_similarity.cpp:9-103: the GLOBAL
 * mean / variance / covariance form; the reference's defaults are k1 = 0.01, k2 = 0.03, image_scale = 1) of
 * `image` against `truth`, n = channels * pixels doubles each (both must have the same size: the reference's
 * resize branch resizes an image to its own size, i.e. does nothing).  psnr / ssim may be NULL. */
srb_status srb_score(srb_ctx* ctx, const double* image_host, const double* truth_host, unsigned long long n,
                     double k1, double k2, double image_scale, double* psnr, double* ssim);
srb_status srb_score_dev(srb_ctx* ctx, const double* image_dev, const double* truth_dev, unsigned long long n,
                         double k1, double k2, double image_scale, double* psnr, double* ssim);

/* AdditiveNoiseModule::ApplyToImage (additive_noise_module.cpp:19-36): data[i] += N(0, (sigma / 255)^2).  The
 * reference draws from cv::randn on OpenCV's process-global generator, which the program never seeds: its noise
 * has no values to reproduce, only a distribution.  Here the samples come from Philox4x32-10 (counter = sample
 * index / 4 and stream_id, key = seed) through Box-Muller: deterministic for a (seed, stream_id) pair and
 * independent of the launch geometry.  sigma must be positive (additive_noise_module.cpp:15-17). */
srb_status srb_add_noise(srb_ctx* ctx, double* data_host_inout, unsigned long long n, double sigma,
                         unsigned long long seed, unsigned long long stream_id);
srb_status srb_add_noise_dev(srb_ctx* ctx, double* data_dev_inout, unsigned long long n, double sigma,
                             unsigned long long seed, unsigned long long stream_id);
/* ImageModel::ApplyToImage(image, k) for every frame k with the noise module last (image_model.cpp:76-84, what
 * generate_data.cpp:118-127 and super_resolution.cpp:286-311 loop over): the LR stack [num_frames][C][h][w] of the
 * HR image [C][H][W] given on the host OR on the device (the other pointer NULL); noise_sigma = 0 leaves the
 * noise module out.  lr_out_host may be NULL; keep_as_observations != 0 makes the stack the context's
 * observations without a round trip through the host (srb_set_observations_dev is implied). */
srb_status srb_generate_observations(srb_ctx* ctx, const double* hr_host, const double* hr_dev, double noise_sigma,
                                     unsigned long long seed, double* lr_out_host, int keep_as_observations);

/* ENVI header / HSI configuration (HSIBinaryDataParameters, hyperspectral_data_loader.h; ReadHeaderFromFile,
 * hyperspectral_data_loader.cpp:226-270 -- "samples" is stored as the row count and "lines" as the column count,
 * as the reference does). */
typedef struct {
  int interleave_bsq;   /* 1 = band sequential (the only supported interleave) */
  int data_type;        /* ENVI data type code; 4 = float32 (the only supported type) */
  int big_endian;       /* "byte order = 1" */
  int header_offset;    /* bytes before the first sample (ENVI convention).  Every file the reference ships or
                           writes has 0; for other values the reference's reader scales the offset by the
                           sample size for the first run and then drops it (:85-101), which is not reproduced */
  int num_data_rows, num_data_cols, num_data_bands;
} srb_envi_header;
srb_status srb_envi_read_header(const char* header_path, srb_envi_header* out);
/* ReadBinaryFileBSQ<float> (hyperspectral_data_loader.cpp:68-118): rows [r0, r1), columns [c0, c1), bands
 * [b0, b1) of a float32 BSQ file as planar doubles [b1-b0][r1-r0][c1-c0].  The selected rows are read in one
 * run per band into pinned memory and copied to the device while the next band is read; byte swap, column
 * crop and float -> double conversion run on the device.  _dev leaves the image in device memory. */
srb_status srb_envi_read(srb_ctx* ctx, const char* data_path, const srb_envi_header* header, int r0, int r1, int c0,
                         int c1, int b0, int b1, double* image_out_host);
srb_status srb_envi_read_dev(srb_ctx* ctx, const char* data_path, const srb_envi_header* header, int r0, int r1,
                             int c0, int c1, int b0, int b1, double* image_out_dev);
/* WriteBinaryFileBSQ<float> (hyperspectral_data_loader.cpp:120-196): float32 BSQ in machine byte order plus
 * `path`.hdr and `path`.config with the reference's keys.  Host only. */
srb_status srb_envi_write(const char* data_path, const double* image_host, int num_bands, int num_rows, int num_cols);

/* SpectralPCA (spectral_pca.cpp:155-199): cv::PCA (data as rows) trained on 10 * num_bands pixel vectors
 * sub-sampled from the images at a fixed stride (GetPCAInputData, :27-96); images_host[i] is [num_bands][num_pixels].
 * num_pca_bands > 0 keeps that many components; num_pca_bands = 0 applies cv::PCA's retained-variance rule: the
 * leading components up to, not including, the one whose cumulative eigenvalue share first exceeds
 * `retained_variance`, and at least 2.  Training is a num_bands x num_bands host problem (mean,
 * covariance, cyclic Jacobi); eigenvectors are defined up to sign, as with cv::PCA.  Host only. */
typedef struct srb_pca srb_pca;
srb_status srb_pca_create(const double* const* images_host, int num_images, int num_bands, unsigned long long num_pixels,
                          int num_pca_bands, double retained_variance, srb_pca** out);
void srb_pca_destroy(srb_pca* pca);
int srb_pca_num_components(const srb_pca* pca);
int srb_pca_num_bands(const srb_pca* pca);
/* mean [num_bands], eigenvectors [num_components][num_bands] (rows), eigenvalues [num_components]; any may be NULL */
srb_status srb_pca_get(const srb_pca* pca, double* mean_out, double* eigenvectors_out, double* eigenvalues_out);
/* SpectralPCA::GetPCAImage / ReconstructImage (spectral_pca.cpp:98-153, 175-197: pca.project / pca.backProject per
 * pixel): [num_bands][P] -> [num_components][P] and back, one thread per pixel on the device. */
srb_status srb_pca_project(srb_ctx* ctx, const srb_pca* pca, const double* image_host, unsigned long long num_pixels,
                           double* pca_image_out_host);
srb_status srb_pca_reconstruct(srb_ctx* ctx, const srb_pca* pca, const double* pca_image_host,
                               unsigned long long num_pixels, double* image_out_host);

/* ---- plumbing --------------------------------------------------------------------------- */
/* Page-locks a caller buffer (e.g. ALGLIB's x / g arrays) so H2D/D2H run at full PCIe rate. */
srb_status srb_pin_host(void* ptr, unsigned long long bytes);
srb_status srb_unpin_host(void* ptr);
/* The context's CUDA stream (a cudaStream_t) and device scratch the host layer may use. */
void* srb_stream(srb_ctx* ctx);
double* srb_dev_x(srb_ctx* ctx);         /* (c1-c0)*H*W estimate buffer */
double* srb_dev_gradient(srb_ctx* ctx);  /* (c1-c0)*H*W + 1 gradient (+cost) buffer */
srb_status srb_synchronize(srb_ctx* ctx);

/* Records CUDA events around the dominant kernel of every evaluation (the fused tile kernel) so
 * that srb_get_timing can report its device time; off by default. */
srb_status srb_set_profiling(srb_ctx* ctx, int on);

typedef struct {
  double last_eval_kernel_ms;   /* device time of the kernels of the last evaluation */
  double last_eval_h2d_ms, last_eval_d2h_ms;
  unsigned long long num_evals; /* evaluations so far */
  unsigned long long kernel_launches; /* CUDA kernels launched by this context so far */
  unsigned long long algorithmic_bytes_per_eval; /* 8*C*P*(2|3 + N/s^2) */
  double last_main_kernel_ms;   /* device time of the last fused tile kernel launch (profiling on) */
} srb_timing;
srb_status srb_get_timing(srb_ctx* ctx, srb_timing* out);
/* Summed over the devices of a multi-GPU context; last_eval_kernel_ms = device time of the last srb_multi_eval
 * (first H2D to last D2H on device 0's clock). */
srb_status srb_multi_get_timing(srb_multi* m, srb_timing* out);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H_ */
