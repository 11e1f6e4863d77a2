/*
 * mbavo.h — C-ABI of the B200-native blur-aware photometric tracking hot path (MBA-VO src/ba_tracker).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point names the reference
 * interface it replaces (paths relative to the reference tree, ethliup/MBA-VO @ 161e1af).  All functions return
 * 0 on success and a negative MBAVO_E* code otherwise, never throw and never exit; mbavo_last_error() returns a
 * thread-local description of the last failure.
 *
 * A context (mbavo_ctx) owns all device scratch, the mapped pinned result buffer and one CUDA stream of one logical
 * tracker on one GPU; it is not thread-safe.  Calls are blocking like the reference's
 * (results are valid on return) unless named *_async.
 *
 * Unknown ordering everywhere: [dt_0 .. dt_{n-1}, dw_0 .. dw_{n-1}] (merge_hessian_gradient_cost.cpp:52-62), rotations
 * updated on the right, R_j <- R_j Exp(dw_j) (Spline.h:317-330); quaternions are stored (x, y, z, w).
 */
#ifndef MBAVO_H_
#define MBAVO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBAVO_VERSION 100

#define MBAVO_OK 0
#define MBAVO_EINVAL (-1)      /* bad argument / unsupported configuration */
#define MBAVO_ECUDA (-2)       /* a CUDA runtime / driver call failed */
#define MBAVO_ERANGE (-3)      /* an exposure sample needs a control knot outside [0, n_knots) */
#define MBAVO_ECAPACITY (-4)   /* exceeds the limits the context was created with */
#define MBAVO_ENOTREADY (-5)   /* level / frame times not set */
#define MBAVO_ENCCL (-6)       /* sharded call: a peer rank did not arrive (collective failed) */

#define MBAVO_MAX_LEVELS 8     /* BlurAwareDirectTrackerOptions per-level arrays, blur_aware_direct_tracker.h:17-31 */
#define MBAVO_MAX_FRAMES 16
#define MBAVO_MAX_KNOT_WINDOW 8 /* control knots touched by one evaluation (k .. k + segments - 1) */

typedef struct mbavo_ctx mbavo_ctx;

/* Capacity of a context — initialize_shared_cuda_storages(max_num_frames, max_num_virtual_poses_per_frame,
 * max_num_keypoints, max_patch_size, max_num_ctrl_knots, spline_deg_k, storages), spline_update_step.cpp:9-58 */
typedef struct mbavo_limits
{
    int device;                          /* CUDA device ordinal, -1 = current */
    int max_num_frames;                  /* <= MBAVO_MAX_FRAMES */
    int max_num_virtual_poses_per_frame; /* exposure samples N per frame */
    int max_num_keypoints;               /* host-map points per level */
    int max_patch_size;                  /* residual-pattern pixels per point */
    int max_num_ctrl_knots;
} mbavo_limits;

#define MBAVO_MEM_HOST 0
#define MBAVO_MEM_DEVICE 1

/* One pyramid level of inputs — what BlurAwareDirectTracker::uploadDataToGpu(int) (blur_aware_direct_tracker.cpp:721-751)
 * and the image arguments of evaluate_cost_hessian_gradient (spline_update_step.h:62-68) provide.
 * With MBAVO_MEM_HOST the data is copied (borrowed for the call); with MBAVO_MEM_DEVICE the pointers are device
 * pointers that must stay valid until the level is replaced. */
typedef struct mbavo_level
{
    int mem;                       /* MBAVO_MEM_HOST or MBAVO_MEM_DEVICE, applies to every pointer below */
    int H, W;                      /* im_size_HW of this level */
    double fx, fy, cx, cy;         /* intrinsics of this level (already divided by 2^level, tracker.cpp:766-776) */
    const unsigned char *ref_I;    /* keyframe image, H*W row-major            (cuda_ref_img) */
    const float *ref_dIxy;         /* keyframe gradient, H*W*2 interleaved      (cuda_dIxy_ref) */
    const unsigned char *const *cur_I; /* n_frames live images, H*W each; the ARRAY is always in host memory   */
    int n_frames;
    const void *keypoint_xy;       /* num_keypoints records holding (x, y) as two doubles */
    int keypoint_xy_stride;        /* bytes between records: 16 for packed double2, 24 for Core::Vector2d */
    int keypoint_xy_offset;        /* byte offset of x inside a record: 0, or 8 for Core::Vector2d (Vector.h:12-16) */
    const double *keypoint_z;      /* num_keypoints depths                       (cuda_keypoint_depth_z) */
    int num_keypoints;
    const int *pattern_xy;         /* patch_size (dx, dy) pairs                  (cuda_local_patch_pattern_xy) */
    int patch_size;
    int num_virtual_poses;         /* exposure samples per frame at this level   (num_virtual_poses_per_frame[level]) */
    /* Optional, device memory, for callers that own the reference's storages (host/spline_update_step.h): */
    unsigned char *ext_outlier_flags; /* use this flag array (cuda_keypoints_outlier_flags) instead of the context's own;
                                         it is NOT cleared by mbavo_set_level */
    double *ext_patch_cost;        /* write patch costs here (cuda_patch_cost_gradient_hessian_tR) ... */
    int ext_patch_cost_stride;     /* ... with this stride in doubles ((6k+1)(6k+2)/2 in the reference) */
} mbavo_level;

/* Spline state of one evaluation — the spline arguments of evaluate_cost_hessian_gradient (spline_update_step.h:69-75)
 * plus the knot data the tracker memcpy's into the storages first (blur_aware_direct_tracker.cpp:755-763, 838-846). */
typedef struct mbavo_spline
{
    int spline_deg_k;              /* control knots per segment: 2 (linear) or 4 (cubic cumulative B-spline) */
    double start_time;             /* spline_start_time */
    double sample_dt;              /* spline_sample_dt (knot spacing) */
    int num_ctrl_knots;            /* n */
    const double *knots_t;         /* host, 3n */
    const double *knots_R;         /* host, 4n, quaternion (x, y, z, w) */
} mbavo_spline;

const char *mbavo_last_error(void);
int mbavo_version(void);

/* initialize_shared_cuda_storages / free_shared_cuda_storages (spline_update_step.cpp:9-95) */
int mbavo_create(const mbavo_limits *limits, mbavo_ctx **out);
int mbavo_destroy(mbavo_ctx *ctx);

/* Run all work of this context on `stream` (a cudaStream_t) instead of the context's own stream; NULL restores it.
 * Used to share a stream with a caller that owns device memory (e.g. a torch stream). */
int mbavo_set_stream(mbavo_ctx *ctx, void *stream);

/* BlurAwareDirectTracker::uploadDataToGpu() (blur_aware_direct_tracker.cpp:701-719): capture / exposure time per frame */
int mbavo_set_frame_times(mbavo_ctx *ctx, int n_frames, const double *cap_time, const double *exp_time);

/* BlurAwareDirectTracker::uploadDataToGpu(int pyra_level) (blur_aware_direct_tracker.cpp:721-751).  Also clears the
 * level's outlier flags (optimizePyramidLevel, :600-601). */
int mbavo_set_level(mbavo_ctx *ctx, int level, const mbavo_level *data);

/* ---- pyramids built on the device ----------------------------------------------------------
 * The step before the path: ImagePyramid<T>::computePyramid (src/core/measurements/ImagePyramid.h:59-99, 2x2 box, float
 * average, truncating cast), compute_image_gradients (src/core/image_proc/Gradient.h:17-75, 0.5 * central differences,
 * zero 1-pixel border) and the per-level cudaMalloc + H2D of Image::uploadToGpu (Image.h:125-136) that the tracker runs
 * on the CPU per keyframe (blur_aware_direct_tracker.cpp:346-353) and per frame (:112-116).  Only the level-0 images
 * cross PCIe; coarser levels, gradients and texels are produced on the GPU, bit-identical to the CPU loops.
 * Level l has H0 / 2^l x W0 / 2^l pixels.  A level is ready for mbavo_evaluate once its keyframe, live frame and points
 * are set; mbavo_set_level on the same level index replaces it. */
typedef struct mbavo_level_points
{
    int mem;                       /* MBAVO_MEM_HOST or MBAVO_MEM_DEVICE, applies to keypoint_xy / keypoint_z */
    double fx, fy, cx, cy;         /* intrinsics of this level (tracker.cpp:766-776) */
    const void *keypoint_xy;       /* as in mbavo_level */
    int keypoint_xy_stride, keypoint_xy_offset;
    const double *keypoint_z;
    int num_keypoints;
    const int *pattern_xy;
    int patch_size;
    int num_virtual_poses;
} mbavo_level_points;

int mbavo_set_keyframe_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *ref_I0, int H0, int W0);
int mbavo_set_live_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *const *cur_I0, int n_frames);
int mbavo_set_level_points(mbavo_ctx *ctx, int level, const mbavo_level_points *points);
/* the points of levels 0 .. n_levels-1 in one call (one synchronisation): points[l] describes level l */
int mbavo_set_points_pyramid(mbavo_ctx *ctx, int n_levels, const mbavo_level_points *points);
/* Everything a new frame brings, in ONE call with ONE synchronisation: the level-0 keyframe (ref_I0, may be NULL: the
 * keyframe pyramid already in the context stays) and live images (cur_I0, may be NULL likewise) go up on the context's
 * stream, followed by the pyramid / gradient / texel kernels, while the points of all levels go up on a second stream
 * underneath them.  Replaces the sequence mbavo_set_keyframe_pyramid + mbavo_set_live_pyramid + mbavo_set_points_pyramid
 * (three synchronisations) — what the tracker does per keyframe / frame in tmpProcessKeyframe + uploadDataToGpu
 * (blur_aware_direct_tracker.cpp:346-409, 701-751).  flags: MBAVO_UPLOAD_ASYNC returns without synchronising; the host
 * buffers must then stay valid and unchanged until the next blocking call on this context (any evaluation, sweep,
 * LM or statistics entry point) has returned. */
#define MBAVO_UPLOAD_ASYNC 1
int mbavo_set_frame(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *ref_I0, int H0, int W0,
                    const unsigned char *const *cur_I0, int n_frames, const mbavo_level_points *points, int flags);

/* A new live (blurred) frame for an already set level — what BlurAwareDirectTracker::trackFrame uploads per frame
 * (blur_aware_direct_tracker.cpp:112-116) while keyframe image, gradient, texels and host-map points stay resident.
 * cur_I: n_frames images of the level's H*W (host or device, as the level was set; the array itself is in host memory).
 * Clears the level's outlier flags like mbavo_set_level. */
int mbavo_set_live_images(mbavo_ctx *ctx, int level, int mem, const unsigned char *const *cur_I, int n_frames);

/* Outlier flags of a level: cuda_keypoints_outlier_flags + num_bad_keypoints (spline_update_step.h:25-26).
 * flags == NULL clears them.  flags is a host array of num_keypoints bytes (1 = outlier). */
int mbavo_set_outliers(mbavo_ctx *ctx, int level, const unsigned char *flags, int num_bad_keypoints);

/* num_bad_keypoints alone (CudaSharedStorages::num_bad_keypoints, read at spline_update_step.cpp:116) when the flag
 * array is owned by the caller (ext_outlier_flags). */
int mbavo_set_num_bad(mbavo_ctx *ctx, int level, int num_bad_keypoints);

/* evaluate_cost_hessian_gradient (spline_update_step.h:60-87, .cpp:97-349).
 *   total_cost : sum of Huber costs / num_residuals
 *   hessian    : 6n x 6n doubles (symmetric, so row- and column-major coincide) or NULL for the cost-only branch
 *   gradient   : 6n doubles, NULL iff hessian is NULL
 * Unlike the reference, each exposure sample's Jacobian goes to that sample's own spline segment, so an exposure
 * window may straddle control knots (up to MBAVO_MAX_KNOT_WINDOW knots touched per evaluation). */
int mbavo_evaluate(mbavo_ctx *ctx, int level, const mbavo_spline *spline, double huber_a, double *total_cost,
                   double *hessian, double *gradient);

/* Per-patch costs of the LAST evaluation of `level`: element 0 of every patch vector of
 * cuda_patch_cost_gradient_hessian_tR (read by detectOutliersAndUploadToGpu, blur_aware_direct_tracker.cpp:646-657).
 * out: host, n_frames * num_keypoints doubles. */
int mbavo_patch_costs(mbavo_ctx *ctx, int level, double *out);

/* detectOutliersAndUploadToGpu (blur_aware_direct_tracker.cpp:639-699) without the P*E D2H: statistics and flags are
 * computed on the device from the last evaluation's patch costs.  Flags are sticky; *num_bad_keypoints receives the
 * number of patches flagged by THIS call and becomes the level's num_bad_keypoints. */
int mbavo_detect_outliers(mbavo_ctx *ctx, int level, double max_chi_square_error, int *num_bad_keypoints);

/* ---- device-resident variants (multi-GPU plumbing) ----------------------------------------------------------- */

/* Length of the packed window vector [cost, g(6 NK), triu(H) row-major] for a knot window of NK knots */
int mbavo_packed_len(int knot_window);

/* Launch one evaluation without synchronising: the packed window vector of THIS context's points (already scaled by
 * 1 / num_residuals_global) is left in device memory at `packed_dev` (>= mbavo_packed_len(NK) doubles) on the
 * context's stream.  kmin / knot_window receive the window.  num_residuals_global <= 0 means "this context's own".
 * Summing the vectors of all shards (e.g. ncclAllReduce) and calling mbavo_unpack gives the global result. */
int mbavo_evaluate_async(mbavo_ctx *ctx, int level, const mbavo_spline *spline, double huber_a, int with_hessian,
                         long long num_residuals_global, double *packed_dev, int *kmin, int *knot_window);

/* merge_hessian_gradient_cost (merge_hessian_gradient_cost.cpp:8-87): scatter a packed window vector (host memory)
 * into the dense H (6n x 6n), g (6n) and cost.  hessian/gradient NULL => cost only. */
int mbavo_unpack(const double *packed_host, int kmin, int knot_window, int num_ctrl_knots, double *total_cost,
                 double *hessian, double *gradient);

/* ---- point sharding over the GPUs of one node ------------------------------------------------------
 * One process (or thread) and one context per GPU, each holding a contiguous block of the host-map points of every
 * level; images, spline and pattern are replicated.  After mbavo_shard_connect, mbavo_evaluate / mbavo_detect_outliers
 * (and everything built on them: mbavo_gn_iteration, mbavo_optimize_level) are COLLECTIVE calls that every rank makes
 * with the same arguments and that return the same global result on every rank: the last block of the tracking kernel
 * writes this rank's packed [cost, g, triu(H)] into every rank's mailbox over NVLink (peer-mapped memory), publishes a
 * sequence number, waits for the other ranks' vectors in its own mailbox and sums them in rank order — a one-shot
 * all-reduce fused into the kernel, no NCCL call, no extra launch.  The reference has no multi-GPU path; this replaces
 * what a ported version would do with ncclAllReduce after kernel_compute_frame_cost_gradient_hessian.
 *
 *   mbavo_shard_export   creates (first call) and ZEROES this context's mailbox; handle_out (MBAVO_IPC_HANDLE_BYTES, may be
 *                        NULL) receives its CUDA IPC handle for ranks in OTHER processes, mailbox_ptr_out (may be NULL) its
 *                        device pointer for ranks in the SAME process.  The mailbox is zeroed here — before its handle can
 *                        reach a peer — and never by mbavo_shard_connect, so no barrier is needed between the ranks'
 *                        connects and the first collective call: a rank that connects first may already deposit its first
 *                        vector in a slower peer's mailbox.  Call it after the last collective of an earlier connection
 *                        has returned.
 *   mbavo_shard_connect  handles: world x MBAVO_IPC_HANDLE_BYTES in rank order (other processes), or mailbox_ptrs: world
 *                        device pointers in rank order (same process); the entry of `rank` itself is ignored.  The
 *                        sequence numbers of the exchange restart at 0, so every connect needs a fresh mbavo_shard_export
 *                        of this context (MBAVO_ENOTREADY otherwise) and all ranks of a group (re)connect together.
 *   mbavo_shard_set_global_points   total number of host-map points of `level` over all ranks (the normaliser
 *                        1 / ((P - num_bad) F S) of spline_update_step.cpp:116-117 is global; num_bad_keypoints given to
 *                        mbavo_set_outliers / mbavo_set_num_bad is the global count too).
 * A rank that waits more than 4 s for a peer gives up: the call returns MBAVO_ENCCL. */
#define MBAVO_IPC_HANDLE_BYTES 64
#define MBAVO_MAX_SHARDS 8
int mbavo_shard_export(mbavo_ctx *ctx, void *handle_out, void **mailbox_ptr_out);
int mbavo_shard_connect(mbavo_ctx *ctx, int world, int rank, const void *handles, void *const *mailbox_ptrs);
int mbavo_shard_disconnect(mbavo_ctx *ctx);
int mbavo_shard_set_global_points(mbavo_ctx *ctx, int level, int num_keypoints_global);

/* ---- host-side solver, mirrors of the tracker's LM loop ------------------ */

#define MBAVO_SOLVER_SVD_JACOBI 0 /* solve_normal_equation.h:18-23 */
#define MBAVO_SOLVER_LDLT 1       /* solve_normal_equation.h:24-28 */

/* computeTrustRegionStep (blur_aware_direct_tracker.cpp:799-831): damps the diagonal of `hessian` IN PLACE by
 * (1 + 1/radius), step = -H^-1 g, *model_cost_change = -(g^T s + s^T H s / 2).  dim = 6n. */
int mbavo_trust_region_step(double *hessian, const double *gradient, int dim, double radius, int solver_type,
                            double *step, double *model_cost_change);

/* SplineSE3::Plus_t / Plus_R (src/core/common/Spline.h:307-330): candidate = knots [+] step */
int mbavo_spline_plus(int num_ctrl_knots, const double *knots_t, const double *knots_R, const double *step,
                      double *cand_t, double *cand_R);

/* One Gauss-Newton / LM iteration of optimizePyramidLevel (blur_aware_direct_tracker.cpp:609-637) without the
 * accept / reject bookkeeping: Hessian pass at the knots -> computeTrustRegionStep at `radius` -> Plus_t / Plus_R ->
 * cost-only pass at the candidate.  step_out (6n), cand_t (3n), cand_R (4n) may be NULL.  This is the benchmark's unit
 * of work ("GN iteration"). */
int mbavo_gn_iteration(mbavo_ctx *ctx, int level, int spline_deg_k, double start_time, double sample_dt,
                       int num_ctrl_knots, const double *knots_t, const double *knots_R, double radius, double huber_a,
                       int solver_type, double *cost, double *candidate_cost, double *step_out, double *cand_t,
                       double *cand_R);

/* One GN iteration (mbavo_gn_iteration) on every level from level_coarse down to level_fine — the coarse-to-fine loop of
 * optimizeTrajectory (blur_aware_direct_tracker.cpp:571-575) with a single iteration per level.  chain != 0: a level whose
 * candidate lowered the cost hands the candidate knots to the next finer level (knots_t / knots_R are updated in place);
 * chain == 0: every level starts from the given knots.  costs (may be NULL): (cost, candidate cost) per level, coarse first.
 * Runs device-resident: the solve, the candidate and the commit happen inside the kernels and the host waits once — as ONE
 * persistent launch for the whole sweep where mbavo_persistent_sweeps (below) says so, else one launch per pass; normal
 * equations that need the SVD branch, or a negative model decrease, make the call fall back to mbavo_gn_iteration per level. */
int mbavo_gn_sweep(mbavo_ctx *ctx, int level_coarse, int level_fine, int chain, int spline_deg_k, double start_time,
                   double sample_dt, int num_ctrl_knots, double *knots_t, double *knots_R, double radius, double huber_a,
                   int solver_type, double *costs);

typedef struct mbavo_lm_options
{
    int max_num_iterations;                 /* 50   blur_aware_direct_tracker.h:39 */
    double min_step_quality;                /* 0.5  :40 */
    double min_abs_cost_decrease;           /* 1e-3 :41 */
    int solver_type;                        /* MBAVO_SOLVER_SVD_JACOBI  :42 */
    int max_consecutive_nonmonotonic_steps; /* 5    :38 */
    double max_chi_square_error;            /* outlier threshold in sigmas :55 */
    double huber_a;                         /* huber_k :35 */
} mbavo_lm_options;

typedef struct mbavo_lm_summary
{
    int num_iterations;
    int num_accepted;
    int num_rejected;
    int num_invalid;
    int num_evaluations;     /* device evaluations issued (Hessian + cost-only) */
    int num_bad_keypoints;
    double initial_cost;
    double final_cost;
    double first_step[6 * 16]; /* first trust-region step (parity quantity) */
    char decisions[64];        /* 'A' accepted / 'R' rejected / 'I' invalid, NUL-terminated */
} mbavo_lm_summary;

void mbavo_lm_default_options(mbavo_lm_options *opt);

/* BlurAwareDirectTracker::optimizePyramidLevel (blur_aware_direct_tracker.cpp:590-637): the whole LM loop of one
 * level.  knots_t (3n) / knots_R (4n) are updated in place with the committed control knots. */
int mbavo_optimize_level(mbavo_ctx *ctx, int level, int spline_deg_k, double start_time, double sample_dt,
                         int num_ctrl_knots, double *knots_t, double *knots_R, const mbavo_lm_options *opt,
                         mbavo_lm_summary *summary);

/* ---- per-frame trajectory bookkeeping on the host -----------------------------------------------
 * Quaternions are (x, y, z, w); tangents are [translation(3), rotation(3)] as Core::Transformation::log / exp
 * (src/core/states/Transformation.cpp:164-178, wrappers of Sophus::SE3d::log / exp).
 * mbavo_spline_pose              SplineSE3::GetPose without Jacobians (src/core/common/Spline.h:222-290)
 * mbavo_spline_transform_by_right  SplineSE3::TransformByRight (Spline.h:212-219): t_i = R_i dt + t_i, R_i = R_i dR, in place
 * mbavo_spline_transform_to      SplineSE3::TransformTo(t, R, t) (Spline.h:184-201): the spline moved so that its pose at `time`
 *                                becomes the target; writes the new knots to knots_t / knots_R (3n / 4n)
 * mbavo_predict_spline           the constant-velocity prediction of trackFrame (blur_aware_direct_tracker.cpp:120-145):
 *                                TransformByRight(exp(velocity * dt_frame)), in place
 * mbavo_frame_velocity           log(T_prev^-1 T_cur) / dt_frame (:155-161) */
int mbavo_se3_exp(const double *tangent, double *t, double *q);
int mbavo_se3_log(const double *t, const double *q, double *tangent);
int mbavo_spline_pose(const mbavo_spline *spline, double time, double *t, double *q);
int mbavo_spline_transform_by_right(int num_ctrl_knots, double *knots_t, double *knots_R, const double *dq, const double *dt);
int mbavo_spline_transform_to(const mbavo_spline *spline, double time, const double *target_t, const double *target_q,
                              double *knots_t, double *knots_R);
int mbavo_predict_spline(int num_ctrl_knots, double *knots_t, double *knots_R, const double *velocity, double dt_frame);
int mbavo_frame_velocity(const double *prev_t, const double *prev_q, const double *cur_t, const double *cur_q, double dt_frame,
                         double *velocity);

/* ---- semi-dense host-map point selection -------------------------------------------------------
 * What BlurAwareDirectTracker::tmpProcessKeyframe does on the CPU for a new keyframe (blur_aware_direct_tracker.cpp:355-409),
 * for levels 0 .. n_levels-1 of the keyframe pyramid already built by mbavo_set_keyframe_pyramid:
 *   FeatureDetectorSemiDense::detect (src/core/feature_detectors/FeatureDetectorSemiDense.cpp:16-59): the pixels whose gradient
 *     magnitude (src/core/image_proc/Gradient.h:57-72) exceeds score_threshold;
 *   FeatureDetectorBase::gridSelection (FeatureDetectorBase.cpp:49-92): per cell of (int)(cell / 1.414^l) pixels the strongest
 *     of them (the first in row-major order on ties), in cell order;
 *   the depth look-up (:389-409): z = depth_z[(int)(y 2^l + .5)][(int)(x 2^l + .5)], points with z < 1e-2 dropped.
 * Bit-exact with the reference.  The selected points become the level's points (as mbavo_set_level_points would set them,
 * intrinsics divided by 2^l as in :765-771) and never leave HBM; num_selected[l] returns how many there are (a level with
 * none cannot be evaluated).  In a sharded context every rank selects the same points and keeps its contiguous block.
 * mbavo_get_points copies a level's points to the host for a caller that wants them (xy: 2 doubles per point; either
 * pointer may be NULL; both NULL just returns the count). */
typedef struct mbavo_point_selection
{
    float score_threshold;         /* 25      (blur_aware_direct_tracker.cpp:357) */
    int cell_H, cell_W;            /* 30, 30  (:358-359); must be positive */
    int depth_mem;                 /* MBAVO_MEM_HOST or MBAVO_MEM_DEVICE */
    const float *depth_z;          /* level-0 depth along the optical axis, H0 x W0 floats, row-major */
    double fx, fy, cx, cy;         /* level-0 intrinsics */
    const int *pattern_xy;         /* residual pattern (host memory), as in mbavo_level */
    int patch_size;
    int num_virtual_poses;
} mbavo_point_selection;
int mbavo_select_points(mbavo_ctx *ctx, int n_levels, const mbavo_point_selection *selection, int *num_selected);
int mbavo_get_points(mbavo_ctx *ctx, int level, int capacity, double *xy, double *z, int *num_points);

/* ---- keyframe test statistics --------------------------------------------------------------
 * BlurAwareDirectTracker::isKeyframe (blur_aware_direct_tracker.cpp:205-248): over the host-map points of `level` (the
 * tracker uses level 0), avg_flow = sqrtf(mean |pi(T0^-1 P) - p|^2) and avg_kernel_len = sqrtf(mean |pi(T-^-1 P) -
 * pi(T+^-1 P)|^2), P the point un-projected with its depth, T the live-camera-to-keyframe pose at the capture time (T0)
 * and at -/+ half the exposure (T-, T+), pi the pinhole projection.  poses_tq: 3 x 7 doubles (tx ty tz qx qy qz qw) in
 * the order T0, T-, T+ — SplineSE3::GetPose at those times.  Collective in a sharded context (means over all ranks).
 * mbavo_is_keyframe applies the thresholds (keyframe_max_flow_mag0 / _mag1, keyframe_max_blur_kernel_mag, :250-262). */
int mbavo_keyframe_stats(mbavo_ctx *ctx, int level, const double *poses_tq, double *avg_flow, double *avg_kernel_len);

/* ---- BlurAwareDirectTracker::trackFrame ------------------------------------------------------------
 * The per-frame driver of the reference (blur_aware_direct_tracker.cpp:88-203), composed from the entry points above.
 * mbavo_tracker is the state the reference keeps between frames: the two-knot spline (:100-105, spacing = options.dt_frame),
 * mNeighFrameVelocity, mTprevB2W, mPrevTimestamp, mTKeyframe.
 *   mbavo_tracker_init          first frame (:91-109): identity spline starting at the keyframe's capture time
 *   mbavo_track_frame           one blurred frame (:112-162, 200-203): uploads its pyramid (level 0 only travels), predicts the
 *                               spline at constant velocity, runs the LM loop of every level coarse to fine, returns the pose
 *                               at the capture time (relative to the keyframe and composed with the keyframe's pose), the
 *                               isKeyframe statistics and the per-level LM summaries; updates velocity / previous pose.
 *                               The context must hold the keyframe pyramid and the points (mbavo_set_keyframe_pyramid +
 *                               mbavo_select_points or mbavo_set_points_pyramid).
 *   mbavo_tracker_new_keyframe  the re-anchoring the reference does when the frame became the keyframe (:186-199); the
 *                               caller (mbavo_is_keyframe decides) uploads the new keyframe + points. */
#define MBAVO_MAX_TRACKER_KNOTS 16
typedef struct mbavo_tracker
{
    int spline_deg_k, num_ctrl_knots;
    double sample_dt, start_time;
    double knots_t[3 * MBAVO_MAX_TRACKER_KNOTS], knots_R[4 * MBAVO_MAX_TRACKER_KNOTS];
    double velocity[6];                  /* mNeighFrameVelocity: [translation, rotation] per second */
    double prev_t[3], prev_q[4];         /* mTprevB2W */
    double prev_timestamp;               /* mPrevTimestamp */
    double keyframe_t[3], keyframe_q[4]; /* mTKeyframe */
} mbavo_tracker;

typedef struct mbavo_frame_result
{
    double t_cur2key[3], q_cur2key[4];       /* spline pose at the capture time */
    double t_cur2world[3], q_cur2world[4];   /* mTKeyframe * that (the return value of trackFrame) */
    double avg_flow, avg_kernel_len;         /* isKeyframe statistics */
    int levels_run;                          /* bit l: level l was optimised */
    mbavo_lm_summary levels[MBAVO_MAX_LEVELS];
} mbavo_frame_result;

int mbavo_tracker_init(mbavo_tracker *tracker, int spline_deg_k, double sample_dt, double keyframe_capture_time);
int mbavo_track_frame(mbavo_ctx *ctx, mbavo_tracker *tracker, int n_levels, int mem, const unsigned char *cur_I0,
                      double capture_time, double exposure_time, const mbavo_lm_options *opt, mbavo_frame_result *result);
int mbavo_tracker_new_keyframe(mbavo_tracker *tracker, double capture_time);
/* the decision of isKeyframe (blur_aware_direct_tracker.cpp:251-262) on the statistics mbavo_track_frame returned, with the
 * options keyframe_max_flow_mag0 / _mag1 / keyframe_max_blur_kernel_mag (blur_aware_direct_tracker.h): 1 = new keyframe */
int mbavo_is_keyframe(double avg_flow, double avg_kernel_len, double max_flow_mag0, double max_flow_mag1,
                      double max_blur_kernel_mag);

/* ---- synthetic blurred frame ---------------------------------------------------------------
 * warp_image + synthesize_motion_blurred_img (src/ba_tracker/generate_synthetic_data.cpp:127-180): the mean over num_poses
 * plane-induced warps of the keyframe (plane Z = plane_depth), each truncated to 8 bits, the mean rounded to nearest even.
 * poses_tq: num_poses x 7 doubles (tx ty tz qx qy qz qw) — what SplineSE3::GetPose returns for every sample time (:166-169);
 * at most 256 poses.  ref_I / out: H*W bytes, both host or both device (`mem`).  device < 0: the current device.
 * Bit-exact with the reference arithmetic (the unit is compiled without FMA contraction). */
int mbavo_synthesize_blurred(int device, int mem, const unsigned char *ref_I, int H, int W, double plane_depth, double fx,
                             double fy, double cx, double cy, const double *poses_tq, int num_poses, unsigned char *out);

/* ---- per-stage intermediates of one Hessian-pass evaluation (SURVEY.md §8d stage gates) ---------------------------------
 * What the reference's module test prints stage by stage (test/test_blur_aware_tracker_modules.cpp:183-342 virtual poses and
 * their Jacobians, :344-500 local patches, :502-895 pixel residuals and Jacobians), produced by the PRODUCT kernels with their
 * debug stores switched on.  All pointers are host memory, each may be NULL.  F frames, N exposure samples, P points, S patch
 * pixels, k = spline_deg_k, NK = knot_window:
 *   poses_tq        [F N 7]     pose of every exposure sample, t then q (x, y, z, w), fp64 — before the rounding into the fp32
 *                               sample record (compute_virtual_camera_poses.cu:26-109)
 *   blend_weights   [F N k]     translation blend weight of the k knots of the sample's segment (J_t = w_j I, SplineFunctor.h:30-40, 74-91)
 *   theta           [F N k 9]   Theta_j (3 x 3 row-major) = d theta / d w_j, theta the right perturbation of the pose rotation; the
 *                               reference's 4 x 3 block is dq / dw_j = L(q) [I / 2; 0] Theta_j (SplineFunctor.h:178-213, 274-361)
 *   segment_start_knot [F N]    first control knot of the sample's segment
 *   patch_centres   [F P 2]     compute_local_patches_xy.cu:26-49, fp64
 *   residuals       [F P S]     raw residual of every pixel, (1/N) sum_i I_i - I_cur (…cost.cu:115-121); 0 for invalid pixels
 *   jacobians       [F P S 6NK] raw Jacobian row of every pixel w.r.t. the knots [kmin, kmin + NK), ordered [t-block | w-block]
 *                               (…cost.cu:123-151 scattered per sample segment)
 * Available for k = 2 with windows of 2 or 3 knots and k = 4 with 4 (MBAVO_ECAPACITY otherwise); unsharded contexts. */
typedef struct mbavo_debug_out
{
    double *poses_tq, *blend_weights, *theta;
    int *segment_start_knot;
    double *patch_centres;
    float *residuals, *jacobians;
    int kmin, knot_window, spline_deg_k; /* out */
    double cost;                         /* out: total cost of the evaluation */
} mbavo_debug_out;
int mbavo_debug_dump(mbavo_ctx *ctx, int level, const mbavo_spline *spline, double huber_a, mbavo_debug_out *out);

/* ---- experiment: TMA-staged cost pass (profiles/r2_tma_variant.md) --------------------------------------------------------
 * The cost-only evaluation of `level` through the variant of the cost pass that stages, per host-map point, a box_w x box_h tile
 * of the 8-bit keyframe in shared memory with cp.async.bulk.tensor.2d (TMA, double-buffered on mbarriers) and takes the bilinear
 * taps from the tile (global fallback outside it) — the mechanism BASELINE's north star names, kept next to the product kernel
 * (which gathers packed texels through L1) so that the two can be measured against each other.  Same result as the cost-only
 * branch of mbavo_evaluate.  kernel_ms: CUDA-event time of the kernel; tile_fraction: share of the valid samples served from
 * the tiles.  box_w a multiple of 16 (<= 256), box_w * box_h a multiple of 128, image width a multiple of 16. */
int mbavo_debug_cost_tma(mbavo_ctx *ctx, int level, const mbavo_spline *spline, double huber_a, int box_w, int box_h,
                         double *total_cost, float *kernel_ms, double *tile_fraction);

/* ---- introspection for benchmarks ---------------------------------------------------------------------------- */

/* Number of kernels this library has launched on behalf of ctx since creation */
long long mbavo_kernel_launches(const mbavo_ctx *ctx);

/* Number of mbavo_gn_sweep calls that ran on the device-resident path (solve, candidate and commit inside the kernels,
 * one host wait) rather than evaluation by evaluation. */
long long mbavo_device_sweeps(const mbavo_ctx *ctx);
/* ... of which as ONE persistent launch: the pose kernel for the starting knots + sweep_kernel, which runs the Hessian pass,
 * the solve, the candidate's sample records and the cost pass of every level inside one resident grid of one block per SM
 * (passes separated by a ticket + flag barrier).  Taken when all levels share the exposure-sample count and the patch size,
 * track one frame, have keyframe texels, and the knot window is one of {k=2: 2..5, k=4: 4}; MBAVO_NO_PERSISTENT=1 disables it. */
long long mbavo_persistent_sweeps(const mbavo_ctx *ctx);
/* Durations (microseconds, device globaltimer) of the passes of the last persistent sweep, coarse level first: [2 li] the
 * Hessian pass of level li — from the release of the previous pass to its own release, i.e. record load, batches, reduction,
 * solve, candidate and the candidate's pose records included — and [2 li + 1] its cost pass.  The stamps are two stores per
 * pass by one thread; they are always taken. */
int mbavo_sweep_pass_times(mbavo_ctx *ctx, double *us_out, int capacity, int *num_passes);

/* 1 if the kernels of `level` read the keyframe through the packed texels built by mbavo_set_level / mbavo_set_frame (16 bytes
 * per pixel holding the 4 x 4 byte neighbourhood, from which a sample's four intensities and four gradients come as byte
 * differences; possible when the gradient image is reproduced exactly by such differences — always the case for Gradient.h's
 * 0.5 * central differences of an 8-bit image), 0 if they gather ref_I / ref_dIxy directly, -1 if the level is not set.
 * Results are identical either way. */
int mbavo_level_uses_texels(const mbavo_ctx *ctx, int level);

/* Device time in milliseconds of the tracking kernel of the last mbavo_evaluate* call on `level` (CUDA events
 * recorded around that kernel on the context's stream); < 0 if timing is disabled. */
int mbavo_enable_kernel_timing(mbavo_ctx *ctx, int enable);
float mbavo_last_kernel_ms(mbavo_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MBAVO_H_ */
