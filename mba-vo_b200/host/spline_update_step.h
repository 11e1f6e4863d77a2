// Drop-in replacement of the reference's evaluation driver, src/ba_tracker/spline_update_step.h:18-87 — same names,
// same argument meaning, same storage members the tracker pokes with raw cudaMemcpy
// (blur_aware_direct_tracker.cpp:600-601, 646-647, 696-698, 711-718, 729-750, 755-763, 838-846) — implemented on top of
// the C-ABI in include/mbavo.h.  A maintainer replaces spline_update_step.{h,cpp} (and drops the five compute_*.cu /
// merge_hessian_gradient_cost.cpp files) with this pair and links libmbavo_b200.so; see INTEGRATION.md.
//
// Differences that are visible to a caller:
//   * scratch that only the reference kernels used (cuda_pixel_jacobians_tR, cuda_vir_pixel_to_ctrl_knots_tR, the
//     J_virtual_pose / temp_* arrays, ...) is no longer allocated; the members stay (nullptr) so code that merely
//     frees or ignores them keeps compiling;
//   * cuda_patch_cost_gradient_hessian_tR keeps its stride ((6k+1)(6k+2)/2 doubles per patch) but only element 0 of
//     every patch vector — the patch cost, the only element the tracker reads (:500, :647) — is written.
#ifndef MBAVO_SPLINE_UPDATE_STEP_SHIM_H_
#define MBAVO_SPLINE_UPDATE_STEP_SHIM_H_

#include "../../include/mbavo.h"

// Minimal stand-ins for core/common/Vector.h and CustomType.h when the reference headers are not on the include path
// (same layout: leading int nDim, Vector.h:12-16).  With the reference tree present, include its headers first.
#ifndef CORE_COMMON_VECTOR_H_
#define CORE_COMMON_VECTOR_H_
namespace SLAM
{
    namespace Core
    {
        template <class T, int nDim_>
        struct VectorX
        {
            int nDim;
            T values[nDim_];
        };
        struct Vector2d : public VectorX<double, 2>
        {
            Vector2d() { this->nDim = 2; }
            Vector2d(double x, double y)
            {
                this->nDim = 2;
                values[0] = x;
                values[1] = y;
            }
            double &operator()(int i) { return values[i]; }
            double operator()(int i) const { return values[i]; }
        };
    } // namespace Core
} // namespace SLAM
#endif
#ifndef SLAM_CORE_CUSTOM_TYPE_H
#define SLAM_CORE_CUSTOM_TYPE_H
namespace SLAM
{
    typedef double FLOAT;
}
#endif

namespace SLAM
{
    namespace VO
    {
        struct CudaSharedStorages
        {
            // written by the tracker with cudaMemcpy / cudaMemset
            double *cuda_img_cap_time = nullptr;
            double *cuda_img_exp_time = nullptr;
            double *cuda_keypoint_depth_z = nullptr;
            Core::Vector2d *cuda_keypoint_xy = nullptr;
            unsigned char *cuda_keypoints_outlier_flags = nullptr;
            int num_bad_keypoints = 0;

            unsigned char **cuda_cur_images = nullptr;
            int *cuda_local_patch_pattern_xy = nullptr;

            double *cuda_spline_ctrl_knots_data_t = nullptr;
            double *cuda_spline_ctrl_knots_data_R = nullptr;

            // reference-kernel scratch: kept for source compatibility, never allocated
            double *cuda_sampled_virtual_poses = nullptr;
            double *cuda_J_virtual_pose_t_to_knots_t = nullptr;
            double *cuda_J_virtual_pose_R_to_knots_R = nullptr;
            double *cuda_jacobian_log_exp = nullptr;
            double *cuda_temp_X_4x4 = nullptr;
            double *cuda_temp_Y_4x4 = nullptr;
            double *cuda_temp_Z_4x4 = nullptr;
            Core::Vector2d *cuda_local_patches_XY = nullptr;
            double *cuda_pixel_residuals = nullptr;
            double *cuda_pixel_jacobians_tR = nullptr;
            FLOAT *cuda_vir_pixel_to_ctrl_knots_tR = nullptr;
            FLOAT *cuda_vir_pixel_residual = nullptr;

            // read back by the tracker: element [patch * E] is the patch cost
            double *cuda_patch_cost_gradient_hessian_tR = nullptr;
            double *cuda_frame_cost_gradient_hessian_tR = nullptr;

            // new: the context that owns the fused-kernel scratch, and the capacities it was created with
            mbavo_ctx *mbavo = nullptr;
            int mbavo_max_num_frames = 0;
            int mbavo_max_num_ctrl_knots = 0;
            // new: the small arrays above (capture / exposure times, live-image pointers, control knots) live in ONE device
            // slab so that an evaluation reads back what the tracker uploaded with a single copy
            unsigned char *mbavo_slab = nullptr;
            int mbavo_slab_bytes = 0;
            // new, opt-in: 1 = keep the level of the previous evaluation (keyframe texels included) while the image /
            // keypoint / pattern pointers, the sizes and mbavo_keyframe_epoch are unchanged, instead of re-deriving it on
            // every call.  A caller that rewrites the keyframe image or its gradient behind an unchanged pointer must bump
            // mbavo_keyframe_epoch (tmpProcessKeyframe is the one place, tracker.cpp:346-409).  0 (default) re-derives always.
            int mbavo_texel_cache = 0;
            int mbavo_keyframe_epoch = 0;
            struct LevelKey *mbavo_level_key = nullptr; // state of that cache (owned by the storages)
        };

        // spline_update_step.h:60-68 of the reference: sizes the context once from the tracker's maxima
        void initialize_shared_cuda_storages(const int max_num_frames, const int max_num_virtual_poses_per_frame, const int max_num_keypoints,
                                             const int max_patch_size, const int max_num_ctrl_knots, const int spline_deg_k,
                                             CudaSharedStorages &storages);

        void free_shared_cuda_storages(CudaSharedStorages &storages);

        // spline_update_step.h:70-87 of the reference.  cpu_hessian_tR == nullptr selects the cost-only branch (.cpp:242-348);
        // H is 6n x 6n symmetric, g is 6n ordered [t-block, omega-block] (merge_hessian_gradient_cost.cpp:41-85)
        void evaluate_cost_hessian_gradient(const int n_vir_poses_per_frame, const int n_frames, const unsigned char *cuda_ref_img,
                                            const float *cuda_dIxy_ref, const int num_keypoints, const int patch_size,
                                            const Core::VectorX<double, 4> &intrinsics, const Core::VectorX<int, 2> &im_size_HW,
                                            const int spline_deg_k, const double spline_start_time, const double spline_sample_dt,
                                            const int *cpu_ctrl_knot_start_indices, const int num_ctrl_knots,
                                            const CudaSharedStorages &storages, const double huber_a, double *total_costs,
                                            double *cpu_hessian_tR, double *cpu_gradient_tR);
    } // namespace VO
} // namespace SLAM

#endif
