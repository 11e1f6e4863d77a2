// Drop-in evaluation driver: the reference API of src/ba_tracker/spline_update_step.{h,cpp} on top of the C-ABI.
#include "spline_update_step.h"

#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace SLAM
{
    namespace VO
    {
        // what the level of the previous evaluation was derived from (CudaSharedStorages::mbavo_texel_cache)
        struct LevelKey
        {
            bool valid = false;
            const void *ref_I = nullptr, *dIxy = nullptr, *cur[MBAVO_MAX_FRAMES] = {};
            int H = 0, W = 0, P = 0, S = 0, N = 0, F = 0, k = 0, epoch = 0;
            double K[4] = {};
        };

        namespace
        {
            constexpr size_t align16(size_t n) { return (n + 15) & ~size_t(15); }
        } // namespace

        // spline_update_step.cpp:9-58.  Only the buffers the tracker itself reads or writes are allocated.
        void initialize_shared_cuda_storages(const int max_num_frames,
                                             const int max_num_virtual_poses_per_frame,
                                             const int max_num_keypoints,
                                             const int max_patch_size,
                                             const int max_num_ctrl_knots,
                                             const int spline_deg_k,
                                             CudaSharedStorages &storages)
        {
            const int num_patches = max_num_frames * max_num_keypoints;
            // capture / exposure times, live-image pointers and control knots share one slab (one read-back per evaluation)
            const size_t o_cap = 0, o_exp = align16(o_cap + sizeof(double) * max_num_frames),
                         o_cur = align16(o_exp + sizeof(double) * max_num_frames), o_kt = align16(o_cur + sizeof(void *) * max_num_frames),
                         o_kR = align16(o_kt + sizeof(double) * max_num_ctrl_knots * 3), total = align16(o_kR + sizeof(double) * max_num_ctrl_knots * 4);
            cudaMalloc((void **)&storages.mbavo_slab, total);
            cudaMemset(storages.mbavo_slab, 0, total);
            storages.mbavo_slab_bytes = (int)total;
            storages.cuda_img_cap_time = reinterpret_cast<double *>(storages.mbavo_slab + o_cap);
            storages.cuda_img_exp_time = reinterpret_cast<double *>(storages.mbavo_slab + o_exp);
            storages.cuda_cur_images = reinterpret_cast<unsigned char **>(storages.mbavo_slab + o_cur);
            storages.cuda_spline_ctrl_knots_data_t = reinterpret_cast<double *>(storages.mbavo_slab + o_kt);
            storages.cuda_spline_ctrl_knots_data_R = reinterpret_cast<double *>(storages.mbavo_slab + o_kR);
            storages.mbavo_level_key = new LevelKey();
            cudaMalloc((void **)&storages.cuda_keypoint_depth_z, sizeof(double) * max_num_keypoints);
            cudaMalloc((void **)&storages.cuda_local_patch_pattern_xy, sizeof(int) * max_patch_size * 2);
            cudaMalloc((void **)&storages.cuda_keypoint_xy, sizeof(Core::Vector2d) * max_num_keypoints);
            cudaMalloc((void **)&storages.cuda_keypoints_outlier_flags, sizeof(unsigned char) * max_num_keypoints);
            cudaMemset(storages.cuda_keypoints_outlier_flags, 0, sizeof(unsigned char) * max_num_keypoints);

            int nelems = spline_deg_k * 6 + 1;
            nelems = (1 + nelems) * nelems / 2;
            cudaMalloc((void **)&storages.cuda_patch_cost_gradient_hessian_tR, sizeof(double) * num_patches * nelems);
            cudaMalloc((void **)&storages.cuda_frame_cost_gradient_hessian_tR, sizeof(double) * max_num_frames * nelems);

            mbavo_limits lim{};
            lim.device = -1;
            lim.max_num_frames = max_num_frames;
            lim.max_num_virtual_poses_per_frame = max_num_virtual_poses_per_frame;
            lim.max_num_keypoints = max_num_keypoints;
            lim.max_patch_size = max_patch_size;
            lim.max_num_ctrl_knots = max_num_ctrl_knots;
            storages.mbavo_max_num_frames = max_num_frames;
            storages.mbavo_max_num_ctrl_knots = max_num_ctrl_knots;
            if (mbavo_create(&lim, &storages.mbavo) != MBAVO_OK)
            {
                std::fprintf(stderr, "mbavo: initialize_shared_cuda_storages: %s\n", mbavo_last_error());
                storages.mbavo = nullptr;
            }
        }

        // spline_update_step.cpp:60-95
        void free_shared_cuda_storages(CudaSharedStorages &storages)
        {
            mbavo_destroy(storages.mbavo);
            storages.mbavo = nullptr;
            cudaFree(storages.mbavo_slab); // cap / exposure times, live-image pointers, control knots
            storages.mbavo_slab = nullptr;
            storages.cuda_img_cap_time = storages.cuda_img_exp_time = nullptr;
            storages.cuda_cur_images = nullptr;
            storages.cuda_spline_ctrl_knots_data_t = storages.cuda_spline_ctrl_knots_data_R = nullptr;
            delete storages.mbavo_level_key;
            storages.mbavo_level_key = nullptr;
            cudaFree(storages.cuda_keypoint_depth_z);
            cudaFree(storages.cuda_local_patch_pattern_xy);
            cudaFree(storages.cuda_keypoint_xy);
            cudaFree(storages.cuda_keypoints_outlier_flags);
            cudaFree(storages.cuda_patch_cost_gradient_hessian_tR);
            cudaFree(storages.cuda_frame_cost_gradient_hessian_tR);
        }

        // spline_update_step.cpp:97-349.  cpu_ctrl_knot_start_indices is accepted for source compatibility; the segment of
        // every exposure sample is derived from its own time (the reference's per-frame index is the special case of an
        // exposure window that stays inside one segment).
        void evaluate_cost_hessian_gradient(const int n_vir_poses_per_frame,
                                            const int n_frames,
                                            const unsigned char *cuda_ref_img,
                                            const float *cuda_dIxy_ref,
                                            const int num_keypoints,
                                            const int patch_size,
                                            const Core::VectorX<double, 4> &intrinsics,
                                            const Core::VectorX<int, 2> &im_size_HW,
                                            const int spline_deg_k,
                                            const double spline_start_time,
                                            const double spline_sample_dt,
                                            const int *cpu_ctrl_knot_start_indices,
                                            const int num_ctrl_knots,
                                            const CudaSharedStorages &storages,
                                            const double huber_a,
                                            double *total_costs,
                                            double *cpu_hessian_tR,
                                            double *cpu_gradient_tR)
        {
            (void)cpu_ctrl_knot_start_indices;
            const double nan = std::numeric_limits<double>::quiet_NaN();
            *total_costs = nan;
            if (!storages.mbavo || n_frames < 1 || n_frames > MBAVO_MAX_FRAMES || num_ctrl_knots < 1 || num_ctrl_knots > 16)
            {
                std::fprintf(stderr, "mbavo: evaluate_cost_hessian_gradient: storages not initialised or sizes out of range\n");
                return;
            }
            // What the tracker uploaded into the storages (tracker.cpp:711-718, 729-733, 755-763) comes back in ONE copy: the
            // five small arrays share a slab (initialize_shared_cuda_storages).
            double cap[MBAVO_MAX_FRAMES], expo[MBAVO_MAX_FRAMES], kt[3 * 16], kR[4 * 16];
            unsigned char *cur[MBAVO_MAX_FRAMES];
            {
                alignas(16) unsigned char host[4096];
                if (!storages.mbavo_slab || storages.mbavo_slab_bytes > (int)sizeof host || n_frames > storages.mbavo_max_num_frames ||
                    num_ctrl_knots > storages.mbavo_max_num_ctrl_knots)
                {
                    std::fprintf(stderr, "mbavo: evaluate_cost_hessian_gradient: sizes exceed what the storages were initialised for\n");
                    return;
                }
                cudaMemcpy(host, storages.mbavo_slab, storages.mbavo_slab_bytes, cudaMemcpyDeviceToHost);
                const unsigned char *base = storages.mbavo_slab;
                auto at = [&](const void *dev) { return host + (reinterpret_cast<const unsigned char *>(dev) - base); };
                std::memcpy(cap, at(storages.cuda_img_cap_time), sizeof(double) * n_frames);
                std::memcpy(expo, at(storages.cuda_img_exp_time), sizeof(double) * n_frames);
                std::memcpy(cur, at(storages.cuda_cur_images), sizeof(void *) * n_frames);
                std::memcpy(kt, at(storages.cuda_spline_ctrl_knots_data_t), sizeof(double) * 3 * num_ctrl_knots);
                std::memcpy(kR, at(storages.cuda_spline_ctrl_knots_data_R), sizeof(double) * 4 * num_ctrl_knots);
            }

            mbavo_ctx *ctx = storages.mbavo;
            int rc = mbavo_set_frame_times(ctx, n_frames, cap, expo);

            mbavo_level lv{};
            lv.mem = MBAVO_MEM_DEVICE;
            lv.H = im_size_HW.values[0], lv.W = im_size_HW.values[1];
            lv.fx = intrinsics.values[0], lv.fy = intrinsics.values[1], lv.cx = intrinsics.values[2], lv.cy = intrinsics.values[3];
            lv.ref_I = cuda_ref_img, lv.ref_dIxy = cuda_dIxy_ref;
            lv.cur_I = cur, lv.n_frames = n_frames;
            lv.keypoint_xy = storages.cuda_keypoint_xy;
            lv.keypoint_xy_stride = (int)sizeof(Core::Vector2d);
            lv.keypoint_xy_offset = (int)(sizeof(Core::Vector2d) - 2 * sizeof(double)); // leading int nDim + padding, Vector.h:12-16
            lv.keypoint_z = storages.cuda_keypoint_depth_z;
            lv.num_keypoints = num_keypoints;
            lv.pattern_xy = storages.cuda_local_patch_pattern_xy;
            lv.patch_size = patch_size;
            lv.num_virtual_poses = n_vir_poses_per_frame;
            lv.ext_outlier_flags = storages.cuda_keypoints_outlier_flags;
            lv.ext_patch_cost = storages.cuda_patch_cost_gradient_hessian_tR;
            const int ndim = 6 * spline_deg_k + 1;
            lv.ext_patch_cost_stride = ndim * (ndim + 1) / 2;
            // the level (texels included) is re-derived on every call unless the caller opted into the cache and nothing it is
            // keyed on has changed
            LevelKey now;
            std::memset(static_cast<void *>(&now), 0, sizeof now); // padding included: the key is compared with memcmp
            now.valid = true, now.ref_I = cuda_ref_img, now.dIxy = cuda_dIxy_ref;
            for (int f = 0; f < n_frames; ++f)
                now.cur[f] = cur[f];
            now.H = lv.H, now.W = lv.W, now.P = num_keypoints, now.S = patch_size, now.N = n_vir_poses_per_frame, now.F = n_frames;
            now.k = spline_deg_k, now.epoch = storages.mbavo_keyframe_epoch;
            for (int e = 0; e < 4; ++e)
                now.K[e] = intrinsics.values[e];
            LevelKey *key = storages.mbavo_level_key;
            const bool hit = storages.mbavo_texel_cache != 0 && key && key->valid && std::memcmp(key, &now, sizeof now) == 0;
            if (rc == MBAVO_OK && !hit)
            {
                if (key)
                    key->valid = false;
                rc = mbavo_set_level(ctx, 0, &lv);
                if (rc == MBAVO_OK && key)
                    *key = now;
            }
            if (rc == MBAVO_OK)
                rc = mbavo_set_num_bad(ctx, 0, storages.num_bad_keypoints);
            mbavo_spline sp{spline_deg_k, spline_start_time, spline_sample_dt, num_ctrl_knots, kt, kR};
            if (rc == MBAVO_OK)
                rc = mbavo_evaluate(ctx, 0, &sp, huber_a, total_costs, cpu_hessian_tR, cpu_gradient_tR);
            if (rc != MBAVO_OK)
            {
                std::fprintf(stderr, "mbavo: evaluate_cost_hessian_gradient: %s\n", mbavo_last_error());
                *total_costs = nan;
            }
        }
    } // namespace VO
} // namespace SLAM
