// Exposure-sample poses on the SE(3) spline and their Jacobians w.r.t. the control knots — the device side of
// kernel_compute_virtual_camera_poses (src/ba_tracker/compute_virtual_camera_poses.cu:9-110) and of the spline
// functors it calls (src/core/common/SplineFunctor.h:13-365, src/core/common/Quaternion.h:61-233), re-designed:
// one thread per (frame, sample), everything in registers / local memory (the reference keeps 24(k-1)+48 doubles
// of scratch per pose in global memory), fp64, and the result is emitted in the fp32 form the fused tracking
// kernel consumes (rotation matrix instead of quaternion, so(3) Jacobian blocks instead of 4 x 3k quaternion
// Jacobians).  The same launch also prepares the per-frame fp64 pose used for the patch centres
// (compute_local_patches_xy.cu:26-43) and the per-frame segment boundaries.
#include "pose_device.cuh"

namespace mbavo
{
    namespace
    {
        // one thread per (frame, sample).  knots_from: 0 = the launch parameter (and, if gn != nullptr, the sweep state is
        // initialised from it), 1 = gn->cur_*, 2 = gn->cand_* (device-resident Gauss-Newton sweep).
        template <int K>
        __global__ void pose_kernel(const __grid_constant__ EvalStage stage, int with_jacobian, float *__restrict__ samples,
                                    double *__restrict__ mid, int *__restrict__ seg_end, GnState *gn, int knots_from, int buf_select,
                                    int samples_stride, int mid_stride, int seg_end_stride, double *dbg)
        {
            cudaTriggerProgrammaticLaunchCompletion(); // the tracking kernel may start its prologue now
            const double *kt = stage.knots_t, *kR = stage.knots_R;
            if (gn)
            {
                // this grid may itself have been launched programmatically behind the previous tracking kernel
                cudaGridDependencySynchronize();
                if (knots_from == 1)
                    kt = gn->cur_t, kR = gn->cur_R;
                else if (knots_from == 2)
                    kt = gn->cand_t, kR = gn->cand_R;
                else if (blockIdx.x == 0)
                {
                    for (int e = threadIdx.x; e < 3 * stage.n_knots; e += blockDim.x)
                        gn->cur_t[e] = stage.knots_t[e];
                    for (int e = threadIdx.x; e < 4 * stage.n_knots; e += blockDim.x)
                        gn->cur_R[e] = stage.knots_R[e];
                    if (threadIdx.x == 0)
                        gn->status = 0, gn->cur_buf = 0;
                }
            }
            int buf = buf_select == kBufB ? 1 : 0;
            if (buf_select == kBufCur || buf_select == kBufCand)
                buf = buf_select == kBufCur ? gn->cur_buf : 1 - gn->cur_buf;
            samples += (size_t)buf * samples_stride, mid += (size_t)buf * mid_stride, seg_end += (size_t)buf * seg_end_stride;
            const int g = blockIdx.x * blockDim.x + threadIdx.x;
            if (g < stage.N * stage.F)
                pose_one<K>(&stage, kt, kR, g, with_jacobian, samples, mid, seg_end, dbg);
        }
    } // namespace

    // Host evaluation of one spline pose (SplineSE3::GetPose without Jacobians, src/core/common/Spline.h:222-290) with the
    // same functions the kernel runs: kt / kR point at the first of the K knots of the segment, u is the normalised time.
    int host_spline_pose(int K, const double *kt, const double *kR, double u, double *t_out, double *q_out)
    {
        Q q{0, 0, 0, 1};
        double wt[4];
        if (K == 2)
            spline_pose<2>(kt, kR, u, t_out, q, wt, nullptr);
        else if (K == 4)
            spline_pose<4>(kt, kR, u, t_out, q, wt, nullptr);
        else
            return -1;
        q_out[0] = q.x, q_out[1] = q.y, q_out[2] = q.z, q_out[3] = q.w;
        return 0;
    }

    // gn / knots_from: see pose_kernel.  dependent: launch programmatically behind the previous kernel of the stream (the
    // kernel then waits for it before it reads the sweep state).
    cudaError_t launch_pose_kernel(int K, const EvalStage &stage, int total_samples, int with_jacobian, float *samples,
                                   double *mid, int *seg_end, cudaStream_t stream, GnState *gn, int knots_from, bool dependent,
                                   int buf_select, int samples_stride, int mid_stride, int seg_end_stride, double *dbg)
    {
        const int threads = 64;
        const int blocks = (total_samples + threads - 1) / threads;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks, 1, 1), cfg.blockDim = dim3(threads, 1, 1), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = (dependent && gn) ? 1 : 0;
        if (K == 2)
            return cudaLaunchKernelEx(&cfg, pose_kernel<2>, stage, with_jacobian, samples, mid, seg_end, gn, knots_from, buf_select, samples_stride,
                                      mid_stride, seg_end_stride, dbg);
        return cudaLaunchKernelEx(&cfg, pose_kernel<4>, stage, with_jacobian, samples, mid, seg_end, gn, knots_from, buf_select, samples_stride,
                                  mid_stride, seg_end_stride, dbg);
    }
} // namespace mbavo
