// TEST / BASELINE INFRASTRUCTURE — not part of the shipped product.
//
// GPU baseline of BASELINE.md §3.2: the REFERENCE's own CUDA kernels (src/ba_tracker/compute_virtual_camera_poses.cu,
// compute_local_patches_xy.cu, compute_hessian_gradients_cost.cu — compiled unmodified from /root/reference by
// oracle/Makefile, nothing copied) driven by a restatement of evaluate_cost_hessian_gradient
// (src/ba_tracker/spline_update_step.cpp:97-349) and merge_hessian_gradient_cost (merge_hessian_gradient_cost.cpp:8-87)
// that needs no Eigen.  Points are processed in chunks of <= 65535 because the reference launches gridDim.y = P
// (compute_hessian_gradients_cost.cu:309).  Every wrapper ends in cudaDeviceSynchronize, as the reference runs.
#include "ba_tracker/compute_hessian_gradients_cost.h"
#include "ba_tracker/compute_local_patches_xy.h"
#include "ba_tracker/compute_virtual_camera_poses.h"
#include "core/common/CustomType.h"
#include "core/common/Vector.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

using namespace SLAM;

namespace
{
    struct RefCtx
    {
        int maxF, maxN, maxP, maxS, maxKnots, k, chunk;
        // spline_update_step.cpp:9-58
        double *cap, *expo, *z, *knots_t, *knots_R, *poses, *Jt, *JR, *Jlogexp, *X, *Y, *Z;
        Core::Vector2d *xy, *centres;
        unsigned char *flags, **cur_imgs;
        int *pattern;
        double *pix_r, *pix_J, *patch, *frame;
        FLOAT *scratch;
        // level data
        unsigned char *ref_I = nullptr, *cur_I[16] = {};
        float *dIxy = nullptr;
        int H = 0, W = 0, P = 0, S = 0, N = 0, F = 0;
        double fx, fy, cx, cy;
    };
    int packed_len(int k)
    {
        const int n = 6 * k + 1;
        return n * (n + 1) / 2;
    }
} // namespace

extern "C"
{
    void *mbavo_refcuda_create(int maxF, int maxN, int maxP, int maxS, int maxKnots, int k)
    {
        RefCtx *c = new RefCtx();
        c->maxF = maxF, c->maxN = maxN, c->maxP = maxP, c->maxS = maxS, c->maxKnots = maxKnots, c->k = k;
        c->chunk = std::min(maxP, 65535);
        const size_t nposes = (size_t)maxF * maxN, npatch = (size_t)maxF * c->chunk, npix = npatch * maxS;
        const int E = packed_len(k);
        cudaMalloc(&c->cap, sizeof(double) * maxF);
        cudaMalloc(&c->expo, sizeof(double) * maxF);
        cudaMalloc(&c->z, sizeof(double) * maxP);
        cudaMalloc(&c->xy, sizeof(Core::Vector2d) * maxP);
        cudaMalloc(&c->flags, maxP);
        cudaMemset(c->flags, 0, maxP);
        cudaMalloc(&c->cur_imgs, sizeof(void *) * maxF);
        cudaMalloc(&c->pattern, sizeof(int) * 2 * maxS);
        cudaMalloc(&c->knots_t, sizeof(double) * 3 * maxKnots);
        cudaMalloc(&c->knots_R, sizeof(double) * 4 * maxKnots);
        cudaMalloc(&c->poses, sizeof(double) * nposes * 7);
        cudaMalloc(&c->Jt, sizeof(double) * nposes * 9 * k);
        cudaMalloc(&c->JR, sizeof(double) * nposes * 12 * k);
        cudaMalloc(&c->Jlogexp, sizeof(double) * nposes * (k - 1) * 24);
        cudaMalloc(&c->X, sizeof(double) * nposes * 16);
        cudaMalloc(&c->Y, sizeof(double) * nposes * 16);
        cudaMalloc(&c->Z, sizeof(double) * nposes * 16);
        cudaMalloc(&c->centres, sizeof(Core::Vector2d) * npatch);
        cudaMalloc(&c->pix_r, sizeof(double) * npix);
        cudaMalloc(&c->pix_J, sizeof(double) * npix * 6 * k);
        cudaMalloc(&c->scratch, sizeof(FLOAT) * npix * 6 * k * maxN);
        cudaMalloc(&c->patch, sizeof(double) * npatch * E);
        cudaMalloc(&c->frame, sizeof(double) * maxF * E);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        {
            delete c;
            return nullptr;
        }
        return c;
    }

    void mbavo_refcuda_destroy(void *h)
    {
        RefCtx *c = (RefCtx *)h;
        if (!c)
            return;
        void *ptrs[] = {c->cap, c->expo, c->z, c->xy, c->flags, c->cur_imgs, c->pattern, c->knots_t, c->knots_R, c->poses,
                        c->Jt, c->JR, c->Jlogexp, c->X, c->Y, c->Z, c->centres, c->pix_r, c->pix_J, c->scratch, c->patch,
                        c->frame, c->ref_I, c->dIxy};
        for (void *p : ptrs)
            cudaFree(p);
        for (auto p : c->cur_I)
            cudaFree(p);
        delete c;
    }

    // what the tracker uploads: tracker.cpp:701-751 and Image<T>::uploadToGpu
    int mbavo_refcuda_set_level(void *h, int H, int W, double fx, double fy, double cx, double cy, const unsigned char *ref_I,
                                const float *dIxy, const unsigned char *const *cur_I, int F, const double *cap,
                                const double *expo, const double *xy, const double *z, int P, const int *pattern, int S, int N)
    {
        RefCtx *c = (RefCtx *)h;
        if (F > c->maxF || P > c->maxP || S > c->maxS || N > c->maxN)
            return 1;
        const size_t npix = (size_t)H * W;
        cudaFree(c->ref_I), cudaFree(c->dIxy);
        cudaMalloc(&c->ref_I, npix);
        cudaMalloc(&c->dIxy, npix * 8);
        cudaMemcpy(c->ref_I, ref_I, npix, cudaMemcpyHostToDevice);
        cudaMemcpy(c->dIxy, dIxy, npix * 8, cudaMemcpyHostToDevice);
        for (int f = 0; f < F; ++f)
        {
            cudaFree(c->cur_I[f]);
            cudaMalloc(&c->cur_I[f], npix);
            cudaMemcpy(c->cur_I[f], cur_I[f], npix, cudaMemcpyHostToDevice);
        }
        cudaMemcpy(c->cur_imgs, c->cur_I, sizeof(void *) * F, cudaMemcpyHostToDevice);
        std::vector<Core::Vector2d> v(P);
        for (int p = 0; p < P; ++p)
            v[p] = Core::Vector2d(xy[2 * p], xy[2 * p + 1]);
        cudaMemcpy(c->xy, v.data(), sizeof(Core::Vector2d) * P, cudaMemcpyHostToDevice);
        cudaMemcpy(c->z, z, sizeof(double) * P, cudaMemcpyHostToDevice);
        cudaMemcpy(c->pattern, pattern, sizeof(int) * 2 * S, cudaMemcpyHostToDevice);
        cudaMemcpy(c->cap, cap, sizeof(double) * F, cudaMemcpyHostToDevice);
        cudaMemcpy(c->expo, expo, sizeof(double) * F, cudaMemcpyHostToDevice);
        cudaMemset(c->flags, 0, c->maxP);
        c->H = H, c->W = W, c->P = P, c->S = S, c->N = N, c->F = F;
        c->fx = fx, c->fy = fy, c->cx = cx, c->cy = cy;
        return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
    }

    // outlier flags as detectOutliersAndUploadToGpu leaves them (tracker.cpp:696-698); flags == nullptr clears them
    int mbavo_refcuda_set_flags(void *h, const unsigned char *flags)
    {
        RefCtx *c = (RefCtx *)h;
        if (flags)
            cudaMemcpy(c->flags, flags, c->P, cudaMemcpyHostToDevice);
        else
            cudaMemset(c->flags, 0, c->maxP);
        return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
    }

    // evaluate_cost_hessian_gradient (spline_update_step.cpp:97-349) + merge (merge_hessian_gradient_cost.cpp:8-87)
    int mbavo_refcuda_evaluate(void *h, double t0, double dt, const double *knots_t, const double *knots_R, int n_knots,
                               const int *seg_start, double huber_a, int num_bad, double *total_cost, double *Hout,
                               double *gout)
    {
        RefCtx *c = (RefCtx *)h;
        const int k = c->k, E = packed_len(k), ndim = 6 * k + 1;
        const bool with_h = Hout != nullptr;
        cudaMemcpy(c->knots_t, knots_t, sizeof(double) * 3 * n_knots, cudaMemcpyHostToDevice); // tracker.cpp:755-763
        cudaMemcpy(c->knots_R, knots_R, sizeof(double) * 4 * n_knots, cudaMemcpyHostToDevice);
        const double inv_nr = 1.0 / ((double)(c->P - num_bad) * c->F * c->S);
        Core::VectorX<double, 4> K;
        K.values[0] = c->fx, K.values[1] = c->fy, K.values[2] = c->cx, K.values[3] = c->cy;
        Core::VectorX<int, 2> HW;
        HW.values[0] = c->H, HW.values[1] = c->W;

        if (with_h)
            VO::compute_virtual_camera_poses(c->N, c->F, c->cap, c->expo, k, t0, dt, c->knots_t, c->knots_R, c->poses, c->Jt,
                                             c->JR, c->Jlogexp, c->X, c->Y, c->Z);
        else
            VO::compute_virtual_camera_poses(c->N, c->F, c->cap, c->expo, k, t0, dt, c->knots_t, c->knots_R, c->poses);

        std::vector<double> frame_total((size_t)c->F * E, 0.0), frame_chunk((size_t)c->F * E);
        for (int p0 = 0; p0 < c->P; p0 += c->chunk)
        {
            const int Pc = std::min(c->chunk, c->P - p0);
            VO::compute_local_patches_xy(c->N, c->F, c->poses, c->xy + p0, c->z + p0, Pc, K, HW, c->centres);
            // the reference leaves pixel_jacobians of invalid pixels stale (Appendix C); give it zeros to start from
            if (with_h)
                cudaMemset(c->pix_J, 0, sizeof(double) * (size_t)c->F * Pc * c->S * 6 * k);
            VO::compute_pixel_jacobian_residual(c->ref_I, c->dIxy, c->cur_imgs, c->N, c->F, c->poses, k, c->Jt, c->JR,
                                                c->centres, c->z + p0, Pc, c->pattern, c->S, K, HW,
                                                with_h ? c->scratch : nullptr, c->pix_r, with_h ? c->pix_J : nullptr);
            VO::compute_patch_cost_gradient_hessian(c->F, Pc, c->S, k, c->pix_r, with_h ? c->pix_J : nullptr, huber_a, inv_nr,
                                                    c->patch);
            VO::compute_frame_cost_gradient_hessian(c->F, Pc, k, c->patch, with_h, c->flags + p0, c->frame);
            cudaMemcpy(frame_chunk.data(), c->frame, sizeof(double) * c->F * E, cudaMemcpyDeviceToHost);
            for (size_t e = 0; e < frame_chunk.size(); ++e)
                frame_total[e] += (with_h || e % E == 0) ? frame_chunk[e] : 0.0;
        }
        if (cudaGetLastError() != cudaSuccess)
            return 2;

        const int Wd = 6 * n_knots;
        *total_cost = 0;
        if (with_h)
        {
            std::memset(Hout, 0, sizeof(double) * Wd * Wd);
            std::memset(gout, 0, sizeof(double) * Wd);
        }
        for (int f = 0; f < c->F; ++f)
        {
            const double *v = frame_total.data() + (size_t)f * E;
            *total_cost += v[0];
            if (!with_h)
                continue;
            const int off0 = seg_start[f] * 3, off1 = (n_knots + seg_start[f]) * 3;
            for (int j = 0; j < 3 * k; ++j)
                gout[off0 + j] += v[j + 1];
            for (int j = 3 * k; j < 6 * k; ++j)
                gout[off1 + j - 3 * k] += v[j + 1];
            const double *ptr = v + ndim;
            for (int j = 0; j < ndim - 1; ++j)
            {
                const int rr = j + (j < 3 * k ? off0 : off1 - 3 * k);
                for (int cidx = j; cidx < ndim - 1; ++cidx, ++ptr)
                {
                    const int cc = cidx + (cidx < 3 * k ? off0 : off1 - 3 * k);
                    Hout[(size_t)rr * Wd + cc] += *ptr;
                    if (cc != rr)
                        Hout[(size_t)cc * Wd + rr] += *ptr;
                }
            }
        }
        return 0;
    }
}
