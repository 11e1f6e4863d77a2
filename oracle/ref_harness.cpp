// TEST INFRASTRUCTURE — not part of the shipped product.
//
// oracle/_ref harness: drives the REFERENCE's own host-callable header arithmetic (included
// from /root/reference/src by include path, never copied into this repo) from restated host
// loops, so that the result is "the reference ba_tracker run on the CPU":
//
//   arithmetic leaves (called, not restated)
//     SplineSegmentStartKnotIdxAndNormalizedU   src/core/common/SplineFunctor.h:13-19
//     C2/C4SplineVec3Functor                    src/core/common/SplineFunctor.h:21-94
//     C2/C4SplineRot3Functor                    src/core/common/SplineFunctor.h:155-365
//     Quaterniond                               src/core/common/Quaternion.h:13-284
//     MatrixMatrixMultiply                      src/core/common/SmallBlas.h:152-225
//     compute_pixel_intensity<double>           src/ba_tracker/compute_pixel_intensity.h:91-209
//   loop structure (restated here, one function per reference kernel)
//     kernel_compute_virtual_camera_poses       src/ba_tracker/compute_virtual_camera_poses.cu:9-110
//     kernel_compute_local_patches_xy           src/ba_tracker/compute_local_patches_xy.cu:9-50
//     kernel_compute_pixel_jacobian_residual    src/ba_tracker/compute_hessian_gradients_cost.cu:23-156
//     kernel_compute_patch_cost_gradient_hessian  ...cost.cu:165-239
//     kernel_compute_frame_cost_gradient_hessian  ...cost.cu:247-283
//     merge_hessian_gradient_cost               src/ba_tracker/merge_hessian_gradient_cost.cpp:8-87
//     evaluate_cost_hessian_gradient            src/ba_tracker/spline_update_step.cpp:97-349
//
// Semantics are the REFERENCE's (per-frame k-knot window, Jacobians of every sample attributed to
// the frame's capture-time segment).  Where the reference kernel has undefined behaviour (threads
// returning before a barrier, stale scratch: SURVEY.md Appendix C) this harness uses: invalid
// pixel -> r = 0, J = 0; invalid sample -> contributes 0, divisor stays N.
//
// Built only where /root/reference exists (oracle/Makefile), output oracle/_ref/libmbavo_ref.so.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.

#include "ba_tracker/compute_pixel_intensity.h"
#include "core/common/CustomType.h"
#include "core/common/Quaternion.h"
#include "core/common/SmallBlas.h"
#include "core/common/SplineFunctor.h"
#include "core/common/Vector.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace SLAM;

namespace
{
    inline int packed_len(int k)
    {
        const int ndim = 6 * k + 1;
        return ndim * (ndim + 1) / 2;
    }
} // namespace

extern "C"
{
    int mbavo_ref_num_threads()
    {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
    }

    void mbavo_ref_set_num_threads(int n)
    {
#ifdef _OPENMP
        omp_set_num_threads(n);
#else
        (void)n;
#endif
    }

    // compute_virtual_camera_poses.cu:26-109, one iteration per (frame, virtual pose)
    // poses [F*N*7]; Jt [F*N*9k] / JR [F*N*12k] nullable together; seg_idx [F*N] nullable
    int mbavo_ref_virtual_poses(int N, int F, const double *cap, const double *exp_time, int k,
                                double t0, double dt, const double *knots_t, const double *knots_R,
                                double *poses, double *Jt, double *JR, int *seg_idx)
    {
        if (k != 2 && k != 4)
            return 1;
        for (int f = 0; f < F; ++f)
        {
            for (int i = 0; i < N; ++i)
            {
                const int g = f * N + i;
                const double t = cap[f] - exp_time[f] * 0.5 + i * exp_time[f] / (N - 1 + 1e-8);
                int idx;
                double u;
                Core::SplineSegmentStartKnotIdxAndNormalizedU(t, t0, dt, idx, u);
                if (seg_idx)
                    seg_idx[g] = idx;

                double jlogexp[3 * 24], X[16], Y[16], Z[16];
                double *jt = Jt ? Jt + (size_t)g * 9 * k : nullptr;
                double *jr = Jt ? JR + (size_t)g * 12 * k : nullptr;
                double *jle = Jt ? jlogexp : nullptr;
                Core::Vector3d tc;
                Core::Quaterniond Rc;
                if (k == 2)
                {
                    tc = Core::C2SplineVec3Functor(knots_t + idx * 3, u, jt);
                    Rc = Core::C2SplineRot3Functor(knots_R + idx * 4, u, jr, jle, jt ? X : nullptr, jt ? Y : nullptr, jt ? Z : nullptr);
                }
                else
                {
                    tc = Core::C4SplineVec3Functor(knots_t + idx * 3, u, jt);
                    Rc = Core::C4SplineRot3Functor(knots_R + idx * 4, u, jr, jle, jt ? X : nullptr, jt ? Y : nullptr, jt ? Z : nullptr);
                }
                double *p = poses + (size_t)g * 7;
                p[0] = tc(0), p[1] = tc(1), p[2] = tc(2);
                p[3] = Rc.x, p[4] = Rc.y, p[5] = Rc.z, p[6] = Rc.w;
            }
        }
        return 0;
    }

    // compute_local_patches_xy.cu:19-49.  xy [P*2], z [P] -> centres [F*P*2]
    int mbavo_ref_local_patches(int N, int F, const double *poses, const double *xy, const double *z, int P,
                                double fx, double fy, double cx, double cy, double *centres)
    {
        for (int f = 0; f < F; ++f)
        {
            const double *pose = poses + (size_t)(f * N + N / 2) * 7;
            const Core::Vector3d t_c2r(pose[0], pose[1], pose[2]);
            const Core::Quaterniond R_c2r(pose[3], pose[4], pose[5], pose[6]);
            const Core::Quaterniond R_r2c = R_c2r.conjugate();
            const Core::Vector3d t_r2c = -(R_r2c * t_c2r);
            for (int p = 0; p < P; ++p)
            {
                Core::Vector3d P3dr;
                P3dr(0) = z[p] * (xy[2 * p] - cx) / fx;
                P3dr(1) = z[p] * (xy[2 * p + 1] - cy) / fy;
                P3dr(2) = z[p];
                const Core::Vector3d P3dc = R_r2c * P3dr + t_r2c;
                centres[((size_t)f * P + p) * 2] = P3dc(0) / P3dc(2) * fx + cx;
                centres[((size_t)f * P + p) * 2 + 1] = P3dc(1) / P3dc(2) * fy + cy;
            }
        }
        return 0;
    }

    // One pixel of kernel_compute_pixel_jacobian_residual (…cost.cu:51-155): residual and (optionally)
    // the 1 x 6k Jacobian row in the frame-local knot window [t-block | w-block].
    static void ref_pixel(const unsigned char *I_ref, const float *dIxy, const unsigned char *I_cur,
                          int N, const double *poses_f, int k, const double *Jt_f, const double *JR_f,
                          double cxp, double cyp, int dx, int dy, double depth,
                          double fx, double fy, double cx, double cy, int H, int W,
                          double *residual, double *Jrow /*nullable, 6k*/)
    {
        const int d = 6 * k;
        *residual = 0;
        if (Jrow)
            std::memset(Jrow, 0, sizeof(double) * d);
        const int X = cxp + dx; // double -> int truncation, …cost.cu:69-70
        const int Y = cyp + dy;
        if (X < 0 || X > W - 1 || Y < 0 || Y > H - 1)
            return;
        const Core::Vector2d cur((double)X, (double)Y);

        double sumI = 0;
        FLOAT row[48];
        for (int i = 0; i < N; ++i)
        {
            const double *t_c2r = poses_f + (size_t)i * 7;
            const double *R_c2r = t_c2r + 3;
            double I, J7[7];
            if (!VO::compute_pixel_intensity<double>(I_ref, dIxy, H, W, R_c2r, t_c2r, depth, fx, fy, cx, cy, cur, &I,
                                                     Jrow ? J7 : nullptr))
                continue; // invalid sample contributes nothing
            sumI += I;
            if (Jrow)
            {
                // …cost.cu:136-142
                Core::MatrixMatrixMultiply<double, double, FLOAT, 0>(J7, 1, 3, Jt_f + (size_t)i * 9 * k, 3, 3 * k,
                                                                     row, 0, 0, 1, 3 * k);
                Core::MatrixMatrixMultiply<double, double, FLOAT, 0>(J7 + 3, 1, 4, JR_f + (size_t)i * 12 * k, 4, 3 * k,
                                                                     row, 0, 3 * k, 1, 3 * k);
                for (int c = 0; c < d; ++c)
                    Jrow[c] += row[c];
            }
        }
        *residual = sumI / float(N) - (double)I_cur[(size_t)Y * W + X]; // …cost.cu:120
        if (Jrow)
            for (int c = 0; c < d; ++c)
                Jrow[c] = Jrow[c] / float(N); // …cost.cu:152
    }

    // kernel 3 over all (frame, point, pixel): r [F*P*S], J [F*P*S*6k] nullable
    int mbavo_ref_pixel_jacobian_residual(const unsigned char *I_ref, const float *dIxy,
                                          const unsigned char *const *I_cur, int N, int F, const double *poses,
                                          int k, const double *Jt, const double *JR, const double *centres,
                                          const double *z, int P, const int *pattern, int S,
                                          double fx, double fy, double cx, double cy, int H, int W,
                                          double *r, double *J)
    {
        const int d = 6 * k;
#pragma omp parallel for schedule(static)
        for (long fp = 0; fp < (long)F * P; ++fp)
        {
            const int f = fp / P, p = fp % P;
            for (int j = 0; j < S; ++j)
            {
                const size_t pix = (size_t)fp * S + j;
                ref_pixel(I_ref, dIxy, I_cur[f], N, poses + (size_t)f * N * 7, k,
                          Jt ? Jt + (size_t)f * N * 9 * k : nullptr, Jt ? JR + (size_t)f * N * 12 * k : nullptr,
                          centres[2 * fp], centres[2 * fp + 1], pattern[2 * j], pattern[2 * j + 1], z[p],
                          fx, fy, cx, cy, H, W, r + pix, J ? J + pix * d : nullptr);
            }
        }
        return 0;
    }

    // One patch of kernel_compute_patch_cost_gradient_hessian (…cost.cu:179-238):
    // packed[E] = [sum rho, g (d), triu(H) row-major] * inv_num_residuals.  J nullable -> only packed[0].
    static void ref_patch(int S, int k, const double *r, const double *J, double huber_a, double inv_num_residuals,
                          double *packed)
    {
        const int ndim = 6 * k + 1;
        const int E = ndim * (ndim + 1) / 2;
        if (J)
            std::memset(packed, 0, sizeof(double) * E);
        double sum_rho = 0;
        double row[25];
        for (int j = 0; j < S; ++j)
        {
            row[0] = r[j];
            const double huber_aa = huber_a * huber_a;
            const double x = 0.5 * row[0] * row[0];
            double sqrt_drho_dx = 1.;
            double rho = x;
            if (x > huber_aa)
            {
                sqrt_drho_dx = sqrtf(huber_a / (sqrtf(x) + 1e-8));
                rho = 2 * huber_a * sqrtf(x) - huber_aa;
            }
            sum_rho += rho;
            row[0] = sqrt_drho_dx * row[0];
            if (J)
            {
                for (int c = 0; c < ndim - 1; ++c)
                    row[c + 1] = sqrt_drho_dx * J[(size_t)j * (ndim - 1) + c];
                int e = 0;
                for (int a = 0; a < ndim; ++a)
                    for (int b = a; b < ndim; ++b)
                        packed[e++] += row[a] * row[b];
            }
        }
        if (J)
            for (int e = 0; e < E; ++e)
                packed[e] *= inv_num_residuals;
        packed[0] = sum_rho * inv_num_residuals;
    }

    // kernel 4 over all (frame, point): packed [F*P*E]
    int mbavo_ref_patch_cost_gradient_hessian(int F, int P, int S, int k, const double *r, const double *J,
                                              double huber_a, double inv_num_residuals, double *packed)
    {
        const int E = packed_len(k);
        const int d = 6 * k;
#pragma omp parallel for schedule(static)
        for (long fp = 0; fp < (long)F * P; ++fp)
            ref_patch(S, k, r + (size_t)fp * S, J ? J + (size_t)fp * S * d : nullptr, huber_a, inv_num_residuals,
                      packed + (size_t)fp * E);
        return 0;
    }

    // kernel 5 (…cost.cu:254-282): frame_packed [F*E] (only element 0 when !with_hessian)
    int mbavo_ref_frame_reduce(int F, int P, int k, const double *packed, int with_hessian,
                               const unsigned char *flags, double *frame_packed)
    {
        const int E = packed_len(k);
        for (int f = 0; f < F; ++f)
            for (int e = 0; e < (with_hessian ? E : 1); ++e)
            {
                double s = 0;
                for (int p = 0; p < P; ++p)
                {
                    if (flags && flags[p] == 1)
                        continue;
                    s += packed[((size_t)f * P + p) * E + e];
                }
                frame_packed[(size_t)f * E + e] = s;
            }
        return 0;
    }

    // merge_hessian_gradient_cost.cpp:25-86.  H is (6n x 6n) column-major == row-major (symmetric); g [6n]
    int mbavo_ref_merge(int F, int k, const double *frame_packed, const int *seg_start, int n_knots,
                        double *total_cost, double *H, double *g)
    {
        const int ndim = 6 * k + 1;
        const int E = packed_len(k);
        const int Wd = 6 * n_knots;
        *total_cost = 0;
        if (H)
        {
            std::memset(H, 0, sizeof(double) * Wd * Wd);
            std::memset(g, 0, sizeof(double) * Wd);
        }
        for (int f = 0; f < F; ++f)
        {
            const double *v = frame_packed + (size_t)f * E;
            *total_cost += v[0];
            if (!H)
                continue;
            const int off0 = seg_start[f] * 3;
            const int off1 = (n_knots + seg_start[f]) * 3;
            int shift = off0;
            for (int j = 0; j < 3 * k; ++j)
                g[shift++] += v[j + 1];
            shift = off1;
            for (int j = 3 * k; j < 6 * k; ++j)
                g[shift++] += v[j + 1];
            const double *ptr = v + ndim;
            for (int j = 0; j < ndim - 1; ++j)
            {
                const int rr = j + (j < 3 * k ? off0 : off1 - 3 * k);
                for (int c = j; c < ndim - 1; ++c, ++ptr)
                {
                    const int cc = c + (c < 3 * k ? off0 : off1 - 3 * k);
                    H[(size_t)rr * Wd + cc] += *ptr;
                    if (cc != rr)
                        H[(size_t)cc * Wd + rr] += *ptr;
                }
            }
        }
        return 0;
    }

    // evaluate_cost_hessian_gradient (spline_update_step.cpp:97-349) without the P*S*N*d scratch: kernels 3,4,5
    // are streamed per point (identical arithmetic per pixel/patch; the across-patch sum runs in point order per
    // thread and thread order across threads instead of the GPU tree order).
    //   patch_costs [F*P] nullable : element 0 of every patch vector (what detectOutliers reads, tracker.cpp:647)
    //   H,g nullable together => cost-only branch (spline_update_step.cpp:242-348)
    int mbavo_ref_evaluate(int N, int F, const unsigned char *I_ref, const float *dIxy,
                           const unsigned char *const *I_cur, const double *cap, const double *exp_time,
                           const double *xy, const double *z, int P, const int *pattern, int S,
                           const unsigned char *flags, int num_bad,
                           double fx, double fy, double cx, double cy, int H_img, int W_img,
                           int k, double t0, double dt, const double *knots_t, const double *knots_R,
                           const int *seg_start, int n_knots, double huber_a,
                           double *total_cost, double *H, double *g, double *patch_costs)
    {
        if (k != 2 && k != 4)
            return 1;
        const int E = packed_len(k);
        const int d = 6 * k;
        const bool with_h = H != nullptr;
        const int num_residuals = (P - num_bad) * F * S; // spline_update_step.cpp:116
        const double inv_num_residuals = 1.0 / num_residuals;

        std::vector<double> poses((size_t)F * N * 7), Jt, JR, centres((size_t)F * P * 2);
        if (with_h)
        {
            Jt.resize((size_t)F * N * 9 * k);
            JR.resize((size_t)F * N * 12 * k);
        }
        mbavo_ref_virtual_poses(N, F, cap, exp_time, k, t0, dt, knots_t, knots_R, poses.data(),
                                with_h ? Jt.data() : nullptr, with_h ? JR.data() : nullptr, nullptr);
        mbavo_ref_local_patches(N, F, poses.data(), xy, z, P, fx, fy, cx, cy, centres.data());

        std::vector<double> frame_packed((size_t)F * E, 0.0);
        int nthreads = 1;
#ifdef _OPENMP
        nthreads = omp_get_max_threads();
#endif
        std::vector<double> partial((size_t)nthreads * F * E, 0.0);
#pragma omp parallel
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            double *acc = partial.data() + (size_t)tid * F * E;
            std::vector<double> r(S), J(with_h ? (size_t)S * d : 0), packed(E);
#pragma omp for schedule(static)
            for (long fp = 0; fp < (long)F * P; ++fp)
            {
                const int f = fp / P, p = fp % P;
                for (int j = 0; j < S; ++j)
                    ref_pixel(I_ref, dIxy, I_cur[f], N, poses.data() + (size_t)f * N * 7, k,
                              with_h ? Jt.data() + (size_t)f * N * 9 * k : nullptr,
                              with_h ? JR.data() + (size_t)f * N * 12 * k : nullptr,
                              centres[2 * fp], centres[2 * fp + 1], pattern[2 * j], pattern[2 * j + 1], z[p],
                              fx, fy, cx, cy, H_img, W_img, &r[j], with_h ? &J[(size_t)j * d] : nullptr);
                ref_patch(S, k, r.data(), with_h ? J.data() : nullptr, huber_a, inv_num_residuals, packed.data());
                if (patch_costs)
                    patch_costs[fp] = packed[0];
                if (flags && flags[p] == 1)
                    continue;
                for (int e = 0; e < (with_h ? E : 1); ++e)
                    acc[(size_t)f * E + e] += packed[e];
            }
        }
        for (int t = 0; t < nthreads; ++t)
            for (size_t e = 0; e < (size_t)F * E; ++e)
                frame_packed[e] += partial[(size_t)t * F * E + e];

        return mbavo_ref_merge(F, k, frame_packed.data(), seg_start, n_knots, total_cost, H, g);
    }

    // compute_pixel_intensity<double> itself (compute_pixel_intensity.h:91-209); pose = [t(3), q(x,y,z,w)]
    int mbavo_ref_pixel_intensity(const unsigned char *I_ref, const float *dIxy, int H, int W, const double *pose, double D,
                                  double fx, double fy, double cx, double cy, double X, double Y, double *intensity,
                                  double *J7)
    {
        Core::VectorX<double, 2> cur;
        cur.values[0] = X;
        cur.values[1] = Y;
        return VO::compute_pixel_intensity<double>(I_ref, dIxy, H, W, pose + 3, pose, D, fx, fy, cx, cy, cur, intensity, J7) ? 1 : 0;
    }

    // One pose on the spline by the reference functors (what SplineSE3::GetPose evaluates, Spline.h:120-170): out7 = t then q (x,y,z,w).
    // Used by tests/golden/make_golden.py to store, next to the golden blurred frame, the very poses it was rendered with.
    int mbavo_ref_spline_pose(int k, double t0, double dt, const double *knots_t, const double *knots_R, double t, double *out7)
    {
        int idx;
        double u;
        Core::SplineSegmentStartKnotIdxAndNormalizedU(t, t0, dt, idx, u);
        Core::Vector3d tc;
        Core::Quaterniond Rc;
        if (k == 2)
        {
            tc = Core::C2SplineVec3Functor(knots_t + idx * 3, u);
            Rc = Core::C2SplineRot3Functor(knots_R + idx * 4, u);
        }
        else if (k == 4)
        {
            tc = Core::C4SplineVec3Functor(knots_t + idx * 3, u);
            Rc = Core::C4SplineRot3Functor(knots_R + idx * 4, u);
        }
        else
            return 1;
        out7[0] = tc(0), out7[1] = tc(1), out7[2] = tc(2);
        out7[3] = Rc.x, out7[4] = Rc.y, out7[5] = Rc.z, out7[6] = Rc.w;
        return 0;
    }

    // generate_synthetic_data.cpp:127-180 (warp_image + synthesize_motion_blurred_img) with the pose supplied by
    // the reference spline functors (SplineSE3::GetPose uses the same functors, Spline.h:120-170).
    int mbavo_ref_synthesize_blurred(const unsigned char *I_ref, int H, int W, double plane_depth,
                                     double fx, double fy, double cx, double cy, int k, double t0, double dt,
                                     const double *knots_t, const double *knots_R, double capture_time,
                                     double exposure_time, int num_samples, unsigned char *out)
    {
        std::vector<float> acc((size_t)H * W, 0.f);
        for (int i = 0; i < num_samples; ++i)
        {
            const double t = capture_time - exposure_time * 0.5 + i * exposure_time / (num_samples - 1);
            int idx;
            double u;
            Core::SplineSegmentStartKnotIdxAndNormalizedU(t, t0, dt, idx, u);
            Core::Vector3d tc;
            Core::Quaterniond Rc;
            if (k == 2)
            {
                tc = Core::C2SplineVec3Functor(knots_t + idx * 3, u);
                Rc = Core::C2SplineRot3Functor(knots_R + idx * 4, u);
            }
            else
            {
                tc = Core::C4SplineVec3Functor(knots_t + idx * 3, u);
                Rc = Core::C4SplineRot3Functor(knots_R + idx * 4, u);
            }
            const double R[4] = {Rc.x, Rc.y, Rc.z, Rc.w};
            const double tt[3] = {tc(0), tc(1), tc(2)};
#pragma omp parallel for schedule(static)
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c)
                {
                    Core::VectorX<double, 2> cur;
                    cur.values[0] = c;
                    cur.values[1] = r;
                    double intensity = 0;
                    VO::compute_pixel_intensity<double>(I_ref, nullptr, H, W, R, tt, plane_depth, fx, fy, cx, cy, cur,
                                                        &intensity, nullptr);
                    acc[(size_t)r * W + c] += (float)(unsigned char)intensity; // `*im_cur_data_ptr = intensity`
                }
        }
        for (size_t i = 0; i < (size_t)H * W; ++i)
        {
            const float v = acc[i] / num_samples;
            long q = lrintf(v); // cv::saturate_cast<uchar>(float) == cvRound (round half to even) + clamp
            out[i] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
        return 0;
    }
}
