// Device / host arithmetic of one exposure-sample pose on the SE(3) spline (shared by pose_kernel.cu and the persistent
// sweep kernel of track_kernel.cuh, whose last block computes the candidate's sample records itself):
// the spline functors of src/core/common/SplineFunctor.h:13-365 and Quaternion.h:61-233 re-derived in so(3) form, and
// pose_one = the body of kernel_compute_virtual_camera_poses (src/ba_tracker/compute_virtual_camera_poses.cu:26-109)
// emitting the fp32 sample record the tracking kernel consumes (layout: mbavo_device.h).
#ifndef MBAVO_POSE_DEVICE_CUH_
#define MBAVO_POSE_DEVICE_CUH_
#include "mbavo_device.h"

#include <math.h>

namespace mbavo
{
    namespace
    {
        struct Q
        {
            double x, y, z, w;
        };

        __host__ __device__ __forceinline__ Q qmul(const Q &a, const Q &b) // Quaternion.h:44-50
        {
            return Q{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                     a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
        }
        __host__ __device__ __forceinline__ Q qconj(const Q &a) { return Q{-a.x, -a.y, -a.z, a.w}; }

        // q (x) p = L(q) p ; q (x) p = Rhat(p) q      (Quaternion.h:239-283), 4x4 row-major
        __host__ __device__ __forceinline__ void left_matrix(const Q &q, double *M)
        {
            M[0] = q.w, M[1] = -q.z, M[2] = q.y, M[3] = q.x;
            M[4] = q.z, M[5] = q.w, M[6] = -q.x, M[7] = q.y;
            M[8] = -q.y, M[9] = q.x, M[10] = q.w, M[11] = q.z;
            M[12] = -q.x, M[13] = -q.y, M[14] = -q.z, M[15] = q.w;
        }
        __host__ __device__ __forceinline__ void right_matrix(const Q &q, double *M)
        {
            M[0] = q.w, M[1] = q.z, M[2] = -q.y, M[3] = q.x;
            M[4] = -q.z, M[5] = q.w, M[6] = q.x, M[7] = q.y;
            M[8] = q.y, M[9] = -q.x, M[10] = q.w, M[11] = q.z;
            M[12] = -q.x, M[13] = -q.y, M[14] = -q.z, M[15] = q.w;
        }

        // rotation vector of q and G = d phi / d q (3x4).  Branches and derivative expressions as in
        // Quaternion::log (Quaternion.h:61-152): the Jacobians of the pose inherit them.
        __host__ __device__ inline void so3_log(const Q &q, double *phi, double *G)
        {
            const double v[3] = {q.x, q.y, q.z};
            const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
            double lam, dl[4] = {0, 0, 0, 0};
            if (n2 < 1e-20)
            {
                const double w3 = q.w * q.w * q.w;
                lam = 2. / q.w - 2. / 3. * n2 / w3;
                for (int c = 0; c < 3; ++c)
                    dl[c] = 2. / q.w - 4. / 3. * v[c] / w3;
                dl[3] = -2 / (q.w * q.w) + 2 * n2 / (w3 * q.w);
            }
            else
            {
                const double n = sqrt(n2);
                if (fabs(q.w) < 1e-10)
                {
                    const double sgn = q.w > 0 ? 1.0 : -1.0;
                    lam = sgn * M_PI / n;
                    for (int c = 0; c < 3; ++c)
                        dl[c] = -sgn * lam / n2 * v[c];
                }
                else
                {
                    lam = 2.0 * atan(n / q.w) / n;
                    const double dn = (2 * q.w - lam) / n;
                    for (int c = 0; c < 3; ++c)
                        dl[c] = dn * v[c] / n;
                    dl[3] = -2.;
                }
            }
            for (int r = 0; r < 3; ++r)
            {
                phi[r] = lam * v[r];
                for (int c = 0; c < 4; ++c)
                    G[r * 4 + c] = dl[c] * v[r] + (r == c ? lam : 0.0);
            }
        }

        // unit quaternion of a rotation vector and E = d q / d phi (4x3)      (Quaternion::exp, Quaternion.h:154-233)
        __host__ __device__ inline void so3_exp(const double *phi, Q &q, double *E)
        {
            const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
            double fi, fr;
            if (t2 < 1e-20)
            {
                const double t4 = t2 * t2;
                fi = 0.5 - 1. / 48. * t2 + 1. / 3840. * t4;
                fr = 1. - 1. / 8. * t2 + 1. / 384. * t4;
                for (int e = 0; e < 12; ++e)
                    E[e] = 0;
                E[0] = E[4] = E[8] = 0.5;
            }
            else
            {
                const double t = sqrt(t2), s = sin(0.5 * t);
                fi = s / t;
                fr = cos(0.5 * t);
                const double dfi = 0.5 * fr / t - fi / t, dfr = -0.5 * s;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        E[r * 3 + c] = dfi * (phi[c] / t) * phi[r] + (r == c ? fi : 0.0);
                for (int c = 0; c < 3; ++c)
                    E[9 + c] = dfr * (phi[c] / t);
            }
            q = Q{fi * phi[0], fi * phi[1], fi * phi[2], fr};
        }

        template <int M, int P, int N>
        __host__ __device__ __forceinline__ void mat_mul(const double *A, const double *B, double *C)
        {
            for (int i = 0; i < M; ++i)
                for (int j = 0; j < N; ++j)
                {
                    double s = 0;
                    for (int l = 0; l < P; ++l)
                        s += A[i * P + l] * B[l * N + j];
                    C[i * N + j] = s;
                }
        }

        // blend weights of the K knots of a segment (SplineFunctor.h:24-28,176 / :47-54,232-234)
        template <int K>
        __host__ __device__ __forceinline__ void spline_weights(double u, double *wt, double *wr)
        {
            if (K == 2)
            {
                wt[0] = 1 - u, wt[1] = u;
                wr[0] = 1, wr[1] = u;
            }
            else
            {
                const double uu = u * u, uuu = uu * u, s = 1. / 6.;
                wt[0] = s - 0.5 * u + 0.5 * uu - s * uuu;
                wt[1] = 4 * s - uu + 0.5 * uuu;
                wt[2] = s + 0.5 * u + 0.5 * uu - 0.5 * uuu;
                wt[3] = s * uuu;
                wr[0] = 1;
                wr[1] = 5 * s + 0.5 * u - 0.5 * uu + s * uuu;
                wr[2] = s + 0.5 * u + 0.5 * uu - 2 * s * uuu;
                wr[3] = s * uuu;
            }
        }

        // One pose:  q = q_0 (x) A_1 (x) ... (x) A_{K-1},  A_j = Exp(c_j Log(q_{j-1}* (x) q_j)).
        // Theta_m (3x3) = d theta / d w_m where q(w) = q (x) Exp(theta): knot m acts as leading factor (m = 0), through
        // d_m = q_{m-1}* q_m, and through d_{m+1} = q_m* q_{m+1}.  With JR_m = dq/dw_m (the reference's 4x3 block),
        // Theta_m = 2 [L(q)^T JR_m]_{rows 0..2}.
        template <int K>
        __host__ __device__ inline void spline_pose(const double *kt, const double *kR, double u, double *t_out, Q &q_out, double *wt,
                                    double *Theta /* K x 9, nullable */)
        {
            double wr[K];
            spline_weights<K>(u, wt, wr);
            for (int a = 0; a < 3; ++a)
            {
                double s = 0;
                for (int j = 0; j < K; ++j)
                    s += wt[j] * kt[3 * j + a];
                t_out[a] = s;
            }
            Q knot[K], A[K], pre[K + 1], post[K];
            double G[K][12], E[K][12];
            for (int j = 0; j < K; ++j)
                knot[j] = Q{kR[4 * j], kR[4 * j + 1], kR[4 * j + 2], kR[4 * j + 3]};
            for (int j = 1; j < K; ++j)
            {
                double phi[3];
                so3_log(qmul(qconj(knot[j - 1]), knot[j]), phi, G[j]);
                phi[0] *= wr[j], phi[1] *= wr[j], phi[2] *= wr[j];
                so3_exp(phi, A[j], E[j]);
            }
            pre[1] = knot[0];
            for (int j = 1; j < K; ++j)
                pre[j + 1] = qmul(pre[j], A[j]);
            q_out = pre[K];
            if (Theta == nullptr)
                return;

            post[K - 1] = Q{0, 0, 0, 1};
            for (int j = K - 2; j >= 0; --j)
                post[j] = qmul(A[j + 1], post[j + 1]);

            double Lq[16];
            left_matrix(q_out, Lq);
            for (int m = 0; m < K; ++m)
            {
                double Lm[16], dq[12], acc[12];
                left_matrix(knot[m], Lm);
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 3; ++c)
                        dq[r * 3 + c] = 0.5 * Lm[r * 4 + c];
                for (int e = 0; e < 12; ++e)
                    acc[e] = 0;
                if (m == 0)
                {
                    double Rp[16], T[12];
                    right_matrix(post[0], Rp);
                    mat_mul<4, 4, 3>(Rp, dq, T);
                    for (int e = 0; e < 12; ++e)
                        acc[e] += T[e];
                }
                for (int path = 0; path < 2; ++path)
                {
                    const int j = path == 0 ? m : m + 1;
                    if (j < 1 || j > K - 1)
                        continue;
                    double Dd[16], T1[12], T2[9], T3[12], T4[12], M4[16];
                    if (path == 0)
                        left_matrix(qconj(knot[m - 1]), Dd);
                    else
                    {
                        right_matrix(knot[m + 1], Dd);
                        for (int r = 0; r < 4; ++r)
                            for (int c = 0; c < 3; ++c)
                                Dd[r * 4 + c] = -Dd[r * 4 + c];
                    }
                    mat_mul<4, 4, 3>(Dd, dq, T1);
                    mat_mul<3, 4, 3>(G[j], T1, T2);
                    for (int e = 0; e < 9; ++e)
                        T2[e] *= wr[j];
                    mat_mul<4, 3, 3>(E[j], T2, T3);
                    right_matrix(post[j], M4);
                    mat_mul<4, 4, 3>(M4, T3, T4);
                    left_matrix(pre[j], M4);
                    mat_mul<4, 4, 3>(M4, T4, T3);
                    for (int e = 0; e < 12; ++e)
                        acc[e] += T3[e];
                }
                // Theta_m = 2 [L(q)^T acc]_{0..2}
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                    {
                        double s = 0;
                        for (int l = 0; l < 4; ++l)
                            s += Lq[l * 4 + r] * acc[l * 3 + c];
                        Theta[m * 9 + r * 3 + c] = 2.0 * s;
                    }
            }
        }

        // The pose alone — same expressions for the pose as spline_pose (bit-identical t, q), none of the derivative blocks and none
        // of their local arrays: what a cost-only evaluation and the candidate of a Gauss-Newton step need.
        // For the rotations between neighbouring control knots of a tracker (a few degrees at most) both maps are evaluated by their
        // power series in the SQUARED norm — no sqrt, no atan, no sin / cos, one division: lam = 2 atan(x) / (x w) with x^2 = n2 / w^2 and
        // (sin(t/2) / t, cos(t/2)) in t^2.  The truncation error is below 1e-17 relative inside the thresholds (x^2 <= 2^-8: 9 terms of
        // the alternating series bound it by x^18 / 19 = 1e-23; t^2 <= 2^-6), i.e. at the rounding level of the closed forms that
        // spline_pose keeps (SplineFunctor.h:155-365 via Quaternion.h:61-233); larger rotations take the closed forms here too.
        // This chain is the latency of the candidate's pose records between the two passes of a Gauss-Newton level.
        __host__ __device__ __forceinline__ double atan_over_x_series(double x2)
        {
            // atan(x) / x = 1 - x^2/3 + x^4/5 - ... (Horner in x2)
            double p = 1.0 / 19.0;
            p = fma(-x2, p, 1.0 / 17.0);
            p = fma(-x2, p, 1.0 / 15.0);
            p = fma(-x2, p, 1.0 / 13.0);
            p = fma(-x2, p, 1.0 / 11.0);
            p = fma(-x2, p, 1.0 / 9.0);
            p = fma(-x2, p, 1.0 / 7.0);
            p = fma(-x2, p, 1.0 / 5.0);
            p = fma(-x2, p, 1.0 / 3.0);
            return fma(-x2, p, 1.0);
        }
        __host__ __device__ inline void so3_log_only(const Q &q, double *phi)
        {
            const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
            double lam;
            if (n2 < 1e-20)
                lam = 2. / q.w - 2. / 3. * n2 / (q.w * q.w * q.w);
            else if (q.w > 0.5 && n2 <= 0.00390625 * q.w * q.w)
            {
                const double iw = 1.0 / q.w;
                lam = 2.0 * iw * atan_over_x_series(n2 * iw * iw);
            }
            else
            {
                const double n = sqrt(n2);
                if (fabs(q.w) < 1e-10)
                    lam = (q.w > 0 ? 1.0 : -1.0) * M_PI / n;
                else
                    lam = 2.0 * atan(n / q.w) / n;
            }
            phi[0] = lam * q.x, phi[1] = lam * q.y, phi[2] = lam * q.z;
        }
        __host__ __device__ inline Q so3_exp_only(const double *phi)
        {
            const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
            double fi, fr;
            if (t2 < 1e-20)
            {
                const double t4 = t2 * t2;
                fi = 0.5 - 1. / 48. * t2 + 1. / 3840. * t4;
                fr = 1. - 1. / 8. * t2 + 1. / 384. * t4;
            }
            else if (t2 <= 0.015625)
            {
                // h = (t/2)^2 <= 2^-8:  sin(t/2) / t = (1/2) (1 - h/6 + h^2/120 - ...),  cos(t/2) = 1 - h/2 + h^2/24 - ...
                const double h = 0.25 * t2;
                double s = -1.0 / 6227020800.0;                 // 13!
                s = fma(h, s, 1.0 / 39916800.0);               // 11!
                s = fma(h, s, -1.0 / 362880.0);                // 9!
                s = fma(h, s, 1.0 / 5040.0);
                s = fma(h, s, -1.0 / 120.0);
                s = fma(h, s, 1.0 / 6.0);
                fi = 0.5 * fma(-h, s, 1.0);
                double c = 1.0 / 87178291200.0;                 // 14!
                c = fma(h, c, -1.0 / 479001600.0);             // 12!
                c = fma(h, c, 1.0 / 3628800.0);                // 10!
                c = fma(h, c, -1.0 / 40320.0);
                c = fma(h, c, 1.0 / 720.0);
                c = fma(h, c, -1.0 / 24.0);
                c = fma(h, c, 0.5);
                fr = fma(-h, c, 1.0);
            }
            else
            {
                const double t = sqrt(t2);
                fi = sin(0.5 * t) / t;
                fr = cos(0.5 * t);
            }
            return Q{fi * phi[0], fi * phi[1], fi * phi[2], fr};
        }
        template <int K>
        __host__ __device__ inline void spline_pose_only(const double *kt, const double *kR, double u, double *t_out, Q &q_out, double *wt)
        {
            double wr[K];
            spline_weights<K>(u, wt, wr);
            for (int a = 0; a < 3; ++a)
            {
                double s = 0;
                for (int j = 0; j < K; ++j)
                    s += wt[j] * kt[3 * j + a];
                t_out[a] = s;
            }
            Q q = Q{kR[0], kR[1], kR[2], kR[3]};
#pragma unroll
            for (int j = 1; j < K; ++j)
            {
                const Q a = Q{kR[4 * j - 4], kR[4 * j - 3], kR[4 * j - 2], kR[4 * j - 1]}, b = Q{kR[4 * j], kR[4 * j + 1], kR[4 * j + 2], kR[4 * j + 3]};
                double phi[3];
                so3_log_only(qmul(qconj(a), b), phi);
                phi[0] *= wr[j], phi[1] *= wr[j], phi[2] *= wr[j];
                q = qmul(q, so3_exp_only(phi));
            }
            q_out = q;
        }

        __host__ __device__ __forceinline__ void rotation_matrix(const Q &q, double *R)
        {
            const double x = q.x, y = q.y, z = q.z, w = q.w;
            R[0] = w * w + x * x - y * y - z * z, R[1] = 2 * (x * y - w * z), R[2] = 2 * (x * z + w * y);
            R[3] = 2 * (x * y + w * z), R[4] = w * w - x * x + y * y - z * z, R[5] = 2 * (y * z - w * x);
            R[6] = 2 * (x * z - w * y), R[7] = 2 * (y * z + w * x), R[8] = w * w - x * x - y * y + z * z;
        }

        // correctly rounded, never contracted: the host and the device evaluate the sample time identically
        __host__ __device__ __forceinline__ double add_rn(double a, double b)
        {
#ifdef __CUDA_ARCH__
            return __dadd_rn(a, b);
#else
            volatile double r = a + b;
            return r;
#endif
        }
        __host__ __device__ __forceinline__ double mul_rn(double a, double b)
        {
#ifdef __CUDA_ARCH__
            return __dmul_rn(a, b);
#else
            volatile double r = a * b;
            return r;
#endif
        }
        __host__ __device__ __forceinline__ double div_rn(double a, double b)
        {
#ifdef __CUDA_ARCH__
            return __ddiv_rn(a, b);
#else
            volatile double r = a / b;
            return r;
#endif
        }

        // Sample g = f * N + i of the evaluation: record, mid-exposure pose of its frame, segment ranges.
        // with_jacobian: 0 = pose only (Theta fields zero), 1 = whole record, kPoseJacobianOnly = only the Theta fields of a record
        // whose pose part is already in place (persistent sweep: the two halves are computed at different times)
        constexpr int kPoseJacobianOnly = 2;
        // dbg (nullable): fp64 values BEFORE the rounding to the fp32 record, per sample [t(3) q(4) wt(K) Theta(9K)] with stride
        // kPoseDebugStride doubles (mbavo_debug_dump)
        constexpr int kPoseDebugStride = 7 + 4 + 36;
        // normalised position of sample g inside its spline segment (compute_virtual_camera_poses.cu:33, SplineFunctor.h:13-19);
        // depends on the frame times and the sample count only, not on the knots
        __host__ __device__ inline double sample_u(const EvalStage *st, int g)
        {
            const int N = st->N;
            const int f = g / N, i = g % N;
            const double t_mu = st->exp_time[f];
            const double ts = add_rn(add_rn(st->cap[f], -mul_rn(t_mu, 0.5)), div_rn(mul_rn((double)i, t_mu), (double)(N - 1) + 1e-8));
            const int idx = st->kmin + st->seg_off[g]; // host-computed with the same expression (SplineFunctor.h:13-19)
            return div_rn(add_rn(ts, -st->t0), st->dt) - (double)idx;
        }
        // u_pre (nullable): sample_u of every sample, computed ahead of time (persistent sweep: off the path between two passes)
        template <int K>
        __host__ __device__ inline void pose_one(const EvalStage *st, const double *knots_t, const double *knots_R, int g,
                                                 int with_jacobian, float *samples, double *mid, int *seg_end, double *dbg = nullptr,
                                                 const double *u_pre = nullptr)
        {
            const int N = st->N;
            const int f = g / N, i = g % N;
            const int idx = st->kmin + st->seg_off[g];
            const double u = u_pre ? u_pre[g] : sample_u(st, g);

            double tt[3], wt[K], Theta[K * 9];
            Q q;
            if (with_jacobian)
                spline_pose<K>(knots_t + 3 * idx, knots_R + 4 * idx, u, tt, q, wt, Theta);
            else
                spline_pose_only<K>(knots_t + 3 * idx, knots_R + 4 * idx, u, tt, q, wt);

            if (dbg)
            {
                double *o = dbg + (size_t)g * kPoseDebugStride;
                o[0] = tt[0], o[1] = tt[1], o[2] = tt[2], o[3] = q.x, o[4] = q.y, o[5] = q.z, o[6] = q.w;
                for (int j = 0; j < K; ++j)
                    o[7 + j] = wt[j];
                for (int e = 0; e < 9 * K; ++e)
                    o[11 + e] = with_jacobian ? Theta[e] : 0.0;
            }
            double R[9];
            rotation_matrix(q, R);
            constexpr int REC = sample_rec_floats(K);
            float *rec = samples + (size_t)g * REC;
            if (with_jacobian == kPoseJacobianOnly)
            {
                // second half of a record whose geometry (and blend weights) are already in place: only Theta_j
                for (int j = 0; j < K; ++j)
                {
                    float *c = rec + kRecGeom + 10 * j;
                    const double *Th = Theta + 9 * j;
                    for (int r = 0; r < 3; ++r)
                        c[1 + r] = (float)Th[3 * r], c[4 + 2 * r] = (float)Th[3 * r + 1], c[5 + 2 * r] = (float)Th[3 * r + 2];
                }
                return;
            }
            // R - I, rounded AFTER the subtraction: the tracking kernel works with the small deviation of the warp from
            // the identity so that its fp32 reference coordinates keep ~1e-6 px accuracy (track_kernel.cu, sample_step).
            // Record layout: mbavo_device.h.
            float Rm[9];
            for (int e = 0; e < 9; ++e)
                Rm[e] = (float)(R[e] - ((e & 3) == 0 ? 1.0 : 0.0));
            rec[0] = Rm[0], rec[1] = Rm[1], rec[2] = Rm[3], rec[3] = Rm[4];
            rec[4] = Rm[6], rec[5] = Rm[7], rec[6] = Rm[2], rec[7] = Rm[5];
            rec[8] = Rm[8], rec[9] = (float)tt[2], rec[10] = (float)tt[0], rec[11] = (float)tt[1];
            for (int j = 0; j < K; ++j)
            {
                float *c = rec + kRecGeom + 10 * j;
                const double *Th = Theta + 9 * j;
                c[0] = (float)wt[j];
                for (int r = 0; r < 3; ++r)
                {
                    c[1 + r] = with_jacobian ? (float)Th[3 * r] : 0.f;
                    c[4 + 2 * r] = with_jacobian ? (float)Th[3 * r + 1] : 0.f;
                    c[5 + 2 * r] = with_jacobian ? (float)Th[3 * r + 2] : 0.f;
                }
            }
            for (int e = kRecGeom + 10 * K; e < REC; ++e)
                rec[e] = 0.f;

            if (i == N / 2)
            {
                // compute_local_patches_xy.cu:37-43: R_r2c = R^T, t_r2c = -(R^T t)
                double *m = mid + f * kMidDoubles;
                for (int r = 0; r < 3; ++r)
                {
                    for (int c = 0; c < 3; ++c)
                        m[r * 3 + c] = R[c * 3 + r];
                    m[9 + r] = -(R[0 * 3 + r] * tt[0] + R[1 * 3 + r] * tt[1] + R[2 * 3 + r] * tt[2]);
                }
            }
            // seg_end[f][s] = number of samples of frame f whose segment offset is <= s  (samples are time-ordered)
            const int off = idx - st->kmin;
            const int next_off = (i + 1 < N) ? (int)st->seg_off[g + 1] : kMaxSegments;
            for (int s = off; s < next_off && s < kMaxSegments; ++s)
                seg_end[f * kMaxSegments + s] = i + 1;
            if (i == 0)
                for (int s = 0; s < off && s < kMaxSegments; ++s)
                    seg_end[f * kMaxSegments + s] = 0;
        }
    } // namespace
} // namespace mbavo
#endif
