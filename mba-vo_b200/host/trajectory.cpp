// Host side of the tracker's per-frame trajectory bookkeeping, C++ behind the C-ABI:
//   constant-velocity prediction     src/ba_tracker/blur_aware_direct_tracker.cpp:120-145
//   velocity from neighbouring frames                                            :155-161
//   SplineSE3::GetPose               src/core/common/Spline.h:222-290 (a1: SplineFunctor.h:13-19)
//   SplineSE3::TransformByRight      src/core/common/Spline.h:212-219
//   SplineSE3::TransformTo(t, R, t)  src/core/common/Spline.h:184-201
//   Transformation::exp / log        src/core/states/Transformation.cpp:164-178 — thin wrappers of Sophus::SE3d::exp / log.
// Sophus is a third-party dependency the reference neither vendors nor pins, so the two maps follow its
// published algorithm (se3.hpp / so3.hpp, 1.0.x): the closed forms of V and V^-1 with the first-order branch below
// theta = 1e-10; parity for them is unpinned and the tests check them against the matrix exponential / logarithm.
// Pure host code, microseconds per frame; nothing here touches the device.
#include "../../include/mbavo.h"

#include <cmath>

namespace mbavo
{
    int host_spline_pose(int K, const double *kt, const double *kR, double u, double *t_out, double *q_out);
}

namespace
{
    constexpr double kEps = 1e-10; // Sophus::Constants<double>::epsilon()
    constexpr double kPi = 3.14159265358979323846;

    // Eigen's Quaterniond * Vector3d: v + w uv + q.vec x uv with uv = 2 q.vec x v
    void rotate(const double *q, const double *v, double *out)
    {
        const double ux = 2.0 * (q[1] * v[2] - q[2] * v[1]), uy = 2.0 * (q[2] * v[0] - q[0] * v[2]), uz = 2.0 * (q[0] * v[1] - q[1] * v[0]);
        out[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
        out[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
        out[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
    }

    void qmul(const double *a, const double *b, double *o) // Hamilton product, (x, y, z, w) storage
    {
        const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
        const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
        const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
        const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
        o[0] = x, o[1] = y, o[2] = z, o[3] = w;
    }

    // out = (a I + b hat(w) + c hat(w)^2) v
    void apply_V(double a, double b, double c, const double *w, const double *v, double *out)
    {
        const double wv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
        const double wwv[3] = {w[1] * wv[2] - w[2] * wv[1], w[2] * wv[0] - w[0] * wv[2], w[0] * wv[1] - w[1] * wv[0]};
        for (int i = 0; i < 3; ++i)
            out[i] = a * v[i] + b * wv[i] + c * wwv[i];
    }
} // namespace

extern "C"
{
    int mbavo_se3_exp(const double *tangent, double *t, double *q)
    {
        if (!tangent || !t || !q)
            return MBAVO_EINVAL;
        const double *ups = tangent, *om = tangent + 3; // [translation, rotation] (Transformation.cpp:166, 174)
        const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
        const double theta = std::sqrt(theta_sq);
        double imag, real; // SO3::expAndTheta
        if (theta < kEps)
        {
            const double theta_po4 = theta_sq * theta_sq;
            imag = 0.5 - theta_sq / 48.0 + theta_po4 / 3840.0;
            real = 1.0 - theta_sq / 8.0 + theta_po4 / 384.0;
        }
        else
        {
            imag = std::sin(0.5 * theta) / theta;
            real = std::cos(0.5 * theta);
        }
        q[0] = imag * om[0], q[1] = imag * om[1], q[2] = imag * om[2], q[3] = real;
        if (theta < kEps)
            rotate(q, ups, t); // V = so3.matrix()
        else
            apply_V(1.0, (1.0 - std::cos(theta)) / theta_sq, (theta - std::sin(theta)) / (theta_sq * theta), om, ups, t);
        return MBAVO_OK;
    }

    int mbavo_se3_log(const double *t, const double *q, double *tangent)
    {
        if (!tangent || !t || !q)
            return MBAVO_EINVAL;
        // SO3::logAndTheta
        const double squared_n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
        const double n = std::sqrt(squared_n), w = q[3];
        double two_atan_nbyw_by_n;
        if (n < kEps)
            two_atan_nbyw_by_n = 2.0 / w - 2.0 * squared_n / (w * w * w);
        else if (std::fabs(w) < kEps)
            two_atan_nbyw_by_n = (w > 0.0 ? kPi : -kPi) / n;
        else
            two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
        const double theta = two_atan_nbyw_by_n * n;
        double *om = tangent + 3;
        om[0] = two_atan_nbyw_by_n * q[0], om[1] = two_atan_nbyw_by_n * q[1], om[2] = two_atan_nbyw_by_n * q[2];
        if (std::fabs(theta) < kEps)
            apply_V(1.0, -0.5, 1.0 / 12.0, om, t, tangent);
        else
        {
            const double half = 0.5 * theta;
            apply_V(1.0, -0.5, (1.0 - theta * std::cos(half) / (2.0 * std::sin(half))) / (theta * theta), om, t, tangent);
        }
        return MBAVO_OK;
    }

    int mbavo_spline_pose(const mbavo_spline *sp, double time, double *t, double *q)
    {
        if (!sp || !t || !q || !sp->knots_t || !sp->knots_R || (sp->spline_deg_k != 2 && sp->spline_deg_k != 4) || !(sp->sample_dt > 0))
            return MBAVO_EINVAL;
        const double s = (time - sp->start_time) / sp->sample_dt; // SplineFunctor.h:13-19
        const int idx = (int)s;
        const double u = s - idx;
        if (idx < 0 || idx + sp->spline_deg_k > sp->num_ctrl_knots) // the reference asserts (Spline.h:232-234)
            return MBAVO_EINVAL;
        return mbavo::host_spline_pose(sp->spline_deg_k, sp->knots_t + 3 * idx, sp->knots_R + 4 * idx, u, t, q) == 0 ? MBAVO_OK : MBAVO_EINVAL;
    }

    int mbavo_spline_transform_by_right(int n, double *knots_t, double *knots_R, const double *dq, const double *dt)
    {
        if (n < 1 || !knots_t || !knots_R || !dq || !dt)
            return MBAVO_EINVAL;
        for (int i = 0; i < n; ++i)
        {
            double r[3];
            rotate(knots_R + 4 * i, dt, r);
            for (int a = 0; a < 3; ++a)
                knots_t[3 * i + a] = r[a] + knots_t[3 * i + a];
            qmul(knots_R + 4 * i, dq, knots_R + 4 * i);
        }
        return MBAVO_OK;
    }

    int mbavo_spline_transform_to(const mbavo_spline *sp, double time, const double *target_t, const double *target_q, double *knots_t,
                                  double *knots_R)
    {
        if (!sp || !target_t || !target_q || !knots_t || !knots_R)
            return MBAVO_EINVAL;
        double t0[3], q0[4];
        int rc = mbavo_spline_pose(sp, time, t0, q0);
        if (rc != MBAVO_OK)
            return rc;
        // Eigen's Quaterniond::inverse(): conjugate / squared norm
        const double n2 = q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3];
        if (!(n2 > 0.0))
            return MBAVO_EINVAL;
        const double qi[4] = {-q0[0] / n2, -q0[1] / n2, -q0[2] / n2, q0[3] / n2};
        double dq[4], dt[3];
        qmul(qi, target_q, dq);
        const double d[3] = {target_t[0] - t0[0], target_t[1] - t0[1], target_t[2] - t0[2]};
        rotate(qi, d, dt);
        for (int i = 0; i < 3 * sp->num_ctrl_knots; ++i)
            knots_t[i] = sp->knots_t[i];
        for (int i = 0; i < 4 * sp->num_ctrl_knots; ++i)
            knots_R[i] = sp->knots_R[i];
        return mbavo_spline_transform_by_right(sp->num_ctrl_knots, knots_t, knots_R, dq, dt);
    }

    int mbavo_predict_spline(int n, double *knots_t, double *knots_R, const double *velocity, double dt_frame)
    {
        if (!velocity)
            return MBAVO_EINVAL;
        double tangent[6], t[3], q[4];
        for (int i = 0; i < 6; ++i)
            tangent[i] = velocity[i] * dt_frame; // tracker.cpp:124
        int rc = mbavo_se3_exp(tangent, t, q);   // :137
        if (rc != MBAVO_OK)
            return rc;
        return mbavo_spline_transform_by_right(n, knots_t, knots_R, q, t); // :145
    }

    int mbavo_frame_velocity(const double *prev_t, const double *prev_q, const double *cur_t, const double *cur_q, double dt_frame,
                             double *velocity)
    {
        if (!prev_t || !prev_q || !cur_t || !cur_q || !velocity || !(dt_frame != 0.0))
            return MBAVO_EINVAL;
        // mTprevB2W.inverse() * T_b2w (tracker.cpp:160): rotation q_prev^* q_cur, translation q_prev^* (t_cur - t_prev)
        const double qi[4] = {-prev_q[0], -prev_q[1], -prev_q[2], prev_q[3]};
        double dq[4], dt[3];
        qmul(qi, cur_q, dq);
        const double d[3] = {cur_t[0] - prev_t[0], cur_t[1] - prev_t[1], cur_t[2] - prev_t[2]};
        rotate(qi, d, dt);
        int rc = mbavo_se3_log(dt, dq, velocity); // :161
        for (int i = 0; i < 6; ++i)
            velocity[i] /= dt_frame;
        return rc;
    }
}
