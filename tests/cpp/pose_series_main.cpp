// Host-side check of the pose-only path the persistent sweep uses for a candidate's sample records (csrc/pose_device.cuh:
// so3_log_only / so3_exp_only / spline_pose_only — power series in the squared norm for the small rotations between
// neighbouring control knots) against the closed forms spline_pose keeps (the restatement of SplineFunctor.h:155-365 via
// Quaternion.h:61-233) and against long-double evaluations of log / exp.  Compiled as host code (nvcc -x cu), no GPU needed.
// Prints one line per check: name, maximum error.  tests/test_host_abi.py::test_pose_only_series_host asserts on them.
#include "../../mba-vo_b200/csrc/pose_device.cuh"

#include <cmath>
#include <cstdio>
#include <random>

using namespace mbavo;

static Q unit_quat_from(const double *phi)
{
    // long-double closed form of Exp(phi)
    const long double t = std::sqrt((long double)phi[0] * phi[0] + (long double)phi[1] * phi[1] + (long double)phi[2] * phi[2]);
    const long double fi = t < 1e-30L ? 0.5L : std::sin(0.5L * t) / t, fr = std::cos(0.5L * t);
    return Q{(double)(fi * phi[0]), (double)(fi * phi[1]), (double)(fi * phi[2]), (double)fr};
}

int main()
{
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    double e_exp = 0, e_log = 0, e_pose_t = 0, e_pose_q = 0, e_round = 0, e_u = 0;
    // rotation magnitudes across every branch of the two maps: tiny (t^2 < 1e-20), series (t^2 <= 2^-6), closed form beyond
    const double mags[] = {0.0, 1e-12, 1e-9, 1e-6, 1e-3, 0.01, 0.05, 0.1, 0.12, 0.1249, 0.1251, 0.2, 0.5, 1.0, 2.0};
    for (double mag : mags)
        for (int rep = 0; rep < 200; ++rep)
        {
            double ax[3] = {U(rng), U(rng), U(rng)};
            const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) + 1e-300;
            const double phi[3] = {mag * ax[0] / n, mag * ax[1] / n, mag * ax[2] / n};
            const Q want = unit_quat_from(phi);
            const Q got = so3_exp_only(phi);
            e_exp = std::fmax(e_exp, std::fmax(std::fmax(std::fabs(got.x - want.x), std::fabs(got.y - want.y)),
                                               std::fmax(std::fabs(got.z - want.z), std::fabs(got.w - want.w))));
            double back[3];
            so3_log_only(want, back);
            for (int a = 0; a < 3; ++a)
                e_log = std::fmax(e_log, std::fabs(back[a] - phi[a]));
            // log(exp(phi)) with the product's own exp: the round trip the spline makes
            so3_log_only(got, back);
            for (int a = 0; a < 3; ++a)
                e_round = std::fmax(e_round, std::fabs(back[a] - phi[a]));
        }
    // whole poses: linear (k = 2) and cubic (k = 4) segments, knots a few degrees apart, against the closed-form path
    for (int rep = 0; rep < 2000; ++rep)
    {
        double kt[12], kR[16];
        const double start[3] = {0.3 * U(rng), 0.3 * U(rng), 0.3 * U(rng)};
        Q q = unit_quat_from(start);
        for (int j = 0; j < 4; ++j)
        {
            for (int a = 0; a < 3; ++a)
                kt[3 * j + a] = U(rng);
            kR[4 * j] = q.x, kR[4 * j + 1] = q.y, kR[4 * j + 2] = q.z, kR[4 * j + 3] = q.w;
            const double step[3] = {0.05 * U(rng), 0.05 * U(rng), 0.05 * U(rng)};
            q = qmul(q, unit_quat_from(step));
        }
        const double u = 0.5 * (U(rng) + 1.0);
        double t1[3], t2[3], wt1[4], wt2[4];
        Q q1, q2;
        spline_pose<2>(kt, kR, u, t1, q1, wt1, nullptr);
        spline_pose_only<2>(kt, kR, u, t2, q2, wt2);
        for (int a = 0; a < 3; ++a)
            e_pose_t = std::fmax(e_pose_t, std::fabs(t1[a] - t2[a]));
        e_pose_q = std::fmax(e_pose_q, std::fmax(std::fmax(std::fabs(q1.x - q2.x), std::fabs(q1.y - q2.y)),
                                                 std::fmax(std::fabs(q1.z - q2.z), std::fabs(q1.w - q2.w))));
        spline_pose<4>(kt, kR, u, t1, q1, wt1, nullptr);
        spline_pose_only<4>(kt, kR, u, t2, q2, wt2);
        for (int a = 0; a < 3; ++a)
            e_pose_t = std::fmax(e_pose_t, std::fabs(t1[a] - t2[a]));
        e_pose_q = std::fmax(e_pose_q, std::fmax(std::fmax(std::fabs(q1.x - q2.x), std::fabs(q1.y - q2.y)),
                                                 std::fmax(std::fabs(q1.z - q2.z), std::fabs(q1.w - q2.w))));
    }
    // sample_u: the position of every exposure sample inside its segment, against the plain expression
    // (compute_virtual_camera_poses.cu:33, SplineFunctor.h:13-19)
    {
        EvalStage st{};
        st.N = 32, st.F = 2, st.kmin = 1, st.t0 = 0.25, st.dt = 0.0125;
        st.cap[0] = 0.3, st.exp_time[0] = 0.02, st.cap[1] = 0.34, st.exp_time[1] = 0.03;
        for (int g = 0; g < st.N * st.F; ++g)
        {
            const int f = g / st.N, i = g % st.N;
            const double ts = st.cap[f] - st.exp_time[f] * 0.5 + (double)i * st.exp_time[f] / ((double)(st.N - 1) + 1e-8);
            const int seg = (int)std::floor((ts - st.t0) / st.dt);
            st.seg_off[g] = (unsigned char)(seg - st.kmin);
            e_u = std::fmax(e_u, std::fabs(sample_u(&st, g) - ((ts - st.t0) / st.dt - (double)seg)));
        }
    }
    std::printf("exp_only_vs_long_double %.3e\n", e_exp);
    std::printf("log_only_vs_long_double %.3e\n", e_log);
    std::printf("log_exp_round_trip %.3e\n", e_round);
    std::printf("pose_only_vs_closed_form_t %.3e\n", e_pose_t);
    std::printf("pose_only_vs_closed_form_q %.3e\n", e_pose_q);
    std::printf("sample_u_vs_plain %.3e\n", e_u);
    return 0;
}
