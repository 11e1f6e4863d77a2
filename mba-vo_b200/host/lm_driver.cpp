// Host side of the tracker's Levenberg–Marquardt loop, C++ behind the C-ABI (include/mbavo.h):
//   solve_normal_equation          src/ba_tracker/solve_normal_equation.h:10-35   (Eigen JacobiSVD / LDLT; Eigen is not
//                                  available, so the symmetric system is solved with a cyclic Jacobi eigen-decomposition,
//                                  resp. an LDL^T factorisation, written here)
//   computeTrustRegionStep         src/ba_tracker/blur_aware_direct_tracker.cpp:799-831
//   Plus_t / Plus_R                src/core/common/Spline.h:307-330
//   LevenbergMarquardtStrategy     src/ba_tracker/levenberg_marquardt_strategy.cpp:9-44
//   TrustRegionStepEvaluator       src/ba_tracker/trust_region_step_evaluator.cpp:45-126  (Algorithm 10.1.2, Conn/Gould/Toint)
//   optimizePyramidLevel           src/ba_tracker/blur_aware_direct_tracker.cpp:590-637, 885-924
// The device work is reached only through mbavo_evaluate / mbavo_detect_outliers.
#include "../../include/mbavo.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace
{
    // Symmetric eigen-decomposition A = V diag(w) V^T by cyclic Jacobi rotations (A is n x n, row-major, destroyed).
    void jacobi_eigen(std::vector<double> &A, int n, std::vector<double> &V, std::vector<double> &w)
    {
        V.assign((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i)
            V[(size_t)i * n + i] = 1.0;
        for (int sweep = 0; sweep < 64; ++sweep)
        {
            double off = 0, diag = 0;
            for (int i = 0; i < n; ++i)
            {
                diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
                for (int j = i + 1; j < n; ++j)
                    off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
            }
            if (off <= 1e-60 || off <= 1e-32 * diag)
                break;
            for (int p = 0; p < n - 1; ++p)
                for (int q = p + 1; q < n; ++q)
                {
                    const double apq = A[(size_t)p * n + q];
                    if (apq == 0.0)
                        continue;
                    const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < n; ++k)
                    {
                        const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                        A[(size_t)k * n + p] = c * akp - s * akq;
                        A[(size_t)k * n + q] = s * akp + c * akq;
                    }
                    for (int k = 0; k < n; ++k)
                    {
                        const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                        A[(size_t)p * n + k] = c * apk - s * aqk;
                        A[(size_t)q * n + k] = s * apk + c * aqk;
                    }
                    for (int k = 0; k < n; ++k)
                    {
                        const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                        V[(size_t)k * n + p] = c * vkp - s * vkq;
                        V[(size_t)k * n + q] = s * vkp + c * vkq;
                    }
                }
        }
        w.resize(n);
        for (int i = 0; i < n; ++i)
            w[i] = A[(size_t)i * n + i];
    }

    // x = A^-1 b through the eigen-decomposition; singular directions (|w| <= n eps |w|max) are dropped like
    // JacobiSVD::solve does with its default threshold.
    void solve_svd(const double *A, const double *b, int n, double *x)
    {
        std::vector<double> M(A, A + (size_t)n * n), V, w;
        jacobi_eigen(M, n, V, w);
        double wmax = 0;
        for (double v : w)
            wmax = std::max(wmax, std::fabs(v));
        const double tol = std::numeric_limits<double>::epsilon() * n * wmax;
        std::vector<double> y(n, 0.0);
        for (int i = 0; i < n; ++i)
        {
            if (std::fabs(w[i]) <= tol)
                continue;
            double s = 0;
            for (int k = 0; k < n; ++k)
                s += V[(size_t)k * n + i] * b[k];
            y[i] = s / w[i];
        }
        for (int k = 0; k < n; ++k)
        {
            double s = 0;
            for (int i = 0; i < n; ++i)
                s += V[(size_t)k * n + i] * y[i];
            x[k] = s;
        }
    }

    // x = A^-1 b by LDL^T (A symmetric, row-major).  min_pivot_ratio > 0: give up (return false) when the smallest
    // pivot falls below that fraction of the largest, or is not positive.
    bool solve_ldlt(const double *A, const double *b, int n, double *x, double min_pivot_ratio = 0.0)
    {
        std::vector<double> L((size_t)n * n, 0.0), D(n, 0.0);
        double dmax = 0.0;
        for (int j = 0; j < n; ++j)
        {
            double d = A[(size_t)j * n + j];
            for (int k = 0; k < j; ++k)
                d -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
            D[j] = d;
            if (d == 0.0)
                return false;
            if (min_pivot_ratio > 0.0)
            {
                dmax = std::max(dmax, d);
                if (!(d > min_pivot_ratio * dmax))
                    return false;
            }
            L[(size_t)j * n + j] = 1.0;
            for (int i = j + 1; i < n; ++i)
            {
                double s = A[(size_t)i * n + j];
                for (int k = 0; k < j; ++k)
                    s -= L[(size_t)i * n + k] * L[(size_t)j * n + k] * D[k];
                L[(size_t)i * n + j] = s / d;
            }
        }
        std::vector<double> y(b, b + n);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < i; ++k)
                y[i] -= L[(size_t)i * n + k] * y[k];
        for (int i = 0; i < n; ++i)
            y[i] /= D[i];
        for (int i = n - 1; i >= 0; --i)
            for (int k = i + 1; k < n; ++k)
                y[i] -= L[(size_t)k * n + i] * y[k];
        std::copy(y.begin(), y.end(), x);
        return true;
    }

    // Hamilton product, (x, y, z, w)                                                        Quaternion.h:44-50
    void q_mul(const double *a, const double *b, double *o)
    {
        const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
        const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
        const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
        const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
        o[0] = x, o[1] = y, o[2] = z, o[3] = w;
    }

    // Sophus::SO3d::exp(w).unit_quaternion() (Spline.h:326): half-angle map with a Taylor branch for tiny angles
    void so3_exp(const double *w, double *q)
    {
        const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
        double fi, fr;
        if (t2 < 1e-20)
        {
            const double t4 = t2 * t2;
            fi = 0.5 - t2 / 48.0 + t4 / 3840.0;
            fr = 1.0 - t2 / 8.0 + t4 / 384.0;
        }
        else
        {
            const double t = std::sqrt(t2);
            fi = std::sin(0.5 * t) / t;
            fr = std::cos(0.5 * t);
        }
        q[0] = fi * w[0], q[1] = fi * w[1], q[2] = fi * w[2], q[3] = fr;
    }

    struct LmStrategy // levenberg_marquardt_strategy.cpp:9-44
    {
        double radius = 1e4, min_radius = 10, max_radius = 1e32, decrease = 2.0;
        void accepted(double q)
        {
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * q - 1.0, 3));
            radius = std::max(std::min(max_radius, radius), min_radius);
            decrease = 2.0;
        }
        void rejected()
        {
            radius = radius / decrease;
            radius = std::max(std::min(max_radius, radius), min_radius);
            decrease *= 2.0;
        }
    };

    struct StepEvaluator // trust_region_step_evaluator.cpp:45-126
    {
        int max_nonmono;
        double minimum, current, reference, candidate, acc_ref = 0, acc_cand = 0;
        int n_nonmono = 0;
        StepEvaluator(int m, double c) : max_nonmono(m), minimum(c), current(c), reference(c), candidate(c) {}
        double quality(double cost, double model) const
        {
            if (cost >= std::numeric_limits<double>::max())
                return std::numeric_limits<double>::lowest();
            const double rel = (current - cost) / model;
            const double hist = (reference - cost) / (acc_ref + model);
            return std::max(rel, hist);
        }
        void accepted(double cost, double model)
        {
            current = cost;
            acc_cand += model;
            acc_ref += model;
            if (current < minimum)
            {
                minimum = current;
                n_nonmono = 0;
                candidate = current;
                acc_cand = 0.0;
            }
            else
            {
                ++n_nonmono;
                if (current > candidate)
                {
                    candidate = current;
                    acc_cand = 0.0;
                }
            }
            if (n_nonmono == max_nonmono)
            {
                reference = candidate;
                acc_ref = acc_cand;
            }
        }
    };
} // namespace

extern "C" int mbavo_gn_sweep_device(mbavo_ctx *ctx, int level_coarse, int level_fine, int chain, int k, double t0, double dt, int n,
                                     double *knots_t, double *knots_R, double radius, double huber_a, double *costs);

extern "C"
{
    int mbavo_trust_region_step(double *H, const double *g, int dim, double radius, int solver_type, double *step,
                                double *model_cost_change)
    {
        if (!H || !g || !step || !model_cost_change || dim < 1 || !(radius > 0))
            return MBAVO_EINVAL;
        for (int i = 0; i < dim; ++i) // in place: the damping compounds over rejected steps (tracker.cpp:803)
            H[(size_t)i * dim + i] += H[(size_t)i * dim + i] * (1.0 / radius);
        if (solver_type == MBAVO_SOLVER_SVD_JACOBI)
        {
            // JacobiSVD::solve returns the pseudo-inverse solution, which IS the inverse solution whenever H has full
            // numerical rank.  A positive-definite H whose LDL^T pivots stay within 1e-9 of the largest (cond < ~1e9,
            // far from the SVD's rank threshold n eps) is solved by LDL^T (~1 us for 12 x 12 instead of ~40 us); anything
            // else goes through the eigen-decomposition, which drops singular directions like the SVD does.
            if (!solve_ldlt(H, g, dim, step, 1e-9))
                solve_svd(H, g, dim, step);
        }
        else if (solver_type == MBAVO_SOLVER_LDLT)
        {
            if (!solve_ldlt(H, g, dim, step))
                return MBAVO_EINVAL;
        }
        else
            return MBAVO_EINVAL;
        double gs = 0, sHs = 0;
        for (int i = 0; i < dim; ++i)
        {
            step[i] = -step[i]; // solve_normal_equation.h:33
            gs += g[i] * step[i];
        }
        for (int i = 0; i < dim; ++i)
        {
            double r = 0;
            for (int j = 0; j < dim; ++j)
                r += H[(size_t)i * dim + j] * step[j];
            sHs += step[i] * r;
        }
        *model_cost_change = -(gs + 0.5 * sHs); // tracker.cpp:821-823
        return MBAVO_OK;
    }

    int mbavo_spline_plus(int n, const double *kt, const double *kR, const double *step, double *ct, double *cR)
    {
        if (n < 1 || !kt || !kR || !step || !ct || !cR)
            return MBAVO_EINVAL;
        for (int i = 0; i < 3 * n; ++i)
            ct[i] = kt[i] + step[i];
        for (int i = 0; i < n; ++i)
        {
            double dq[4];
            so3_exp(step + 3 * n + 3 * i, dq);
            q_mul(kR + 4 * i, dq, cR + 4 * i);
        }
        return MBAVO_OK;
    }

    void mbavo_lm_default_options(mbavo_lm_options *opt)
    {
        if (!opt)
            return;
        opt->max_num_iterations = 50;
        opt->min_step_quality = 0.5;
        opt->min_abs_cost_decrease = 1e-3;
        opt->solver_type = MBAVO_SOLVER_SVD_JACOBI;
        opt->max_consecutive_nonmonotonic_steps = 5;
        opt->max_chi_square_error = 3.0;
        opt->huber_a = 10.0;
    }

    int mbavo_gn_iteration(mbavo_ctx *ctx, int level, int k, double t0, double dt, int n, const double *knots_t,
                           const double *knots_R, double radius, double huber_a, int solver_type, double *cost,
                           double *candidate_cost, double *step_out, double *cand_t, double *cand_R)
    {
        if (!ctx || !knots_t || !knots_R || !cost || !candidate_cost || n < 2 || n > 16)
            return MBAVO_EINVAL;
        const int dim = 6 * n;
        double H[96 * 96], g[96], step[96], ct[3 * 16], cR[4 * 16], model = 0;
        mbavo_spline sp{k, t0, dt, n, knots_t, knots_R};
        int rc = mbavo_evaluate(ctx, level, &sp, huber_a, cost, H, g);
        if (rc != MBAVO_OK)
            return rc;
        rc = mbavo_trust_region_step(H, g, dim, radius, solver_type, step, &model);
        if (rc != MBAVO_OK)
            return rc;
        mbavo_spline_plus(n, knots_t, knots_R, step, ct, cR);
        mbavo_spline cand{k, t0, dt, n, ct, cR};
        rc = mbavo_evaluate(ctx, level, &cand, huber_a, candidate_cost, nullptr, nullptr);
        if (rc != MBAVO_OK)
            return rc;
        if (step_out)
            std::memcpy(step_out, step, sizeof(double) * dim);
        if (cand_t)
            std::memcpy(cand_t, ct, sizeof(double) * 3 * n);
        if (cand_R)
            std::memcpy(cand_R, cR, sizeof(double) * 4 * n);
        return MBAVO_OK;
    }

    int mbavo_gn_sweep(mbavo_ctx *ctx, int level_coarse, int level_fine, int chain, int k, double t0, double dt, int n,
                       double *knots_t, double *knots_R, double radius, double huber_a, int solver_type, double *costs)
    {
        if (!ctx || !knots_t || !knots_R || level_coarse < level_fine || level_fine < 0 || level_coarse >= MBAVO_MAX_LEVELS ||
            n < 2 || n > 16)
            return MBAVO_EINVAL;
        // fast path: the whole sweep on the device, one wait (mbavo_api.cu); it declines (returns 1) when it cannot
        // reproduce the host semantics, e.g. normal equations that need the SVD branch
        if (solver_type == MBAVO_SOLVER_SVD_JACOBI || solver_type == MBAVO_SOLVER_LDLT)
        {
            std::vector<double> kt(knots_t, knots_t + 3 * n), kR(knots_R, knots_R + 4 * n);
            const int rc = mbavo_gn_sweep_device(ctx, level_coarse, level_fine, chain, k, t0, dt, n, kt.data(), kR.data(), radius,
                                                 huber_a, costs);
            if (rc == MBAVO_OK)
            {
                std::memcpy(knots_t, kt.data(), sizeof(double) * 3 * n);
                std::memcpy(knots_R, kR.data(), sizeof(double) * 4 * n);
                return MBAVO_OK;
            }
            if (rc < 0)
                return rc;
        }
        double ct[3 * 16], cR[4 * 16];
        for (int level = level_coarse; level >= level_fine; --level) // optimizeTrajectory, tracker.cpp:571-575
        {
            double c = 0, cc = 0;
            int rc = mbavo_gn_iteration(ctx, level, k, t0, dt, n, knots_t, knots_R, radius, huber_a, solver_type, &c, &cc, nullptr, ct, cR);
            if (rc != MBAVO_OK)
                return rc;
            if (costs)
                costs[2 * (level_coarse - level)] = c, costs[2 * (level_coarse - level) + 1] = cc;
            if (chain && cc < c) // the finer level starts from the candidate if it lowered the cost
            {
                std::memcpy(knots_t, ct, sizeof(double) * 3 * n);
                std::memcpy(knots_R, cR, sizeof(double) * 4 * n);
            }
        }
        return MBAVO_OK;
    }

    int mbavo_optimize_level(mbavo_ctx *ctx, int level, int k, double t0, double dt, int n, double *knots_t, double *knots_R,
                             const mbavo_lm_options *opt, mbavo_lm_summary *sum)
    {
        if (!ctx || !knots_t || !knots_R || !opt || n < 2 || n > 16)
            return MBAVO_EINVAL;
        const int dim = 6 * n;
        std::vector<double> H((size_t)dim * dim), g(dim), step(dim), ct(3 * n), cR(4 * n);
        mbavo_lm_summary local;
        mbavo_lm_summary &S = sum ? *sum : local;
        std::memset(&S, 0, sizeof S);

        int rc = mbavo_set_outliers(ctx, level, nullptr, 0); // tracker.cpp:600-601
        if (rc != MBAVO_OK)
            return rc;
        mbavo_spline sp{k, t0, dt, n, knots_t, knots_R};
        double eval_cost = 0, cand_cost = 0;
        rc = mbavo_evaluate(ctx, level, &sp, opt->huber_a, &eval_cost, H.data(), g.data()); // iteration 0, :604
        if (rc != MBAVO_OK)
            return rc;
        S.num_evaluations = 1;
        S.initial_cost = eval_cost;
        LmStrategy lm;
        StepEvaluator ev(opt->max_consecutive_nonmonotonic_steps, eval_cost);
        int iter = 0, ndec = 0;
        double abs_decrease = 1e10;
        bool have_first = false;
        for (;;)
        {
            ++iter; // finalizeIterationAndCheckIfMinimizerCanContinue, :910-924
            if (iter > opt->max_num_iterations || abs_decrease < opt->min_abs_cost_decrease)
                break;
            double model = 0;
            rc = mbavo_trust_region_step(H.data(), g.data(), dim, lm.radius, opt->solver_type, step.data(), &model);
            if (rc != MBAVO_OK)
                return rc;
            if (!have_first)
            {
                std::memcpy(S.first_step, step.data(), sizeof(double) * dim);
                have_first = true;
            }
            if (model < 0) // :825-829 -> handleInvalidStep
            {
                lm.rejected();
                ++S.num_invalid;
                if (ndec < 63)
                    S.decisions[ndec++] = 'I';
                continue;
            }
            mbavo_spline_plus(n, knots_t, knots_R, step.data(), ct.data(), cR.data());
            mbavo_spline cand{k, t0, dt, n, ct.data(), cR.data()};
            rc = mbavo_evaluate(ctx, level, &cand, opt->huber_a, &cand_cost, nullptr, nullptr);
            if (rc != MBAVO_OK)
                return rc;
            ++S.num_evaluations;
            abs_decrease = eval_cost - cand_cost;
            const double q = ev.quality(cand_cost, model);
            if (q > opt->min_step_quality && cand_cost < eval_cost) // isStepSuccessful, :890-894
            {
                int nbad = 0;
                rc = mbavo_detect_outliers(ctx, level, opt->max_chi_square_error, &nbad); // :624
                if (rc != MBAVO_OK)
                    return rc;
                S.num_bad_keypoints = nbad;
                std::memcpy(knots_t, ct.data(), sizeof(double) * 3 * n); // InvalidParameter, Spline.h:332-341
                std::memcpy(knots_R, cR.data(), sizeof(double) * 4 * n);
                rc = mbavo_evaluate(ctx, level, &sp, opt->huber_a, &eval_cost, H.data(), g.data());
                if (rc != MBAVO_OK)
                    return rc;
                ++S.num_evaluations;
                lm.accepted(q);
                ev.accepted(eval_cost, model);
                ++S.num_accepted;
                if (ndec < 63)
                    S.decisions[ndec++] = 'A';
            }
            else
            {
                lm.rejected();
                ++S.num_rejected;
                if (ndec < 63)
                    S.decisions[ndec++] = 'R';
            }
        }
        S.decisions[ndec] = 0;
        S.num_iterations = iter - 1;
        S.final_cost = eval_cost;
        return MBAVO_OK;
    }
}
