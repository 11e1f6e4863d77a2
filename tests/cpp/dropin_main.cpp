// Drives the drop-in shim (mba-vo_b200/host/spline_update_step.h) exactly the way the reference tracker drives the
// reference API: storages initialised once, inputs poked into the storages with raw cudaMemcpy
// (blur_aware_direct_tracker.cpp:701-763), evaluate_cost_hessian_gradient for the Hessian pass and the cost-only pass
// (:753-797, 833-883), patch costs read back with the reference's stride (:646-657).
// Input: a flat binary problem file written by tests/test_gpu_dropin.py; output: JSON on stdout.
#include "../../mba-vo_b200/host/spline_update_step.h"

#include <cuda_runtime_api.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace SLAM;

template <class T>
static std::vector<T> read_vec(FILE *f, size_t n)
{
    std::vector<T> v(n);
    if (fread(v.data(), sizeof(T), n, f) != n)
    {
        std::fprintf(stderr, "short read\n");
        std::exit(2);
    }
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 2)
        return 1;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f)
        return 1;
    auto hdr = read_vec<int>(f, 8); // H, W, P, S, N, n_knots, k, num_flagged
    const int H = hdr[0], W = hdr[1], P = hdr[2], S = hdr[3], N = hdr[4], n = hdr[5], k = hdr[6], nbad = hdr[7];
    auto dbl = read_vec<double>(f, 9); // fx, fy, cx, cy, cap, exp, t0, dt, huber
    auto ref_I = read_vec<unsigned char>(f, (size_t)H * W);
    auto dIxy = read_vec<float>(f, (size_t)H * W * 2);
    auto cur_I = read_vec<unsigned char>(f, (size_t)H * W);
    auto xy = read_vec<double>(f, (size_t)P * 2);
    auto z = read_vec<double>(f, P);
    auto pattern = read_vec<int>(f, (size_t)S * 2);
    auto kt = read_vec<double>(f, 3 * n);
    auto kR = read_vec<double>(f, 4 * n);
    auto flags = read_vec<unsigned char>(f, P);
    std::fclose(f);

    VO::CudaSharedStorages st;
    VO::initialize_shared_cuda_storages(1, 64, P, 128, 16, k, st);

    unsigned char *d_ref, *d_cur;
    float *d_g;
    cudaMalloc((void **)&d_ref, ref_I.size());
    cudaMalloc((void **)&d_cur, cur_I.size());
    cudaMalloc((void **)&d_g, dIxy.size() * sizeof(float));
    cudaMemcpy(d_ref, ref_I.data(), ref_I.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cur, cur_I.data(), cur_I.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_g, dIxy.data(), dIxy.size() * sizeof(float), cudaMemcpyHostToDevice);

    // uploadDataToGpu() / uploadDataToGpu(level)
    cudaMemcpy(st.cuda_img_cap_time, &dbl[4], sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(st.cuda_img_exp_time, &dbl[5], sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(st.cuda_cur_images, &d_cur, sizeof(void *), cudaMemcpyHostToDevice);
    std::vector<Core::Vector2d> kp(P);
    for (int p = 0; p < P; ++p)
        kp[p] = Core::Vector2d(xy[2 * p], xy[2 * p + 1]);
    cudaMemcpy(st.cuda_keypoint_xy, kp.data(), sizeof(Core::Vector2d) * P, cudaMemcpyHostToDevice);
    cudaMemcpy(st.cuda_keypoint_depth_z, z.data(), sizeof(double) * P, cudaMemcpyHostToDevice);
    cudaMemcpy(st.cuda_local_patch_pattern_xy, pattern.data(), sizeof(int) * 2 * S, cudaMemcpyHostToDevice);
    // evaluateCostGradientAndHessian()
    cudaMemcpy(st.cuda_spline_ctrl_knots_data_t, kt.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(st.cuda_spline_ctrl_knots_data_R, kR.data(), sizeof(double) * 4 * n, cudaMemcpyHostToDevice);

    Core::VectorX<double, 4> K;
    K.values[0] = dbl[0], K.values[1] = dbl[1], K.values[2] = dbl[2], K.values[3] = dbl[3];
    Core::VectorX<int, 2> HW;
    HW.values[0] = H, HW.values[1] = W;
    int seg = 0;
    const int dim = 6 * n;
    std::vector<double> Hm((size_t)dim * dim), g(dim);
    double cost = 0, cost_only = 0, cost_flagged = 0;
    VO::evaluate_cost_hessian_gradient(N, 1, d_ref, d_g, P, S, K, HW, k, dbl[6], dbl[7], &seg, n, st, dbl[8], &cost, Hm.data(),
                                       g.data());
    VO::evaluate_cost_hessian_gradient(N, 1, d_ref, d_g, P, S, K, HW, k, dbl[6], dbl[7], &seg, n, st, dbl[8], &cost_only, nullptr,
                                       nullptr);
    // detectOutliers read-back: P * E doubles, element [i * E] is the patch cost
    const int ndim = 6 * k + 1, E = ndim * (ndim + 1) / 2;
    std::vector<double> patch((size_t)P * E);
    cudaMemcpy(patch.data(), st.cuda_patch_cost_gradient_hessian_tR, sizeof(double) * P * E, cudaMemcpyDeviceToHost);
    // upload outlier flags the way the tracker does (:696-698) and evaluate again
    cudaMemcpy(st.cuda_keypoints_outlier_flags, flags.data(), P, cudaMemcpyHostToDevice);
    st.num_bad_keypoints = nbad;
    VO::evaluate_cost_hessian_gradient(N, 1, d_ref, d_g, P, S, K, HW, k, dbl[6], dbl[7], &seg, n, st, dbl[8], &cost_flagged, nullptr,
                                       nullptr);

    // ---- the same call with the opt-in level cache (CudaSharedStorages::mbavo_texel_cache), and wall-clock per evaluation both ways
    cudaMemset(st.cuda_keypoints_outlier_flags, 0, P);
    st.num_bad_keypoints = 0;
    auto now_us = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const int reps = 50;
    double cost_again = 0, cost_cached = 0, cost_epoch = 0, us_plain_h, us_plain_c, us_cached_h, us_cached_c;
    std::vector<double> Hc((size_t)dim * dim), gc(dim);
    auto eval = [&](bool with_h, double *c, double *Hout, double *gout) {
        VO::evaluate_cost_hessian_gradient(N, 1, d_ref, d_g, P, S, K, HW, k, dbl[6], dbl[7], &seg, n, st, dbl[8], c, with_h ? Hout : nullptr,
                                           with_h ? gout : nullptr);
    };
    eval(true, &cost_again, Hc.data(), gc.data());
    double t0 = now_us();
    for (int i = 0; i < reps; ++i)
        eval(true, &cost_again, Hc.data(), gc.data());
    us_plain_h = (now_us() - t0) / reps;
    t0 = now_us();
    for (int i = 0; i < reps; ++i)
        eval(false, &cost_again, nullptr, nullptr);
    us_plain_c = (now_us() - t0) / reps;
    st.mbavo_texel_cache = 1;
    eval(true, &cost_cached, Hc.data(), gc.data()); // fills the cache
    t0 = now_us();
    for (int i = 0; i < reps; ++i)
        eval(true, &cost_cached, Hc.data(), gc.data());
    us_cached_h = (now_us() - t0) / reps;
    t0 = now_us();
    for (int i = 0; i < reps; ++i)
        eval(false, &cost_epoch, nullptr, nullptr);
    us_cached_c = (now_us() - t0) / reps;
    double h_diff = 0;
    for (int i = 0; i < dim * dim; ++i)
        h_diff = std::max(h_diff, std::fabs(Hc[i] - Hm[i]));
    // new keyframe content behind the same pointer: the caller bumps the epoch and the level is re-derived
    std::vector<unsigned char> ref2(ref_I);
    for (auto &b : ref2)
        b = (unsigned char)(255 - b);
    cudaMemcpy(d_ref, ref2.data(), ref2.size(), cudaMemcpyHostToDevice);
    double cost_stale = 0;
    eval(false, &cost_stale, nullptr, nullptr);       // still the cached texels: the old keyframe
    ++st.mbavo_keyframe_epoch;
    eval(false, &cost_epoch, nullptr, nullptr);       // re-derived: differs
    cudaMemcpy(d_ref, ref_I.data(), ref_I.size(), cudaMemcpyHostToDevice);
    st.mbavo_texel_cache = 0;

    std::printf("{\"shim_us\": {\"hessian\": %.2f, \"cost\": %.2f, \"hessian_cached\": %.2f, \"cost_cached\": %.2f}, "
                "\"cost_cached\": %.17g, \"H_cached_max_abs_diff\": %.3g, \"cost_stale\": %.17g, \"cost_new_epoch\": %.17g, ",
                us_plain_h, us_plain_c, us_cached_h, us_cached_c, cost_cached, h_diff, cost_stale, cost_epoch);
    std::printf("\"cost\": %.17g, \"cost_only\": %.17g, \"cost_flagged\": %.17g, \"g\": [", cost, cost_only, cost_flagged);
    for (int i = 0; i < dim; ++i)
        std::printf("%s%.17g", i ? ", " : "", g[i]);
    std::printf("], \"H\": [");
    for (int i = 0; i < dim * dim; ++i)
        std::printf("%s%.17g", i ? ", " : "", Hm[i]);
    std::printf("], \"patch_costs\": [");
    for (int p = 0; p < P; ++p)
        std::printf("%s%.17g", p ? ", " : "", patch[(size_t)p * E]);
    std::printf("]}\n");
    VO::free_shared_cuda_storages(st);
    cudaFree(d_ref), cudaFree(d_cur), cudaFree(d_g);
    return 0;
}
