// Synthetic motion-blurred image on the GPU: warp_image + synthesize_motion_blurred_img of
// src/ba_tracker/generate_synthetic_data.cpp:127-180 — every live pixel is warped through the plane Z = D of the keyframe
// with each of the N virtual poses (compute_pixel_intensity<double>, compute_pixel_intensity.h:91-153, intensity only),
// the warped intensity is truncated to 8 bits, the N images are averaged in float and the mean is converted to 8 bits with
// round-to-nearest-even (cv::Mat::convertTo).
//
// This is byte work, so the result has to be BIT-EXACT: the unit is compiled with --fmad=false and follows the
// reference's fp64 operation order (one rounding per operation, as the host compiler evaluates the reference headers),
// including its float sqrt of a double argument (:119), its float reciprocal constant (:137) and its fp32 bilinear
// weights (:43-68).  One thread per pixel, samples in order; the N poses travel as a launch parameter.
#include "../../include/mbavo.h"

#include <cuda_runtime.h>

#include <cstring>

namespace
{
    constexpr int kMaxSynthPoses = 256;
    struct SynthPoses
    {
        double tq[kMaxSynthPoses][7]; // tx ty tz qx qy qz qw
    };

    __device__ __forceinline__ void q_mul(const double *a, const double *b, double *o) // Quaternion.h:44-50
    {
        const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
        const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
        const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
        const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
        o[0] = x, o[1] = y, o[2] = z, o[3] = w;
    }

    // compute_pixel_intensity<double> without the Jacobian; false: the reference coordinate leaves the keyframe
    __device__ bool warp_intensity(const unsigned char *__restrict__ I, int H, int W, const double *t, const double *q, double D,
                                   double fx, double fy, double cx, double cy, double X, double Y, double *out)
    {
        const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
        double xh = (X - cx) / fx, yh = (Y - cy) / fy;
        const double zh = 1. / sqrtf((float)(1. + xh * xh + yh * yh)); // :119
        xh *= zh, yh *= zh;
        const double r20 = 2 * (qx * qz - qw * qy), r21 = 2 * (qy * qz + qw * qx), r22 = qw * qw - qx * qx - qy * qy + qz * qz;
        const double lam = r20 * xh + r21 * yh + r22 * zh; // :124-126
        const double s = (D - t[2]) / lam;                 // :128
        // :130-134: q (s ray, 0) q*
        const double p[4] = {s * xh, s * yh, s * zh, 0.0}, c[4] = {-qx, -qy, -qz, qw};
        double tmp[4], r[4];
        q_mul(q, p, tmp);
        q_mul(tmp, c, r);
        const double Px = r[0] + t[0], Py = r[1] + t[1], Pz = r[2] + t[2];
        const double iz = 1.f / (Pz + 1e-8); // :137
        const double u = fx * (Px * iz) + cx, v = fy * (Py * iz) + cy;
        // bilinear_interpolation, :25-72
        if (u < 0 || u > W - 1 || v < 0 || v > H - 1)
            return false;
        const int xi = (int)u, yi = (int)v;
        const float dx = (float)(u - xi), dy = (float)(v - yi);
        const float dxdy = dx * dy;
        const float w00 = 1.0f - dx - dy + dxdy, w01 = dx - dxdy, w10 = dy - dxdy, w11 = dxdy;
        const size_t last = (size_t)H * W - 1;
        const size_t i00 = (size_t)yi * W + xi, i01 = i00 + 1, i10 = i00 + W, i11 = i10 + 1;
        const size_t j01 = i01 > last ? last : i01, j10 = i10 > last ? last : i10, j11 = i11 > last ? last : i11;
        *out = w11 * I[j11] + w10 * I[j10] + w01 * I[j01] + w00 * I[i00];
        return true;
    }

    __global__ void synth_blur_kernel(const unsigned char *__restrict__ I, int H, int W, double D, double fx, double fy, double cx,
                                      double cy, const __grid_constant__ SynthPoses poses, int N, unsigned char *__restrict__ out)
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
        if (c >= W)
            return;
        float acc = 0.f;
        for (int i = 0; i < N; ++i)
        {
            double v = 0;
            warp_intensity(I, H, W, poses.tq[i], poses.tq[i] + 3, D, fx, fy, cx, cy, (double)c, (double)r, &v);
            acc += (float)(unsigned char)v; // warp_image stores the double into an uchar image (:148), convertTo float (:173)
        }
        const float mean = acc / (float)N;                        // :176
        const int q = __float2int_rn(mean);                       // convertTo CV_8U: round to nearest even, saturate
        out[(size_t)r * W + c] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
    }
} // namespace

extern "C" int mbavo_synthesize_blurred(int device, int mem, const unsigned char *ref_I, int H, int W, double plane_depth, double fx,
                                        double fy, double cx, double cy, const double *poses_tq, int num_poses, unsigned char *out)
{
    if (!ref_I || !poses_tq || !out || H < 2 || W < 2 || num_poses < 1 || num_poses > kMaxSynthPoses ||
        (mem != MBAVO_MEM_HOST && mem != MBAVO_MEM_DEVICE))
        return MBAVO_EINVAL;
    int prev = -1;
    if (device >= 0)
    {
        cudaGetDevice(&prev);
        if (cudaSetDevice(device) != cudaSuccess)
            return MBAVO_ECUDA;
    }
    int rc = MBAVO_OK;
    const size_t npix = (size_t)H * W;
    unsigned char *d_in = nullptr, *d_out = nullptr;
    const unsigned char *src = ref_I;
    unsigned char *dst = out;
    if (mem == MBAVO_MEM_HOST)
    {
        if (cudaMalloc(&d_in, npix) != cudaSuccess || cudaMalloc(&d_out, npix) != cudaSuccess ||
            cudaMemcpy(d_in, ref_I, npix, cudaMemcpyHostToDevice) != cudaSuccess)
            rc = MBAVO_ECUDA;
        src = d_in, dst = d_out;
    }
    if (rc == MBAVO_OK)
    {
        SynthPoses poses; // 14 KB launch parameter
        std::memset(&poses, 0, sizeof poses);
        std::memcpy(poses.tq, poses_tq, sizeof(double) * 7 * num_poses);
        const dim3 block(128, 1, 1), grid((W + 127) / 128, H, 1);
        synth_blur_kernel<<<grid, block>>>(src, H, W, plane_depth, fx, fy, cx, cy, poses, num_poses, dst);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
            rc = MBAVO_ECUDA;
    }
    if (rc == MBAVO_OK && mem == MBAVO_MEM_HOST && cudaMemcpy(out, d_out, npix, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = MBAVO_ECUDA;
    cudaFree(d_in);
    cudaFree(d_out);
    if (prev >= 0)
        cudaSetDevice(prev);
    return rc;
}
