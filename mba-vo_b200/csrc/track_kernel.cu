// Fused blur-aware photometric tracking kernel for one pyramid level.
//
// Replaces, in ONE launch, the reference's kernel chain (src/ba_tracker):
//   kernel_compute_local_patches_xy              compute_local_patches_xy.cu:9-50
//   kernel_compute_pixel_jacobian_residual       compute_hessian_gradients_cost.cu:23-156   (+ compute_pixel_intensity.h:25-209)
//   kernel_compute_patch_cost_gradient_hessian   compute_hessian_gradients_cost.cu:165-239
//   kernel_compute_frame_cost_gradient_hessian   compute_hessian_gradients_cost.cu:247-283
// without the P*S*N*d fp64 scratch, the (d+1) block barriers per pixel and the E barriers per patch.
//
// Mapping.  A warp owns a batch of TP whole host-map points of one frame and works through it in chunks of 32 residual
// pixels:
//   A  lane = pixel: patch centre (fp64), integer live pixel, ray -> 8-word pixel record in shared memory;
//   B  lane = (pixel slot, exposure phase): the 32 lanes are PH phases x Q = 32/PH pixel slots.  Q is the patch size
//      (8 for the reference pattern), so at any instruction the warp gathers around ONE host-map point: the taps of a
//      load instruction fall into the handful of 128-byte lines the patch covers instead of the ~20 that 32 pixels of
//      4 different points touch — the L1 data pipe (1 wavefront / clk / SM) was the limiter of the first version
//      (profiles/r1a_*: l1tex data-pipe wavefronts 92 % of peak).  A lane walks the samples i = phase, phase + PH, ...
//      sequentially in registers: plane-induced warp (fp32), 4-tap bilinear gather, dI/dt (1x3) and dI/dtheta (1x3,
//      right perturbation of the pose rotation) chained to the control knots through the per-sample spline blocks of
//      pose_kernel.  The PH partial sums of a pixel are combined by an xor-butterfly of warp shuffles;
//   C  residual, Huber, row = sqrt(w) [r, J]: the 32 rows of the chunk are staged in shared memory and the packed
//      upper triangle of rows^T rows is accumulated with one element set per lane (fp32 over 32 rows, fp64 across).
// Epilogue: deterministic block reduction -> per-block partials -> last block sums all partials in block order.
//
// Keyframe texels.  With LevelDev::ref_pair / ref_quad (built by pack_kernel below when the gradient image is exactly
// representable in fp16) a Hessian-pass sample needs 2 x 128-bit loads and a cost-only sample 1 x 32-bit load instead
// of 4 + 4 / 4 scattered loads; the values are bit-identical to ref_I / ref_dIxy.  Without them the same kernels
// gather ref_I / ref_dIxy directly (PACKED = false).
//
// Precision: per-sample arithmetic fp32 (the reference's bilinear taps/weights are fp32 too, compute_pixel_intensity.h:43-68),
// patch centre fp64 (its truncation picks the live pixel, …cost.cu:69-70), every sum across pixels fp64.
#include "mbavo_device.h"

#include <cuda_fp16.h>

#ifndef MBAVO_MINB_H
#define MBAVO_MINB_H 2
#endif

namespace mbavo
{
    namespace
    {
        __device__ __forceinline__ float u8_to_float(unsigned int b)
        {
            // exact for 0..255: place the byte in the mantissa of 2^23 and subtract 2^23 (full-rate LOP3 + FADD
            // instead of a quarter-rate I2F)
            return __uint_as_float(0x4B000000u | b) - 8388608.0f;
        }
        // byte k of a packed word as float: one PRMT builds 0x4B0000bb
        template <int BYTE>
        __device__ __forceinline__ float byte_to_float(unsigned int v)
        {
            return __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540 | BYTE)) - 8388608.0f;
        }
        __device__ __forceinline__ float2 halves(unsigned int v)
        {
            return __half22float2(*reinterpret_cast<const __half2 *>(&v));
        }
        __device__ __forceinline__ float rcp_approx(float x)
        {
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
            return r;
        }

        // Pixel record written by phase A (8 words, two 128-bit shared loads in phase B)
        struct PixelRec
        {
            float rx, ry;     // ray of the live pixel, (X - cx)/fx, (Y - cy)/fy (z = 1; the result is scale-invariant)
            float D;          // plane depth of the point
            float iD;         // 1 / (D + 1e-8): projection with P_z == D        compute_pixel_intensity.h:137
            float icur;       // live-image intensity at the pixel
            int X, Y;         // live pixel (also the origin of the reference coordinate, see sample_step)
            int state;        // bit 0: pixel inside the live image and point slot in range; bit 1: point flagged as outlier
        };
        static_assert(sizeof(PixelRec) == 32, "PixelRec must be 8 words");

        // Per-item setup: patch centre (fp64), integer live pixel, ray.  compute_local_patches_xy.cu:26-49,
        // compute_hessian_gradients_cost.cu:63-78.
        __device__ __forceinline__ PixelRec setup_pixel(const LevelDev &lv, const double *__restrict__ mid, int f, int p,
                                                        int j, const int2 *__restrict__ pattern_s)
        {
            PixelRec ps;
            ps.state = 0;
            ps.rx = ps.ry = 0.f;
            ps.D = 1.f;
            ps.iD = 1.f;
            ps.X = ps.Y = 0;
            ps.icur = 0.f;
            if (p >= lv.P)
                return ps;
            if (lv.flags[p] == 1)
                ps.state = 2;
            const double *xy = reinterpret_cast<const double *>(lv.xy + (size_t)p * lv.xy_stride + lv.xy_offset);
            const double x = xy[0], y = xy[1], z = lv.z[p];
            const double Prx = z * (x - lv.cx) / lv.fx, Pry = z * (y - lv.cy) / lv.fy, Prz = z;
            const double Pcx = mid[0] * Prx + mid[1] * Pry + mid[2] * Prz + mid[9];
            const double Pcy = mid[3] * Prx + mid[4] * Pry + mid[5] * Prz + mid[10];
            const double Pcz = mid[6] * Prx + mid[7] * Pry + mid[8] * Prz + mid[11];
            const double ccx = Pcx / Pcz * lv.fx + lv.cx;
            const double ccy = Pcy / Pcz * lv.fy + lv.cy;
            const int2 d = pattern_s[j];
            const double px = ccx + d.x, py = ccy + d.y;
            // (int) of a double: truncation toward zero; guard the conversion range first
            if (!(px > -1.0 && px < (double)lv.W && py > -1.0 && py < (double)lv.H))
                return ps;
            const int X = (int)px, Y = (int)py;
            if (X < 0 || X > lv.W - 1 || Y < 0 || Y > lv.H - 1)
                return ps;
            ps.state |= 1;
            ps.X = X, ps.Y = Y;
            ps.rx = (float)(((double)X - lv.cx) / lv.fx);
            ps.ry = (float)(((double)Y - lv.cy) / lv.fy);
            ps.D = (float)z;
            ps.iD = (float)(1.0 / (z + 1e-8));
            ps.icur = u8_to_float(__ldg(lv.cur_I[f] + (size_t)Y * lv.W + X));
            return ps;
        }

        // What a lane keeps in registers while it walks the samples of its pixel
        struct PixelRegs
        {
            float rx, ry, D, iD;
            float lox, hix, loy, hiy; // the reference coordinate (X + du, Y + dv) is inside [0, W-1] x [0, H-1] iff
                                      // lox <= du <= hix and loy <= dv <= hiy (all four are exact small integers)
            float fxiD, fyiD;         // fx / (D + 1e-8), fy / (D + 1e-8)
            int X, Y;
        };

        // One exposure sample of one pixel.  K knots per segment; OFF = segment offset inside the knot window.
        // Accumulates sumI and, if WITH_J, the 1 x 6NK row (Jt: translation block, Jw: rotation block).
        //
        // Reference coordinate.  compute_pixel_intensity.h:108-144 evaluates u = fx P_x / (D + 1e-8) + c_x in fp64 and
        // rounds only the fractional part to fp32.  A plain fp32 evaluation of u loses ~W * 2^-24 px (4e-5 px at VGA),
        // which is what limits the parity of H, g and the step.  Here the coordinate is evaluated RELATIVE TO THE LIVE
        // PIXEL X:  with A = (R - I) ray (small), m = ray + A, tau = t_z / (D + 1e-8),
        //     u - X = fx [ (A_x - r_x A_z - tau m_x) / m_z + t_x / (D + 1e-8) ]
        // in which every term is of the size of the blur, so fp32 keeps ~1e-6 px; the integer tap is X + floor(u - X).
        template <int K, int NK, bool WITH_J, bool PACKED, int OFF>
        __device__ __forceinline__ void sample_step(const float *__restrict__ rec, const PixelRegs &ps, const LevelDev &lv,
                                                    float fxf, float fyf, float &sumI,
                                                    float (&Jt)[WITH_J ? NK : 1][3], float (&Jw)[WITH_J ? NK : 1][3])
        {
            const float4 a0 = *reinterpret_cast<const float4 *>(rec);      // Rm0 Rm1 Rm2 Rm3   (Rm = R - I)
            const float4 a1 = *reinterpret_cast<const float4 *>(rec + 4);  // Rm4 Rm5 Rm6 Rm7
            const float4 a2 = *reinterpret_cast<const float4 *>(rec + 8);  // Rm8 tx ty tz
            const float A0 = fmaf(a0.x, ps.rx, fmaf(a0.y, ps.ry, a0.z));
            const float A1 = fmaf(a0.w, ps.rx, fmaf(a1.x, ps.ry, a1.y));
            const float A2 = fmaf(a1.z, ps.rx, fmaf(a1.w, ps.ry, a2.x));
            const float m0 = ps.rx + A0, m1 = ps.ry + A1, m2 = 1.0f + A2;
            const float il = rcp_approx(m2);                    // 1 / lambda (1 ulp: the terms it scales are blur-sized)
            const float tau = a2.w * ps.iD;
            const float du = fxf * fmaf(fmaf(-tau, m0, fmaf(-ps.rx, A2, A0)), il, a2.y * ps.iD);
            const float dv = fyf * fmaf(fmaf(-tau, m1, fmaf(-ps.ry, A2, A1)), il, a2.z * ps.iD);
            // inside [0, W-1] x [0, H-1] (compute_pixel_intensity.h:35); an invalid sample contributes nothing while the
            // divisor stays N (…cost.cu:107-110).  NaN / inf coordinates fail the comparisons.
            if (!(du >= ps.lox && du <= ps.hix && dv >= ps.loy && dv <= ps.hiy))
                return;
            const float xf = floorf(du), yf = floorf(dv);
            const float dx = du - xf, dy = dv - yf; // exact
            const int xi = ps.X + (int)xf, yi = ps.Y + (int)yf;

            // bilinear_interpolation, compute_pixel_intensity.h:40-68
            const float dxdy = dx * dy;
            const float w00 = 1.0f - dx - dy + dxdy, w01 = dx - dxdy, w10 = dy - dxdy, w11 = dxdy;
            // the +1 taps carry weight 0 on the last column / row; they are clamped instead of read out of bounds
            const int idx = yi * lv.W + xi;
            const int rowoff = yi < lv.H - 1 ? lv.W : 0;

            float gx, gy;
            if constexpr (PACKED && WITH_J)
            {
                const uint4 ta = __ldg(lv.ref_pair + idx);
                const uint4 tb = __ldg(lv.ref_pair + idx + rowoff);
                const float2 p00 = halves(ta.x), q00 = halves(ta.y), r00 = halves(ta.z); // I00 gx00 | gy00 I01 | gx01 gy01
                const float2 p10 = halves(tb.x), q10 = halves(tb.y), r10 = halves(tb.z); // I10 gx10 | gy10 I11 | gx11 gy11
                sumI += w11 * q10.y + w10 * p10.x + w01 * q00.y + w00 * p00.x;
                gx = w11 * r10.x + w10 * p10.y + w01 * r00.x + w00 * p00.y;
                gy = w11 * r10.y + w10 * q10.x + w01 * r00.y + w00 * q00.x;
            }
            else if constexpr (PACKED)
            {
                const unsigned int t = __ldg(lv.ref_quad + idx);
                sumI += w11 * byte_to_float<3>(t) + w10 * byte_to_float<2>(t) + w01 * byte_to_float<1>(t) +
                        w00 * byte_to_float<0>(t);
            }
            else
            {
                const int coloff = xi < lv.W - 1 ? 1 : 0;
                const int i00 = idx, i01 = idx + coloff, i10 = idx + rowoff, i11 = i10 + coloff;
                const float I00 = u8_to_float(__ldg(lv.ref_I + i00)), I01 = u8_to_float(__ldg(lv.ref_I + i01));
                const float I10 = u8_to_float(__ldg(lv.ref_I + i10)), I11 = u8_to_float(__ldg(lv.ref_I + i11));
                sumI += w11 * I11 + w10 * I10 + w01 * I01 + w00 * I00;
                if constexpr (WITH_J)
                {
                    const float2 g00 = __ldg(lv.ref_dIxy + i00), g01 = __ldg(lv.ref_dIxy + i01);
                    const float2 g10 = __ldg(lv.ref_dIxy + i10), g11 = __ldg(lv.ref_dIxy + i11);
                    gx = w11 * g11.x + w10 * g10.x + w01 * g01.x + w00 * g00.x;
                    gy = w11 * g11.y + w10 * g10.y + w01 * g01.y + w00 * g00.y;
                }
            }

            if constexpr (WITH_J)
            {
                const float s = (ps.D - a2.w) * il;                 // (D - t_z) / lambda       compute_pixel_intensity.h:128
                // dI/dt = dI/dP (I - m e_z^T / lambda)                                   compute_pixel_intensity.h:196-202
                const float gtx = gx * ps.fxiD, gty = gy * ps.fyiD;
                const float gtz = -(gtx * m0 + gty * m1) * il;
                // dI/dtheta = s (r x R^T dI/dt): right perturbation R <- R Exp(theta) of the pose rotation; equals
                // dI/dq (compute_pixel_intensity.h:179-206) contracted with dq/dtheta = L(q)[I/2;0]
                const float b0 = gtx + fmaf(a0.x, gtx, fmaf(a0.w, gty, a1.z * gtz));
                const float b1 = gty + fmaf(a0.y, gtx, fmaf(a1.x, gty, a1.w * gtz));
                const float b2 = gtz + fmaf(a0.z, gtx, fmaf(a1.y, gty, a2.x * gtz));
                const float v0 = s * (ps.ry * b2 - b1);
                const float v1 = s * (b0 - ps.rx * b2);
                const float v2 = s * (ps.rx * b1 - ps.ry * b0);
#pragma unroll
                for (int j = 0; j < K; ++j)
                {
                    const float wt = rec[12 + j];
                    const float *Th = rec + 12 + K + 9 * j;
                    Jt[OFF + j][0] = fmaf(wt, gtx, Jt[OFF + j][0]);
                    Jt[OFF + j][1] = fmaf(wt, gty, Jt[OFF + j][1]);
                    Jt[OFF + j][2] = fmaf(wt, gtz, Jt[OFF + j][2]);
                    Jw[OFF + j][0] = fmaf(v0, Th[0], fmaf(v1, Th[3], fmaf(v2, Th[6], Jw[OFF + j][0])));
                    Jw[OFF + j][1] = fmaf(v0, Th[1], fmaf(v1, Th[4], fmaf(v2, Th[7], Jw[OFF + j][1])));
                    Jw[OFF + j][2] = fmaf(v0, Th[2], fmaf(v1, Th[5], fmaf(v2, Th[8], Jw[OFF + j][2])));
                }
            }
        }

        template <int K, int NK, bool PACKED, int OFF>
        struct SegmentLoop
        {
            __device__ __forceinline__ static void run(const float *__restrict__ samples_s, const int *__restrict__ seg_end_s,
                                                       int &i, int PH, const PixelRegs &ps, const LevelDev &lv, float fxf,
                                                       float fyf, float &sumI, float (&Jt)[NK][3], float (&Jw)[NK][3])
            {
                constexpr int REC = sample_rec_floats(K);
                const int end = seg_end_s[OFF];
                for (; i < end; i += PH)
                    sample_step<K, NK, true, PACKED, OFF>(samples_s + i * REC, ps, lv, fxf, fyf, sumI, Jt, Jw);
                if constexpr (OFF + 1 <= NK - K)
                    SegmentLoop<K, NK, PACKED, (OFF + 1 <= NK - K ? OFF + 1 : OFF)>::run(samples_s, seg_end_s, i, PH, ps, lv, fxf,
                                                                                      fyf, sumI, Jt, Jw);
            }
        };

        // Huber on x = r^2/2 with threshold a^2 (compute_hessian_gradients_cost.cu:188-199)
        __device__ __forceinline__ float huber(float r, float a, float &sqrt_w)
        {
            const float aa = a * a, x = 0.5f * r * r;
            sqrt_w = 1.f;
            if (x > aa)
            {
                const float sx = sqrtf(x);
                sqrt_w = sqrtf(a / (sx + 1e-8f));
                return 2.f * a * sx - aa;
            }
            return x;
        }

        // K: knots per segment, NK: knots in the window (NK - K + 1 segments touched), WITH_J: Hessian pass or cost only,
        // PACKED: keyframe texels available.
        template <int K, int NK, bool WITH_J, bool PACKED>
        __global__ void __launch_bounds__(kThreads, WITH_J ? (NK <= 3 ? MBAVO_MINB_H : 1) : 4) track_kernel(const TrackParams prm)
        {
            constexpr int REC = sample_rec_floats(K);
            constexpr int NJ = WITH_J ? NK : 1;
            constexpr int D1 = WITH_J ? 6 * NK + 1 : 1;     // row length: [r | J]
            constexpr int D1P = D1 | 1;                      // odd row pitch: conflict-free row writes
            constexpr int E = WITH_J ? packed_len(NK) : 1;   // packed upper triangle
            constexpr int ME = (E + 31) / 32;                // elements owned by one lane

            const LevelDev &lv = prm.lv;
            const int f = blockIdx.y;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int S = lv.S, N = lv.N, TP = prm.TP, PH = prm.PH;
            const int Q = 32 / PH;                           // pixel slots per pass
            // lane -> (slot, phase): slot-fastest (a quarter-warp = 8 pixels of one phase) or phase-fastest (a quarter-warp
            // = 8 consecutive exposure samples of one pixel, which mostly share a 128-byte line)
            const bool pf = prm.phase_fast != 0;
            const int slot = pf ? lane / PH : lane & (Q - 1), phase = pf ? lane & (PH - 1) : lane / Q;
            const int bfly_lo = pf ? 1 : Q, bfly_hi = pf ? PH : 32;

            extern __shared__ __align__(16) unsigned char smem_raw[];
            float *samples_s = reinterpret_cast<float *>(smem_raw);                  // N * REC
            int *seg_end_s = reinterpret_cast<int *>(samples_s + N * REC);          // kMaxSegments (+1 pad)
            PixelRec *pix_s = reinterpret_cast<PixelRec *>(seg_end_s + 8);          // warps * 32
            int2 *pattern_s = reinterpret_cast<int2 *>(pix_s + kWarpsPerBlock * 32); // S
            float *rho_s = reinterpret_cast<float *>(pattern_s + S);                // warps * max(32, TP*S)
            const int rho_per_warp = max(32, TP * S);
            unsigned short *pair_s = reinterpret_cast<unsigned short *>(rho_s + kWarpsPerBlock * rho_per_warp); // E (padded)
            float *rows_s = reinterpret_cast<float *>(pair_s + ((E + 7) & ~7));    // warps * 32 * D1P
            double *red_s = reinterpret_cast<double *>(smem_raw);                   // epilogue: warps * E doubles (aliases all)

            for (int e = threadIdx.x; e < N * REC; e += blockDim.x)
                samples_s[e] = prm.samples[(size_t)f * N * REC + e];
            if (threadIdx.x < kMaxSegments)
                seg_end_s[threadIdx.x] = prm.seg_end[f * kMaxSegments + threadIdx.x];
            for (int e = threadIdx.x; e < S; e += blockDim.x)
                pattern_s[e] = lv.pattern[e];
            if (WITH_J)
            {
                // pair table: packed index e -> (a, b), a <= b, row-major upper triangle of the D1 x D1 matrix
                for (int e = threadIdx.x; e < E; e += blockDim.x)
                {
                    int a = 0, rem = e;
                    while (rem >= D1 - a)
                    {
                        rem -= D1 - a;
                        ++a;
                    }
                    pair_s[e] = (unsigned short)((a << 8) | (a + rem));
                }
            }
            __syncthreads();

            const double *mid = prm.mid + f * kMidDoubles;
            const float fxf = (float)lv.fx, fyf = (float)lv.fy;
            const float inv_N = 1.0f / (float)N;
            const float huber_a = prm.huber_a;
            const double inv_num_residuals = prm.inv_num_residuals;
            float *my_rows = rows_s + warp * 32 * D1P;
            float *my_rho = rho_s + warp * rho_per_warp;
            PixelRec *my_pix = pix_s + warp * 32;

            double acc[ME];
#pragma unroll
            for (int m = 0; m < ME; ++m)
                acc[m] = 0.0;
            double cost_acc = 0.0;

            const int items = TP * S;
            for (int wb = blockIdx.x * kWarpsPerBlock + warp; wb < prm.batches_per_frame; wb += gridDim.x * kWarpsPerBlock)
            {
                const int p0 = wb * TP;
                for (int base = 0; base < items; base += 32)
                {
                    // ---- A: pixel records -----------------------------------------------------------------------------
                    {
                        const int it = base + lane;
                        const bool in_batch = it < items;
                        my_pix[lane] = setup_pixel(lv, mid, f, in_batch ? p0 + it / S : lv.P, in_batch ? it % S : 0, pattern_s);
                    }
                    __syncwarp();
                    // ---- B: exposure samples, PH phases per pixel ---------------------------------------------------
                    const int chunk = min(32, items - base);
                    for (int sub = 0; sub < chunk; sub += Q)
                    {
                        const int q = sub + slot;
                        const float4 r0 = *reinterpret_cast<const float4 *>(my_pix + q);
                        const int4 r1 = *(reinterpret_cast<const int4 *>(my_pix + q) + 1);
                        const bool valid = r1.w & 1;
                        PixelRegs ps;
                        ps.rx = r0.x, ps.ry = r0.y, ps.D = r0.z, ps.iD = r0.w;
                        ps.X = r1.y, ps.Y = r1.z;
                        ps.lox = -(float)ps.X, ps.hix = (float)(lv.W - 1 - ps.X);
                        ps.loy = -(float)ps.Y, ps.hiy = (float)(lv.H - 1 - ps.Y);
                        ps.fxiD = fxf * ps.iD, ps.fyiD = fyf * ps.iD;

                        float sumI = 0.f;
                        float Jt[NJ][3], Jw[NJ][3];
#pragma unroll
                        for (int a = 0; a < NJ; ++a)
                            Jt[a][0] = Jt[a][1] = Jt[a][2] = Jw[a][0] = Jw[a][1] = Jw[a][2] = 0.f;

                        if (valid)
                        {
                            if constexpr (WITH_J)
                            {
                                int i = phase;
                                SegmentLoop<K, NK, PACKED, 0>::run(samples_s, seg_end_s, i, PH, ps, lv, fxf, fyf, sumI, Jt, Jw);
                            }
                            else
                            {
                                for (int i = phase; i < N; i += PH) // cost only: the segment of a sample is irrelevant
                                    sample_step<K, NK, false, PACKED, 0>(samples_s + i * REC, ps, lv, fxf, fyf, sumI, Jt, Jw);
                            }
                        }
                        // combine the PH phases of every pixel (fixed butterfly order: deterministic)
                        for (int o = bfly_lo; o < bfly_hi; o <<= 1)
                        {
                            sumI += __shfl_xor_sync(0xffffffffu, sumI, o);
                            if constexpr (WITH_J)
                            {
#pragma unroll
                                for (int a = 0; a < NK; ++a)
                                {
#pragma unroll
                                    for (int c = 0; c < 3; ++c)
                                    {
                                        Jt[a][c] += __shfl_xor_sync(0xffffffffu, Jt[a][c], o);
                                        Jw[a][c] += __shfl_xor_sync(0xffffffffu, Jw[a][c], o);
                                    }
                                }
                            }
                        }
                        // residual (…cost.cu:115-121), Huber, weighted row (…cost.cu:185-206)
                        const float icur = __int_as_float(r1.x);
                        const float r = valid ? sumI * inv_N - icur : 0.f;
                        float sw;
                        const float rho = huber(r, huber_a, sw);
                        if (phase == 0 && q < chunk)
                        {
                            my_rho[base + q] = rho;
                            if constexpr (WITH_J)
                            {
                                const float scale = (r1.w == 1) ? sw : 0.f; // invalid pixels and outliers add nothing (…cost.cu:267)
                                float *row = my_rows + q * D1P;
                                row[0] = scale * r;
                                const float sj = scale * inv_N;
#pragma unroll
                                for (int a = 0; a < NK; ++a)
                                {
                                    row[1 + 3 * a + 0] = sj * Jt[a][0];
                                    row[1 + 3 * a + 1] = sj * Jt[a][1];
                                    row[1 + 3 * a + 2] = sj * Jt[a][2];
                                    row[1 + 3 * NK + 3 * a + 0] = sj * Jw[a][0];
                                    row[1 + 3 * NK + 3 * a + 1] = sj * Jw[a][1];
                                    row[1 + 3 * NK + 3 * a + 2] = sj * Jw[a][2];
                                }
                            }
                        }
                    }
                    if constexpr (WITH_J)
                    {
                        // rows of a short last chunk: zero
                        if (lane >= chunk)
                        {
                            float *row = my_rows + lane * D1P;
#pragma unroll
                            for (int c = 0; c < D1; ++c)
                                row[c] = 0.f;
                        }
                        __syncwarp();
                        // ---- C: packed upper triangle of rows^T rows (…cost.cu:214-230): lane owns elements lane, lane+32, …
#pragma unroll
                        for (int m = 0; m < ME; ++m)
                        {
                            const int e = lane + 32 * m;
                            if (e < E)
                            {
                                const int a = pair_s[e] >> 8, b = pair_s[e] & 0xff;
                                float sacc = 0.f;
#pragma unroll 8
                                for (int qq = 0; qq < 32; ++qq)
                                    sacc = fmaf(my_rows[qq * D1P + a], my_rows[qq * D1P + b], sacc);
                                acc[m] += (double)sacc;
                            }
                        }
                    }
                    __syncwarp();
                }
                // per-patch cost (…cost.cu:232-238) in fixed pixel order; element 0 of the patch vector
                if (lane < TP && p0 + lane < lv.P)
                {
                    double sp = 0.0;
                    for (int jj = 0; jj < S; ++jj)
                        sp += (double)my_rho[lane * S + jj];
                    lv.patch_cost[((size_t)f * lv.P + p0 + lane) * lv.patch_cost_stride] = sp * inv_num_residuals;
                    if (lv.flags[p0 + lane] != 1)
                        cost_acc += sp;
                }
                __syncwarp();
            }

            // ---- epilogue: warp -> block -> grid, all in fixed order (deterministic) --------------------------------
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);
            if (lane == 0)
                acc[0] = cost_acc; // packed[0] is the cost, not r^2 (…cost.cu:232-238)

            __syncthreads(); // every warp is done with samples_s / rows_s before red_s (aliasing them) is written
#pragma unroll
            for (int m = 0; m < ME; ++m)
            {
                const int e = lane + 32 * m;
                if (e < E)
                    red_s[warp * E + e] = acc[m];
            }
            __syncthreads();
            const int block_linear = blockIdx.y * gridDim.x + blockIdx.x;
            const int num_blocks = gridDim.x * gridDim.y;
            for (int e = threadIdx.x; e < E; e += blockDim.x)
            {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kWarpsPerBlock; ++w)
                    s += red_s[w * E + e];
                prm.block_partials[(size_t)block_linear * E + e] = s;
            }
            __threadfence();
            __shared__ unsigned int ticket_s;
            __syncthreads();
            if (threadIdx.x == 0)
                ticket_s = atomicAdd(prm.counter, 1u);
            __syncthreads();
            if (ticket_s != (unsigned int)(num_blocks - 1))
                return;
            // last block: sum the partials of all blocks in block order
            __threadfence();
            for (int e = threadIdx.x; e < E; e += blockDim.x)
            {
                double s = 0.0;
                for (int b = 0; b < num_blocks; ++b)
                    s += __ldcg(prm.block_partials + (size_t)b * E + e);
                s *= inv_num_residuals;
                prm.packed_out[e] = s;
                if (prm.host_out)
                    prm.host_out[e] = s;
            }
            if (prm.host_out)
            {
                __threadfence_system(); // the vector is visible to the host before the sequence number is
                __syncthreads();
            }
            if (threadIdx.x == 0)
            {
                *prm.counter = 0u; // re-arm for the next launch
                if (prm.host_out)
                    *prm.host_seq = prm.seq;
            }
        }

        // Keyframe texels (see LevelDev).  One thread per pixel; *inexact counts gradient values that fp16 cannot hold.
        __global__ void pack_kernel(const unsigned char *__restrict__ I, const float2 *__restrict__ g, int H, int W,
                                    uint4 *__restrict__ pair, unsigned int *__restrict__ quad, int *__restrict__ inexact)
        {
            const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
            if (x >= W)
                return;
            const int x1 = min(x + 1, W - 1), y1 = min(y + 1, H - 1);
            const int i00 = y * W + x, i01 = y * W + x1, i10 = y1 * W + x, i11 = y1 * W + x1;
            const unsigned int b00 = I[i00], b01 = I[i01], b10 = I[i10], b11 = I[i11];
            quad[i00] = b00 | (b01 << 8) | (b10 << 16) | (b11 << 24);
            const float2 g0 = g[i00], g1 = g[i01];
            const __half hx0 = __float2half_rn(g0.x), hy0 = __float2half_rn(g0.y);
            const __half hx1 = __float2half_rn(g1.x), hy1 = __float2half_rn(g1.y);
            // bitwise round trip (also rejects NaN and values that overflow to inf)
            if (__float_as_uint(__half2float(hx0)) != __float_as_uint(g0.x) ||
                __float_as_uint(__half2float(hy0)) != __float_as_uint(g0.y))
                atomicAdd(inexact, 1);
            const unsigned int hI0 = __half_as_ushort(__float2half_rn((float)b00)), hI1 = __half_as_ushort(__float2half_rn((float)b01));
            uint4 t;
            t.x = hI0 | ((unsigned int)__half_as_ushort(hx0) << 16);
            t.y = (unsigned int)__half_as_ushort(hy0) | (hI1 << 16);
            t.z = (unsigned int)__half_as_ushort(hx1) | ((unsigned int)__half_as_ushort(hy1) << 16);
            t.w = 0u;
            pair[i00] = t;
        }

        template <int K, int NK, bool WITH_J, bool PACKED>
        cudaError_t launch_one(const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream)
        {
            static unsigned long long configured = 0; // per instantiation and per device (attribute of the device function)
            int dev = 0;
            cudaGetDevice(&dev);
            if (!(configured >> dev & 1ull))
            {
                cudaError_t e = cudaFuncSetAttribute(track_kernel<K, NK, WITH_J, PACKED>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e != cudaSuccess)
                    return e;
                configured |= 1ull << dev;
            }
            track_kernel<K, NK, WITH_J, PACKED><<<grid, kThreads, smem, stream>>>(prm);
            return cudaGetLastError();
        }

        // Occupancy-derived grid width for one instantiation
        template <int K, int NK, bool WITH_J, bool PACKED>
        int blocks_per_sm(size_t smem)
        {
            int n = 0;
            cudaFuncSetAttribute(track_kernel<K, NK, WITH_J, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, track_kernel<K, NK, WITH_J, PACKED>, kThreads, smem);
            return n > 0 ? n : 1;
        }

        template <int K, int NK, bool WITH_J>
        cudaError_t dispatch_packed(bool packed, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                    int *query_occupancy)
        {
            if (query_occupancy)
            {
                *query_occupancy = packed ? blocks_per_sm<K, NK, WITH_J, true>(smem) : blocks_per_sm<K, NK, WITH_J, false>(smem);
                return cudaSuccess;
            }
            return packed ? launch_one<K, NK, WITH_J, true>(prm, grid, smem, stream)
                          : launch_one<K, NK, WITH_J, false>(prm, grid, smem, stream);
        }
    } // namespace

    size_t track_kernel_smem_bytes(int K, int NK, bool with_j, int N, int S, int TP)
    {
        const int REC = sample_rec_floats(K);
        const int D1 = with_j ? 6 * NK + 1 : 1, D1P = D1 | 1;
        const int E = with_j ? packed_len(NK) : 1;
        const int rho_per_warp = max(32, TP * S);
        size_t main_bytes = (size_t)N * REC * 4 + 8 * 4 + (size_t)kWarpsPerBlock * 32 * sizeof(PixelRec) + (size_t)S * 8 +
                            (size_t)kWarpsPerBlock * rho_per_warp * 4 + (size_t)((E + 7) & ~7) * 2 +
                            (size_t)kWarpsPerBlock * 32 * D1P * 4;
        size_t red_bytes = (size_t)kWarpsPerBlock * E * 8;
        return (main_bytes > red_bytes ? main_bytes : red_bytes) + 16;
    }

    cudaError_t launch_pack_kernel(const unsigned char *I, const float *dIxy, int H, int W, uint4 *pair, unsigned int *quad,
                                   int *inexact, cudaStream_t stream)
    {
        const dim3 block(128, 1, 1), grid((W + 127) / 128, H, 1);
        pack_kernel<<<grid, block, 0, stream>>>(I, reinterpret_cast<const float2 *>(dIxy), H, W, pair, quad, inexact);
        return cudaGetLastError();
    }

#define MBAVO_DISPATCH(K_, NK_)   \
    if (K == K_ && NK == NK_)     \
        return dispatch_packed<K_, NK_, true>(packed, prm, grid, smem, stream, query_occupancy);

    // with_j: Hessian pass (templated on the window) or cost-only pass (one instantiation per K)
    cudaError_t launch_track_kernel(int K, int NK, bool with_j, const TrackParams &prm, dim3 grid, size_t smem,
                                    cudaStream_t stream, int *query_occupancy)
    {
        const bool packed = prm.lv.ref_pair != nullptr || (query_occupancy && prm.PH < 0);
        if (!with_j)
        {
            if (K == 2)
                return dispatch_packed<2, 2, false>(packed, prm, grid, smem, stream, query_occupancy);
            if (K == 4)
                return dispatch_packed<4, 4, false>(packed, prm, grid, smem, stream, query_occupancy);
            return cudaErrorInvalidValue;
        }
        MBAVO_DISPATCH(2, 2)
        MBAVO_DISPATCH(2, 3)
        MBAVO_DISPATCH(2, 4)
        MBAVO_DISPATCH(2, 5)
        MBAVO_DISPATCH(2, 6)
        MBAVO_DISPATCH(4, 4)
        MBAVO_DISPATCH(4, 5)
        MBAVO_DISPATCH(4, 6)
        MBAVO_DISPATCH(4, 7)
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
