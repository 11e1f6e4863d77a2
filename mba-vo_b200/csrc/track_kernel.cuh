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
//   B  lane = pixel (optionally lane = (pixel slot, exposure phase), MBAVO_PHASES): the lane walks the exposure samples
//      sequentially in registers: plane-induced warp (fp32), 4-tap bilinear gather, dI/dt (1x3) and dI/dtheta (1x3,
//      right perturbation of the pose rotation) chained to the control knots through the per-sample spline blocks of
//      pose_kernel.  The arithmetic is written in 2-vectors — (x, y) image components, matrix column / row pairs, pairs
//      of Jacobian columns — and issued as packed FFMA2 / FMUL2 / FADD2 (sm_100 packed fp32, one operand may be a
//      broadcast scalar): the kernel is issue-bound, and packing removes ~30 % of its instructions;
//   C  residual, Huber, row = sqrt(w) [r, J]: the 32 rows of the chunk are staged in shared memory and rows^T rows is
//      accumulated in 2x2 register tiles, one or more tiles per lane (fp32 over the 32 rows, fp64 across chunks).
// Epilogue: deterministic block reduction -> per-block partials -> last block sums all partials in block order.
//
// Keyframe texels.  With LevelDev::ref_pair / ref_quad (built by pack_kernel, track_common.cu, when the gradient image is
// reproducible from byte differences, LevelDev) a Hessian-pass sample needs 1 x 128-bit load and a cost-only sample 1 x 32-bit
// load instead of 4 + 4 / 4 scattered loads; the values are bit-identical to ref_I / ref_dIxy.  Without them the same kernels
// gather ref_I / ref_dIxy directly (PACKED = false).
//
// Precision: per-sample arithmetic fp32 (the reference's bilinear taps/weights are fp32 too, compute_pixel_intensity.h:43-68),
// patch centre fp64 (its truncation picks the live pixel, …cost.cu:69-70), every sum across pixels fp64.
#ifndef MBAVO_TRACK_KERNEL_CUH_
#define MBAVO_TRACK_KERNEL_CUH_
#include "mbavo_device.h"
#include "pose_device.cuh"

#include <cuda_fp16.h>

// Hessian pass: branch-free sample step, sample loop unrolled by 2 (measured best: profiles/r1_history.md); the cost-only
// pass keeps the early exit of an invalid sample.
#ifndef MBAVO_MINB_C
#define MBAVO_MINB_C 4 // resident blocks per SM of the cost-only pass (8 warps each)
#endif
#ifndef MBAVO_BRANCHLESS
#define MBAVO_BRANCHLESS 1
#endif
#ifndef MBAVO_UNROLL
#define MBAVO_UNROLL 2
#endif
// cost-only pass: branch-free sample step (an invalid sample runs with zero weights and a safe tap) so that the unrolled
// loop keeps several texel loads in flight per lane; MBAVO_COST_BRANCHLESS=0 restores the early exit
#ifndef MBAVO_COST_BRANCHLESS
#define MBAVO_COST_BRANCHLESS 1
#endif
#ifndef MBAVO_UNROLL_C
#define MBAVO_UNROLL_C 4
#endif
// floor + float->int of the blur-sized tap offset by the 1.5 * 2^23 trick (full-rate FADD.RM / IADD) instead of
// FRND + F2I (quarter-rate XU pipe)
#ifndef MBAVO_MAGIC_FLOOR
#define MBAVO_MAGIC_FLOOR 1
#endif
// Hessian pass with patch texels: sample loop software-pipelined by hand (sample_front / sample_back below)
#ifndef MBAVO_PIPELINE
#define MBAVO_PIPELINE 0
#endif
#ifndef MBAVO_PIPE_UNROLL
#define MBAVO_PIPE_UNROLL 2
#endif
// last block: sum of the per-block partials with 128-bit loads, all of a thread's rows in flight at once (pass_finish)
#ifndef MBAVO_WIDE_SUM
#define MBAVO_WIDE_SUM 1
#endif
// persistent sweep, serial section of a Hessian pass: the knots the step starts from are fetched into shared memory under the
// sum of the partials, the candidate reaches the pose records through shared memory (no L2 round trips), the sample positions
// u are computed once at kernel entry, the candidate's quaternions take the power series of so3_exp_only
#ifndef MBAVO_FAST_FINISH
#define MBAVO_FAST_FINISH 1
#endif
// serial sections, third cut: pivot reciprocals seeded by rcp.approx.f64 (one MUFU instead of a float division between two
// conversions), the record-buffer selector and the inputs of the cost pass's commit fetched under the sum of the partials
#ifndef MBAVO_FINISH3
#define MBAVO_FINISH3 1
#endif
#ifndef MBAVO_RCP64
#define MBAVO_RCP64 MBAVO_FINISH3
#endif
#ifndef MBAVO_PIPE_LEAN
#define MBAVO_PIPE_LEAN 1 // 8 registers cross the split instead of 10 (m01 is recomputed from the re-read rotation groups)
#endif

namespace mbavo
{
    namespace
    {
        constexpr int kSampleUnroll = MBAVO_UNROLL;
        constexpr int kCostUnroll = MBAVO_UNROLL_C;
        constexpr int kPipeUnroll = MBAVO_PIPE_UNROLL;
        constexpr int MBAVO_MAX_LEVELS_DEV = 8; // host_out layout of a sweep: 4 scalars per level, then the final knots

#ifdef MBAVO_PROFILE_PHASES
#define MBAVO_STAMP(slot)                                                                          \
    do                                                                                             \
    {                                                                                              \
        if (prm.phase_times && threadIdx.x == 0 && (blockIdx.x == 0 || (slot) >= 8))               \
            prm.phase_times[16 * (prm.trace_row & 63) + (slot)] = global_timer_ns();                                             \
    } while (0)
#else
#define MBAVO_STAMP(slot) \
    do                    \
    {                     \
    } while (0)
#endif
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
        // 2^23 + byte k of a packed word (exact): the bias cancels in a difference of two of them
        template <int BYTE>
        __device__ __forceinline__ float biased_byte(unsigned int v)
        {
            return __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540 | BYTE));
        }
        // packed fp32 pairs (FFMA2 / FMUL2 / FADD2); bc() broadcasts a scalar, which the instruction takes as a .F32 operand
        __device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
        __device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
        __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
        __device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
        __device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

        // the bilinear blend (compute_pixel_intensity.h:40-68) with the contraction spelled out, so that every texel format and
        // the direct gather round identically
        __device__ __forceinline__ float blend4(float w00, float w01, float w10, float w11, float i00, float i01, float i10, float i11)
        {
            return fmaf(w00, i00, fmaf(w01, i01, fmaf(w10, i10, w11 * i11)));
        }
        __device__ __forceinline__ float2 halves(unsigned int v)
        {
            return __half22float2(*reinterpret_cast<const __half2 *>(&v));
        }
        // floor(x) as float, and as int in `i`, for |x| < 2^22 (tap offsets are bounded by the image size): adding 1.5 * 2^23
        // with round-down leaves the integer part in the mantissa — two full-rate FADDs and one IADD where floorf + (int)
        // cost an FRND and an F2I on the quarter-rate XU pipe.  Out-of-range / NaN inputs give garbage that callers discard.
        __device__ __forceinline__ float floor_to_int(float x, int &i)
        {
#if MBAVO_MAGIC_FLOOR
            const float t = __fadd_rd(x, 12582912.0f);
            i = __float_as_int(t) - 0x4B400000;
            return t - 12582912.0f;
#else
            const float f = floorf(x);
            i = (int)f;
            return f;
#endif
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
                                                        int j, const int2 *__restrict__ pattern_s, double2 *centre_out = nullptr)
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
            if (centre_out)
                *centre_out = make_double2(ccx, ccy);
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
            // (no discrete decision depends on these: reciprocal multiplies instead of fp64 divisions)
            ps.rx = (float)(((double)X - lv.cx) * lv.inv_fx);
            ps.ry = (float)(((double)Y - lv.cy) * lv.inv_fy);
            ps.D = (float)z;
            ps.iD = __frcp_rn((float)(z + 1e-8));
            ps.icur = u8_to_float(__ldg(lv.cur_I[f] + (size_t)Y * lv.W + X));
            return ps;
        }

        // A sample record in shared memory; q(k) = k-th group of four floats
        struct SmemRec
        {
            const float *p;
            __device__ __forceinline__ float4 q(int k) const { return *reinterpret_cast<const float4 *>(p + 4 * k); }
        };

        // What a lane keeps in registers while it walks the samples of its pixel
        struct PixelRegs
        {
            float2 rxy;               // ray of the live pixel (z = 1)
            float D, iD;              // plane depth, 1 / (D + 1e-8)
            float lox, hix, loy, hiy; // the reference coordinate (X + du, Y + dv) is inside [0, W-1] x [0, H-1] iff
                                      // lox <= du <= hix and loy <= dv <= hiy (all four are exact small integers)
            float2 fxyiD;             // (fx, fy) / (D + 1e-8); halved when the texels deliver doubled gradients
            int X, Y;
        };

        // One exposure sample of one pixel.  K knots per segment; OFF = segment offset inside the knot window.
        // Accumulates sumI and, if WITH_J, the 1 x 6NK row as three pairs per knot a:
        //     J[a][0] = (dt_x, dt_y)   J[a][1] = (dt_z, dw_x)   J[a][2] = (dw_y, dw_z)
        //
        // Reference coordinate.  compute_pixel_intensity.h:108-144 evaluates u = fx P_x / (D + 1e-8) + c_x in fp64 and
        // rounds only the fractional part to fp32.  A plain fp32 evaluation of u loses ~W * 2^-24 px (4e-5 px at VGA),
        // which is what limits the parity of H, g and the step.  Here the coordinate is evaluated RELATIVE TO THE LIVE
        // PIXEL X:  with A = (R - I) ray (small), m = ray + A, tau = t_z / (D + 1e-8),
        //     u - X = fx [ (A_x - r_x A_z - tau m_x) / m_z + t_x / (D + 1e-8) ]
        // in which every term is of the size of the blur, so fp32 keeps ~1e-6 px; the integer tap is X + floor(u - X).
        template <int K, int NK, bool WITH_J, bool PACKED, int OFF, class Rec>
        __device__ __forceinline__ void sample_step(const Rec rec, const PixelRegs &ps, const LevelDev &lv,
                                                    float2 fxy, float &sumI, float2 (&J)[WITH_J ? NK : 1][3])
        {
            const float4 g0 = rec.q(0); // (Rm00 Rm01) (Rm10 Rm11)      Rm = R - I
            const float4 g1 = rec.q(1); // (Rm20 Rm21) Rm02 Rm12
            const float4 g2 = rec.q(2); // Rm22 tz (tx ty)
            const float2 A01 = f2(fmaf(g0.x, ps.rxy.x, fmaf(g0.y, ps.rxy.y, g1.z)), fmaf(g0.z, ps.rxy.x, fmaf(g0.w, ps.rxy.y, g1.w)));
            const float A2 = fmaf(g1.x, ps.rxy.x, fmaf(g1.y, ps.rxy.y, g2.x));
            const float2 m01 = add2(ps.rxy, A01);
            const float m2 = 1.0f + A2;
            const float il_raw = rcp_approx(m2);                // 1 / lambda (1 ulp: the terms it scales are blur-sized)
            const float tau = g2.y * ps.iD;
            const float2 num = fma2(bc(-tau), m01, fma2(bc(-A2), ps.rxy, A01));
            const float2 duv = mul2(fxy, fma2(num, bc(il_raw), mul2(f2(g2.z, g2.w), bc(ps.iD))));
            // inside [0, W-1] x [0, H-1] (compute_pixel_intensity.h:35); an invalid sample contributes nothing while the
            // divisor stays N (…cost.cu:107-110).  NaN / inf coordinates fail the comparisons.
            float il, w00, w01, w10, w11;
            int idx, rowoff, xi;
            if constexpr (WITH_J ? MBAVO_BRANCHLESS : MBAVO_COST_BRANCHLESS)
            {
            // Branch-free form: an invalid sample keeps going with zero weights, a safe tap address and il = 0 (so that
            // nothing non-finite reaches the sums).  Straight-line code lets the compiler overlap the loads of one sample
            // with the arithmetic of the previous one when the sample loop is unrolled.
                const bool ok = duv.x >= ps.lox && duv.x <= ps.hix && duv.y >= ps.loy && duv.y <= ps.hiy;
                il = ok ? il_raw : 0.f;
                int xo, yo;
                const float xf = floor_to_int(duv.x, xo), yf = floor_to_int(duv.y, yo);
                const float dx = duv.x - xf, dy = duv.y - yf; // exact
                const int yi = ps.Y + yo;
                xi = ok ? ps.X + xo : 0;
                const float dxdy = dx * dy;
                w00 = ok ? 1.0f - dx - dy + dxdy : 0.f, w01 = ok ? dx - dxdy : 0.f, w10 = ok ? dy - dxdy : 0.f, w11 = ok ? dxdy : 0.f;
                idx = ok ? yi * lv.W + xi : 0;
                rowoff = (ok && yi < lv.H - 1) ? lv.W : 0;
            }
            else
            {
                if (!(duv.x >= ps.lox && duv.x <= ps.hix && duv.y >= ps.loy && duv.y <= ps.hiy))
                    return;
                il = il_raw;
                int xo, yo;
                const float xf = floor_to_int(duv.x, xo), yf = floor_to_int(duv.y, yo);
                const float dx = duv.x - xf, dy = duv.y - yf; // exact
                const int yi = ps.Y + yo;
                xi = ps.X + xo;
                // bilinear_interpolation, compute_pixel_intensity.h:40-68
                const float dxdy = dx * dy;
                w00 = 1.0f - dx - dy + dxdy, w01 = dx - dxdy, w10 = dy - dxdy, w11 = dxdy;
                // the +1 taps carry weight 0 on the last column / row; they are clamped instead of read out of bounds
                idx = yi * lv.W + xi;
                rowoff = yi < lv.H - 1 ? lv.W : 0;
            }

            float2 gxy;
            if constexpr (PACKED && WITH_J)
            {
#if MBAVO_TEXEL == 3
                // patch texel: the 4 x 4 byte neighbourhood of the tap in one 128-bit load.  biased(b) = 2^23 + b exactly; the
                // difference of two biased bytes is the DOUBLED central difference, exactly (PixelRegs::fxyiD carries the 1/2)
                const uint4 t = __ldg(lv.ref_pair + idx);
                (void)rowoff;
                const float r0c1 = biased_byte<1>(t.x), r0c2 = biased_byte<2>(t.x);
                const float r1c0 = biased_byte<0>(t.y), r1c1 = biased_byte<1>(t.y), r1c2 = biased_byte<2>(t.y), r1c3 = biased_byte<3>(t.y);
                const float r2c0 = biased_byte<0>(t.z), r2c1 = biased_byte<1>(t.z), r2c2 = biased_byte<2>(t.z), r2c3 = biased_byte<3>(t.z);
                const float r3c1 = biased_byte<1>(t.w), r3c2 = biased_byte<2>(t.w);
                const float2 g00 = f2(r1c2 - r1c0, r2c1 - r0c1), g01 = f2(r1c3 - r1c1, r2c2 - r0c2);
                const float2 g10 = f2(r2c2 - r2c0, r3c1 - r1c1), g11 = f2(r2c3 - r2c1, r3c2 - r1c2);
                const float2 i0 = f2(r1c1 - 8388608.0f, r1c2 - 8388608.0f), i1 = f2(r2c1 - 8388608.0f, r2c2 - 8388608.0f);
#else
                const uint4 ta = __ldg(lv.ref_pair + idx);
                const uint4 tb = __ldg(lv.ref_pair + idx + rowoff);
                const float2 g00 = halves(ta.x), g01 = halves(ta.y), i0 = halves(ta.z); // (gx gy)(x,y) | (gx gy)(x+1,y) | I(x,y) I(x+1,y)
                const float2 g10 = halves(tb.x), g11 = halves(tb.y), i1 = halves(tb.z);
#endif
                sumI += blend4(w00, w01, w10, w11, i0.x, i0.y, i1.x, i1.y);
                gxy = fma2(bc(w00), g00, fma2(bc(w01), g01, fma2(bc(w10), g10, mul2(bc(w11), g11))));
            }
            else if constexpr (PACKED)
            {
                const unsigned int t = __ldg(lv.ref_quad + idx);
                sumI += blend4(w00, w01, w10, w11, byte_to_float<0>(t), byte_to_float<1>(t), byte_to_float<2>(t), byte_to_float<3>(t));
            }
            else
            {
                const int coloff = xi < lv.W - 1 ? 1 : 0;
                const int i00 = idx, i01 = idx + coloff, i10 = idx + rowoff, i11 = i10 + coloff;
                const float I00 = u8_to_float(__ldg(lv.ref_I + i00)), I01 = u8_to_float(__ldg(lv.ref_I + i01));
                const float I10 = u8_to_float(__ldg(lv.ref_I + i10)), I11 = u8_to_float(__ldg(lv.ref_I + i11));
                sumI += blend4(w00, w01, w10, w11, I00, I01, I10, I11);
                if constexpr (WITH_J)
                {
                    const float2 g00 = __ldg(lv.ref_dIxy + i00), g01 = __ldg(lv.ref_dIxy + i01);
                    const float2 g10 = __ldg(lv.ref_dIxy + i10), g11 = __ldg(lv.ref_dIxy + i11);
                    gxy = fma2(bc(w00), g00, fma2(bc(w01), g01, fma2(bc(w10), g10, mul2(bc(w11), g11))));
                }
            }

            if constexpr (WITH_J)
            {
                const float s = (ps.D - g2.y) * il;                 // (D - t_z) / lambda       compute_pixel_intensity.h:128
                // dI/dt = dI/dP (I - m e_z^T / lambda)                                   compute_pixel_intensity.h:196-202
                const float2 gt = mul2(gxy, ps.fxyiD);
                const float gtz = -(gt.x * m01.x + gt.y * m01.y) * il;
                // dI/dtheta = s (r x R^T dI/dt): right perturbation R <- R Exp(theta) of the pose rotation; equals
                // dI/dq (compute_pixel_intensity.h:179-206) contracted with dq/dtheta = L(q)[I/2;0].   b = s R^T dI/dt
                const float2 b01 = mul2(bc(s), fma2(bc(gt.x), f2(g0.x, g0.y),
                                                    fma2(bc(gt.y), f2(g0.z, g0.w), fma2(bc(gtz), f2(g1.x, g1.y), gt))));
                const float b2 = s * fmaf(g1.z, gt.x, fmaf(g1.w, gt.y, fmaf(g2.x, gtz, gtz)));
                const float v0 = fmaf(ps.rxy.y, b2, -b01.y);
                const float v1 = fmaf(-ps.rxy.x, b2, b01.x);
                const float v2 = fmaf(ps.rxy.x, b01.y, -ps.rxy.y * b01.x);
                // per knot: wt | Th00 Th10 Th20 | (Th01 Th02) (Th11 Th12) (Th21 Th22)
                constexpr int NC = (10 * K + 3) / 4 * 4;
                float c[NC];
#pragma unroll
                for (int q = 0; q < NC / 4; ++q)
                {
                    const float4 v = rec.q(kRecGeom / 4 + q);
                    c[4 * q] = v.x, c[4 * q + 1] = v.y, c[4 * q + 2] = v.z, c[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < K; ++j)
                {
                    const float *k = c + 10 * j;
                    J[OFF + j][0] = fma2(bc(k[0]), gt, J[OFF + j][0]);
                    J[OFF + j][1].x = fmaf(k[0], gtz, J[OFF + j][1].x);
                    J[OFF + j][1].y = fmaf(v0, k[1], fmaf(v1, k[2], fmaf(v2, k[3], J[OFF + j][1].y)));
                    J[OFF + j][2] = fma2(bc(v0), f2(k[4], k[5]), fma2(bc(v1), f2(k[6], k[7]), fma2(bc(v2), f2(k[8], k[9]), J[OFF + j][2])));
                }
            }
        }

        // ---- Hessian-pass sample in two halves (MBAVO_PIPELINE, patch texels) ------------------------------------------------
        // ptxas keeps the texel load of sample_step right in front of its first use (17 instructions), so a warp has ONE gather in
        // flight and every L1 miss (31 % of the sectors) is paid in full: long-scoreboard is the largest stall of the sweep kernel
        // (profiles/r2w_sweep_kernel_c3_patch_texel_ncu_raw.csv).  Split at the load, the sample loop is software-pipelined by
        // hand: front(i + 1) — geometry, tap address, LDG — is issued BEFORE back(i) — texel decode, chain rule, accumulation —
        // so the gather of the next sample flies under ~90 instructions of the current one.  Same arithmetic in the same order as
        // sample_step (bit-identical sums); what crosses the split is 14 registers, and the two rotation groups of the record
        // are read from shared memory a second time.
        struct TapState
        {
            uint4 t;     // patch texel of the tap
            float dx, dy; // fractional tap offsets (the bilinear weights are rebuilt from them)
            float il, s; // 1 / lambda (0: invalid sample), (D - t_z) / lambda
#if !MBAVO_PIPE_LEAN
            float2 m01;  // ray + (R - I) ray, x and y (lean form: recomputed from the rotation groups the second half re-reads anyway)
#endif
        };
        template <class Rec>
        __device__ __forceinline__ TapState sample_front(const Rec rec, const PixelRegs &ps, const LevelDev &lv, float2 fxy)
        {
            const float4 g0 = rec.q(0), g1 = rec.q(1), g2 = rec.q(2);
            const float2 A01 = f2(fmaf(g0.x, ps.rxy.x, fmaf(g0.y, ps.rxy.y, g1.z)), fmaf(g0.z, ps.rxy.x, fmaf(g0.w, ps.rxy.y, g1.w)));
            const float A2 = fmaf(g1.x, ps.rxy.x, fmaf(g1.y, ps.rxy.y, g2.x));
            const float2 m01 = add2(ps.rxy, A01);
            const float m2 = 1.0f + A2;
            const float il_raw = rcp_approx(m2);
            const float tau = g2.y * ps.iD;
            const float2 num = fma2(bc(-tau), m01, fma2(bc(-A2), ps.rxy, A01));
            const float2 duv = mul2(fxy, fma2(num, bc(il_raw), mul2(f2(g2.z, g2.w), bc(ps.iD))));
#if MBAVO_PIPE_LEAN >= 2
            // fewer registers live across the loop: the upper bounds are rebuilt from the lower ones (exact small integers)
            const bool ok = duv.x >= ps.lox && duv.x <= (float)(lv.W - 1) + ps.lox && duv.y >= ps.loy && duv.y <= (float)(lv.H - 1) + ps.loy;
#else
            const bool ok = duv.x >= ps.lox && duv.x <= ps.hix && duv.y >= ps.loy && duv.y <= ps.hiy;
#endif
            TapState st;
            st.il = ok ? il_raw : 0.f;
            int xo, yo;
            const float xf = floor_to_int(duv.x, xo), yf = floor_to_int(duv.y, yo);
            st.dx = duv.x - xf, st.dy = duv.y - yf; // exact
            const int idx = ok ? (ps.Y + yo) * lv.W + ps.X + xo : 0;
            st.t = __ldg(lv.ref_pair + idx);
            st.s = (ps.D - g2.y) * st.il;
#if !MBAVO_PIPE_LEAN
            st.m01 = m01;
#endif
            return st;
        }
        template <int K, int NK, int OFF, class Rec>
        __device__ __forceinline__ void sample_back(const Rec rec, const PixelRegs &ps, float2 fxy, const TapState &st, float &sumI, float2 (&J)[NK][3])
        {
            // an invalid sample (il == 0; a valid one has 1 / lambda > 0) runs with zero weights
            const bool ok = st.il != 0.f;
            const float dxdy = st.dx * st.dy;
            const float w00 = ok ? 1.0f - st.dx - st.dy + dxdy : 0.f, w01 = ok ? st.dx - dxdy : 0.f, w10 = ok ? st.dy - dxdy : 0.f, w11 = ok ? dxdy : 0.f;
            const uint4 t = st.t;
            const float r0c1 = biased_byte<1>(t.x), r0c2 = biased_byte<2>(t.x);
            const float r1c0 = biased_byte<0>(t.y), r1c1 = biased_byte<1>(t.y), r1c2 = biased_byte<2>(t.y), r1c3 = biased_byte<3>(t.y);
            const float r2c0 = biased_byte<0>(t.z), r2c1 = biased_byte<1>(t.z), r2c2 = biased_byte<2>(t.z), r2c3 = biased_byte<3>(t.z);
            const float r3c1 = biased_byte<1>(t.w), r3c2 = biased_byte<2>(t.w);
            const float2 g00 = f2(r1c2 - r1c0, r2c1 - r0c1), g01 = f2(r1c3 - r1c1, r2c2 - r0c2);
            const float2 g10 = f2(r2c2 - r2c0, r3c1 - r1c1), g11 = f2(r2c3 - r2c1, r3c2 - r1c2);
            const float2 i0 = f2(r1c1 - 8388608.0f, r1c2 - 8388608.0f), i1 = f2(r2c1 - 8388608.0f, r2c2 - 8388608.0f);
            sumI += blend4(w00, w01, w10, w11, i0.x, i0.y, i1.x, i1.y);
            const float2 gxy = fma2(bc(w00), g00, fma2(bc(w01), g01, fma2(bc(w10), g10, mul2(bc(w11), g11))));
            const float4 g0 = rec.q(0), g1 = rec.q(1);
            const float rm22 = rec.p[8];
#if MBAVO_PIPE_LEAN
            const float2 m01 = add2(ps.rxy, f2(fmaf(g0.x, ps.rxy.x, fmaf(g0.y, ps.rxy.y, g1.z)), fmaf(g0.z, ps.rxy.x, fmaf(g0.w, ps.rxy.y, g1.w))));
#else
            const float2 m01 = st.m01;
#endif
#if MBAVO_PIPE_LEAN >= 2
            const float2 gt = mul2(gxy, mul2(fxy, bc(0.5f * ps.iD))); // = PixelRegs::fxyiD, rebuilt
#else
            const float2 gt = mul2(gxy, ps.fxyiD);
#endif
            const float gtz = -(gt.x * m01.x + gt.y * m01.y) * st.il;
            const float2 b01 = mul2(bc(st.s), fma2(bc(gt.x), f2(g0.x, g0.y), fma2(bc(gt.y), f2(g0.z, g0.w), fma2(bc(gtz), f2(g1.x, g1.y), gt))));
            const float b2 = st.s * fmaf(g1.z, gt.x, fmaf(g1.w, gt.y, fmaf(rm22, gtz, gtz)));
            const float v0 = fmaf(ps.rxy.y, b2, -b01.y);
            const float v1 = fmaf(-ps.rxy.x, b2, b01.x);
            const float v2 = fmaf(ps.rxy.x, b01.y, -ps.rxy.y * b01.x);
            constexpr int NC = (10 * K + 3) / 4 * 4;
            float c[NC];
#pragma unroll
            for (int q = 0; q < NC / 4; ++q)
            {
                const float4 v = rec.q(kRecGeom / 4 + q);
                c[4 * q] = v.x, c[4 * q + 1] = v.y, c[4 * q + 2] = v.z, c[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < K; ++j)
            {
                const float *k = c + 10 * j;
                J[OFF + j][0] = fma2(bc(k[0]), gt, J[OFF + j][0]);
                J[OFF + j][1].x = fmaf(k[0], gtz, J[OFF + j][1].x);
                J[OFF + j][1].y = fmaf(v0, k[1], fmaf(v1, k[2], fmaf(v2, k[3], J[OFF + j][1].y)));
                J[OFF + j][2] = fma2(bc(v0), f2(k[4], k[5]), fma2(bc(v1), f2(k[6], k[7]), fma2(bc(v2), f2(k[8], k[9]), J[OFF + j][2])));
            }
        }

        // Samples of the segments OFF, OFF + 1, ... in time order, PH exposure phases (lane-dependent start, step PH)
        template <int K, int NK, bool PACKED, int OFF>
        struct SegmentLoop
        {
            __device__ __forceinline__ static void run(const float *__restrict__ samples_s, const int *__restrict__ seg_end_s,
                                                       int &i, int PH, const PixelRegs &ps, const LevelDev &lv, float2 fxy,
                                                       float &sumI, float2 (&J)[NK][3])
            {
                constexpr int REC = sample_rec_floats(K);
                const int end = seg_end_s[OFF];
                if constexpr (MBAVO_PIPELINE && PACKED && MBAVO_TEXEL == 3 && MBAVO_BRANCHLESS)
                {
                    if (i < end)
                    {
                        TapState cur = sample_front(SmemRec{samples_s + i * REC}, ps, lv, fxy);
                        int n = i + PH;
#pragma unroll kPipeUnroll
                        for (; n < end; n += PH)
                        {
                            const TapState nxt = sample_front(SmemRec{samples_s + n * REC}, ps, lv, fxy);
                            sample_back<K, NK, OFF>(SmemRec{samples_s + (n - PH) * REC}, ps, fxy, cur, sumI, J);
                            cur = nxt;
                        }
                        sample_back<K, NK, OFF>(SmemRec{samples_s + (n - PH) * REC}, ps, fxy, cur, sumI, J);
                        i = n;
                    }
                }
                else
                {
#pragma unroll kSampleUnroll
                for (; i < end; i += PH)
                    sample_step<K, NK, true, PACKED, OFF>(SmemRec{samples_s + i * REC}, ps, lv, fxy, sumI, J);
                }
                if constexpr (OFF + 1 <= NK - K)
                    SegmentLoop<K, NK, PACKED, (OFF + 1 <= NK - K ? OFF + 1 : OFF)>::run(samples_s, seg_end_s, i, PH, ps, lv, fxy,
                                                                                      sumI, J);
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

        // ---- device-resident Gauss-Newton step (one warp of the last block) -------------------------------------------------
        // computeTrustRegionStep (blur_aware_direct_tracker.cpp:799-831) + Plus_t / Plus_R (Spline.h:307-330) on the packed
        // window vector v = [cost, g(D), triu(H) row-major], D = 6 NK: damp the diagonal by (1 + 1/radius), solve by LDL^T
        // (same pivot test as the host's fast path, lm_driver.cpp), step = -H^-1 g, model decrease with the damped H, candidate
        // = knots [+] step.  Knots outside the window keep a zero step (what the pseudo-inverse of the full system gives).
        // A (2 D x D) and w (3 D) are shared-memory scratch.  Executed by ONE warp; right-looking factorisation so that every
        // step is a handful of parallel FMAs instead of long serial dot products.
        // fp64 reciprocal for the pivots: float seed (24 bits) + two Newton steps in FMA form (error at the fp64 rounding
        // level); an IEEE division costs about twice the latency and the factorisation is a chain of D of them
        __device__ __forceinline__ double rcp_newton(double d)
        {
#if MBAVO_RCP64
            // rcp.approx.ftz.f64 (MUFU.RCP64H): ~20 good bits from the upper word of d in ONE instruction; two Newton steps square
            // the error twice (2^-20 -> 2^-40 -> 2^-80).  The float route costs two conversions, a range check and a float
            // division on the chain of every pivot.  (d <= 0 / NaN: garbage in, flagged by the pivot test of the callers)
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#else
            double r = (double)(1.0f / (float)d);
#endif
            r = fma(r, fma(-d, r, 1.0), r);
            r = fma(r, fma(-d, r, 1.0), r);
            return r;
        }

        // Register-resident LDL^T solve of the damped window system for D = 6 NK <= 32: lane i owns row i of the lower
        // triangle, columns are exchanged by warp shuffles, every loop is unrolled at compile time.  On return y (shared)
        // holds H^-1 g; *ok_out tells whether every pivot passed the host's test.  Executed by one warp.
        template <int D>
        __device__ __noinline__ void ldlt_solve_regs(const double *__restrict__ Hd, const double *__restrict__ g,
                                                     double *__restrict__ Ls, double *__restrict__ y, bool *ok_out)
        {
            const int lane = threadIdx.x & 31;
            const int row = lane < D ? lane : D - 1; // lanes >= D shadow the last row (their results are ignored)
            double a[D];
#pragma unroll
            for (int c = 0; c < D; ++c)
                a[c] = Hd[row * D + c];
            bool ok = true;
            double dmax = 0.0, yi = g[row], idiag = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j)
            {
                const double d = __shfl_sync(0xffffffffu, a[j], j);
                dmax = d > dmax ? d : dmax;
                if (!(d > 1e-9 * dmax))
                    ok = false;
                const double id = rcp_newton(d);
                const double aj = a[j];           // A(row, j) before scaling
                const double l = aj * id;         // L(row, j)
#pragma unroll
                for (int k = j + 1; k < D; ++k)
                {
                    const double akj = __shfl_sync(0xffffffffu, aj, k); // A(k, j) before scaling
                    a[k] -= l * akj;              // A(row, k) -= L(row, j) d_j L(k, j)   (only row >= k is used later)
                }
                if (row > j)
                    a[j] = l;
                if (row == j)
                    idiag = id;
            }
            // L to shared memory for the transposed sweep
#pragma unroll
            for (int c = 0; c < D; ++c)
                if (lane < D)
                    Ls[lane * D + c] = a[c];
            // forward: L y = g
#pragma unroll
            for (int k = 0; k < D; ++k)
            {
                const double yk = __shfl_sync(0xffffffffu, yi, k);
                if (row > k)
                    yi -= a[k] * yk;
            }
            yi *= idiag;
            __syncwarp();
            // backward: L^T x = y
#pragma unroll
            for (int k = D - 1; k >= 0; --k)
            {
                const double xk = __shfl_sync(0xffffffffu, yi, k);
                if (row < k)
                    yi -= Ls[k * D + row] * xk;
            }
            if (lane < D)
                y[lane] = yi;
            *ok_out = ok;
            __syncwarp();
        }

        // Compact shared-memory LDL^T solve of the damped window system for 12 < D <= 32, one warp, lane = row: a runtime loop over
        // the pivots (code that stays in the instruction cache, unlike the fully unrolled register form, whose 40 KB of straight-line
        // code executed once per level costs more in fetches than it saves in arithmetic).  A (D x D, the damped H) keeps its columns
        // unscaled — lane i reads A(k, j) of the rows above its own while updating its row — and the unit-lower factor goes to Lm.
        // On return y holds H^-1 g; same pivot test as the host's fast path.
        template <int D>
        __device__ __noinline__ void ldlt_solve_rows(double *__restrict__ A, const double *__restrict__ g, double *__restrict__ Lm,
                                                     double *__restrict__ y, bool *ok_out)
        {
            const int lane = threadIdx.x & 31;
            const int i = lane < D ? lane : D - 1; // lanes >= D shadow the last row without storing
            const bool mine = lane < D;
            bool ok = true;
            double dmax = 0.0, idiag = 0.0;
            for (int j = 0; j < D; ++j)
            {
                const double d = A[j * D + j];
                dmax = d > dmax ? d : dmax;
                if (!(d > 1e-9 * dmax))
                    ok = false;
                const double id = rcp_newton(d);
                if (i == j)
                    idiag = id;
                if (i > j)
                {
                    const double l = A[i * D + j] * id; // L(i, j)
                    if (mine)
                        Lm[i * D + j] = l;
#pragma unroll 4
                    for (int k = j + 1; k <= i; ++k) // A(i, k) -= L(i, j) d_j L(k, j) = l * A(k, j)
                    {
                        const double v = A[i * D + k] - l * A[k * D + j];
                        if (mine)
                            A[i * D + k] = v;
                    }
                }
                __syncwarp();
            }
            // forward L z = g, scale by 1 / d, backward L^T x = z: the running value of row i lives in a register, finished
            // entries travel by shuffle
            double yi = g[i];
            for (int k = 0; k < D; ++k)
            {
                const double yk = __shfl_sync(0xffffffffu, yi, k);
                if (i > k)
                    yi -= Lm[i * D + k] * yk;
            }
            yi *= idiag;
            for (int k = D - 1; k >= 0; --k)
            {
                const double xk = __shfl_sync(0xffffffffu, yi, k);
                if (i < k)
                    yi -= Lm[k * D + i] * xk;
            }
            if (mine)
                y[lane] = yi;
            *ok_out = ok;
            __syncwarp();
        }

        // ---- block-wide set-up and factorisation of the damped window system (persistent sweep: the whole last block is there anyway,
        // and the one-warp forms spend 1.3 us building the system and 6 us factoring it for an 18 x 18 window) ----------------------
        // A (D x D), Hd (D x D) = the damped H, g = y = the gradient; every thread of the block takes elements
        template <int NK>
        __device__ __forceinline__ void gn_build_block(const double *__restrict__ v, double radius, double *__restrict__ A, double *__restrict__ w)
        {
            constexpr int D = 6 * NK, D1 = D + 1;
            double *Hd = A + D * D, *g = w, *y = w + D;
            const double damp = 1.0 / radius;
            for (int e = threadIdx.x; e < D * D; e += blockDim.x)
            {
                const int r = e / D, c = e % D, ra = (r < c ? r : c) + 1, cb = (r < c ? c : r) + 1;
                double h = v[ra * D1 - ra * (ra - 1) / 2 + (cb - ra)];
                if (r == c)
                    h += h * damp;
                A[e] = h, Hd[e] = h;
            }
            for (int e = threadIdx.x; e < D; e += blockDim.x)
                g[e] = v[1 + e], y[e] = v[1 + e];
        }
        // LDL^T with ONE THREAD PER ELEMENT of the lower triangle (D (D + 1) / 2 threads of the block, the others return at once):
        // per pivot, column j is published to shared memory (double-buffered: one named barrier per pivot), then every element (i, k),
        // k > j, takes its rank-1 update.  A receives L below the diagonal and d on it.  ~150 cycles per pivot, no unrolled code.
        template <int D>
        __device__ __noinline__ void ldlt_factor_block(double *__restrict__ A, double *__restrict__ col, int *__restrict__ ok_out)
        {
            constexpr int NEL = D * (D + 1) / 2, NTH = (NEL + 31) / 32 * 32;
            const int t = threadIdx.x;
            if (t >= NTH)
                return;
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= t)
                ++i;
            const int k = t - i * (i + 1) / 2;
            const bool active = t < NEL;
            double a = active ? A[i * D + k] : 0.0;
            double dmax = 0.0;
            bool ok = true;
            for (int j = 0; j < D; ++j)
            {
                double *buf = col + (j & 1) * (D + 2);
                if (active && k == j)
                    buf[i] = a; // column j, pivot included
                asm volatile("bar.sync 1, %0;" ::"r"(NTH) : "memory");
                const double d = buf[j];
                dmax = d > dmax ? d : dmax;
                if (!(d > 1e-9 * dmax))
                    ok = false;
                const double id = rcp_newton(d);
                if (active && k >= j)
                {
                    const double l = buf[i] * id; // L(i, j)
                    if (k > j)
                        a -= l * buf[k];          // A(i, k) -= L(i, j) d_j L(k, j)
                    else if (i > j)
                        A[i * D + j] = l;
                    else
                        A[j * D + j] = d;
                }
            }
            if (t == 0)
                *ok_out = ok ? 1 : 0;
        }

        // PREFACTORED: gn_build_block + ldlt_factor_block have run (A holds the factor, *prefactored_ok its pivot verdict)
        template <int NK, bool PREFACTORED = false>
        // knots_s (nullable, shared memory, 2 x 112 doubles): the knots the step starts from are read from its first half instead
        // of GnState, and the candidate is ALSO left in its second half (persistent sweep: the pose records follow at once)
        __device__ void gn_solve_step(const double *__restrict__ v, const GnParams &gp, double *__restrict__ A, double *__restrict__ w,
                                      unsigned long long *ts = nullptr, const int *prefactored_ok = nullptr, double *knots_s = nullptr)
        {
#ifdef MBAVO_PROFILE_PHASES
#define MBAVO_TS(k) do { if (ts && (threadIdx.x & 31) == 0) ts[k] = global_timer_ns(); } while (0)
#else
#define MBAVO_TS(k) do { } while (0)
#endif
            constexpr int D = 6 * NK, D1 = D + 1;
            const int lane = threadIdx.x & 31;
            GnState *st = gp.state;
            double *Hd = A + D * D;                 // the damped H (kept for the model decrease); A becomes its LDL^T factor
            double *g = w, *y = w + D, *sv = w + 2 * D;
            const double damp = 1.0 / gp.radius;
            if constexpr (!PREFACTORED)
            {
                for (int e = lane; e < D * D; e += 32)
                {
                    const int r = e / D, c = e % D, ra = (r < c ? r : c) + 1, cb = (r < c ? c : r) + 1;
                    double h = v[ra * D1 - ra * (ra - 1) / 2 + (cb - ra)]; // element (ra, cb) of the (D+1) x (D+1) upper triangle
                    if (r == c)
                        h += h * damp;
                    A[e] = h, Hd[e] = h;
                }
                for (int e = lane; e < D; e += 32)
                    g[e] = v[1 + e], y[e] = v[1 + e];
                __syncwarp();
            }
            (void)damp, (void)D1;
            MBAVO_TS(11);
            bool ok = true;
            if constexpr (PREFACTORED)
            {
                // forward L z = g, scale by 1 / d, backward L^T x = z with the factor in A: lane = row, finished entries by shuffle
                ok = *prefactored_ok != 0;
                const int i = lane < D ? lane : D - 1;
                double yi = g[i];
                for (int k = 0; k < D; ++k)
                {
                    const double yk = __shfl_sync(0xffffffffu, yi, k);
                    if (i > k)
                        yi -= A[i * D + k] * yk;
                }
                yi *= rcp_newton(A[i * D + i]);
                for (int k = D - 1; k >= 0; --k)
                {
                    const double xk = __shfl_sync(0xffffffffu, yi, k);
                    if (i < k)
                        yi -= A[k * D + i] * xk;
                }
                if (lane < D)
                    y[lane] = yi;
                __syncwarp();
            }
            else
            {
#ifdef MBAVO_SMEM_SOLVE
            constexpr int kRegSolveMaxD = 0;
#else
            constexpr int kRegSolveMaxD = 24; // windows of 3 and 4 knots take the lane-per-row shared-memory form (ldlt_solve_rows); the
                                              // right-looking form below takes 10.5 us for D = 18 (3 warp barriers + runtime-index
                                              // updates per pivot), the fully unrolled register form 6 us (instruction fetch)
#endif
#ifndef MBAVO_SOLVE_ROWS
#define MBAVO_SOLVE_ROWS 0 // 1: windows of 3 and 4 knots take ldlt_solve_rows (measured slower: 9.5 vs 6.0 us for D = 18)
#endif
            if constexpr (D <= (MBAVO_SOLVE_ROWS ? 12 : kRegSolveMaxD))
            {
                ldlt_solve_regs<D>(Hd, g, A, y, &ok);
            }
            else if constexpr (D <= kRegSolveMaxD)
            {
                // A holds a copy of the damped H (Hd stays intact for the model decrease); the factor goes behind the three vectors
                ldlt_solve_rows<D>(A, g, w + 3 * D + 2, y, &ok);
            }
            else
            {
                // right-looking LDL^T: after step j column j of A holds L(:, j) below the diagonal and d_j on it
                double dmax = 0.0;
                for (int j = 0; j < D; ++j)
                {
                    const double d = A[j * D + j];
                    dmax = d > dmax ? d : dmax;
                    if (!(d > 1e-9 * dmax))
                        ok = false;
                    const double id = 1.0 / d;
                    __syncwarp();
                    const int m = D - j - 1; // trailing size
                    for (int e = lane; e < m * m; e += 32)
                    {
                        const int i = j + 1 + e / m, k = j + 1 + e % m;
                        if (k <= i) // lower triangle only
                            A[i * D + k] -= A[i * D + j] * A[k * D + j] * id;
                    }
                    __syncwarp();
                    for (int i = j + 1 + lane; i < D; i += 32)
                        A[i * D + j] *= id; // L(i, j)
                    __syncwarp();
                }
                // L y = g (column sweeps), y /= d, L^T x = y
                for (int k = 0; k < D; ++k)
                {
                    const double yk = y[k];
                    __syncwarp();
                    for (int i = k + 1 + lane; i < D; i += 32)
                        y[i] -= A[i * D + k] * yk;
                    __syncwarp();
                }
                for (int e = lane; e < D; e += 32)
                    y[e] /= A[e * D + e];
                __syncwarp();
                for (int k = D - 1; k >= 0; --k)
                {
                    const double xk = y[k];
                    __syncwarp();
                    for (int i = lane; i < k; i += 32)
                        y[i] -= A[k * D + i] * xk;
                    __syncwarp();
                }
            }
            }
            MBAVO_TS(12);
            for (int e = lane; e < D; e += 32)
                sv[e] = -y[e]; // solve_normal_equation.h:33
            __syncwarp();
            // model decrease -(g^T s + s^T H s / 2) with the damped H (tracker.cpp:821-823)
            double part = 0.0;
            for (int r = lane; r < D; r += 32)
            {
                double hr = 0.0;
                for (int c = 0; c < D; ++c)
                    hr += Hd[r * D + c] * sv[c];
                part += g[r] * sv[r] + 0.5 * sv[r] * hr;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                part += __shfl_xor_sync(0xffffffffu, part, o);
            const double model = -part;
            MBAVO_TS(13);
            const int n = gp.n_knots;
            // full-ordering step and candidate knots
            for (int e = lane; e < 6 * n; e += 32)
                st->step[e] = 0.0;
            __syncwarp();
            for (int e = lane; e < 3 * NK; e += 32)
            {
                st->step[3 * gp.kmin + e] = sv[e];
                st->step[3 * n + 3 * gp.kmin + e] = sv[3 * NK + e];
            }
            const double *cur_t = knots_s ? knots_s : st->cur_t, *cur_R = knots_s ? knots_s + 48 : st->cur_R;
            for (int e = lane; e < 3 * n; e += 32)
            {
                const int a = e / 3 - gp.kmin;
                const double ct = cur_t[e] + ((a >= 0 && a < NK) ? sv[3 * a + e % 3] : 0.0);
                st->cand_t[e] = ct;
                if (knots_s)
                    knots_s[112 + e] = ct;
            }
            for (int j = lane; j < n; j += 32)
            {
                const int a = j - gp.kmin;
                const double *q = cur_R + 4 * j;
                double dq[4] = {0, 0, 0, 1};
#if MBAVO_FAST_FINISH
                if (a >= 0 && a < NK)
                {
                    // Sophus::SO3d::exp(w).unit_quaternion() (Spline.h:326): power series for the small steps of a tracker, closed
                    // form otherwise (so3_exp_only, pose_device.cuh; equal to lm_driver.cpp so3_exp at the rounding level)
                    const Q e = so3_exp_only(sv + 3 * NK + 3 * a);
                    dq[0] = e.x, dq[1] = e.y, dq[2] = e.z, dq[3] = e.w;
                }
#else
                if (a >= 0 && a < NK)
                {
                    // Sophus::SO3d::exp(w).unit_quaternion() (Spline.h:326), as lm_driver.cpp so3_exp
                    const double *om = sv + 3 * NK + 3 * a;
                    const double t2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
                    double fi, fr;
                    if (t2 < 1e-20)
                    {
                        const double t4 = t2 * t2;
                        fi = 0.5 - t2 / 48.0 + t4 / 3840.0;
                        fr = 1.0 - t2 / 8.0 + t4 / 384.0;
                    }
                    else
                    {
                        const double t = sqrt(t2);
                        double sh, ch;
                        sincos(0.5 * t, &sh, &ch);
                        fi = sh / t;
                        fr = ch;
                    }
                    dq[0] = fi * om[0], dq[1] = fi * om[1], dq[2] = fi * om[2], dq[3] = fr;
                }
#endif
                double o[4];
                o[0] = q[3] * dq[0] + q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1];
                o[1] = q[3] * dq[1] + q[1] * dq[3] + q[2] * dq[0] - q[0] * dq[2];
                o[2] = q[3] * dq[2] + q[2] * dq[3] + q[0] * dq[1] - q[1] * dq[0];
                o[3] = q[3] * dq[3] - q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    st->cand_R[4 * j + c] = o[c];
                    if (knots_s)
                        knots_s[112 + 48 + 4 * j + c] = o[c];
                }
            }
            MBAVO_TS(14);
            if (lane == 0)
            {
                st->cost = v[0];
                st->model = model;
                if (!ok || !(model == model))
                    st->status = 1;
                else if (model < 0.0 && st->status == 0)
                    st->status = 2;
            }
        }

        // tile id -> (ta << 8) | tb, row-major upper triangle of the T x T tile grid
        __device__ __forceinline__ unsigned int tile_s_copy(int id, int T)
        {
            int a = 0, rem = id;
            while (rem >= T - a)
            {
                rem -= T - a;
                ++a;
            }
            return (unsigned int)((a << 8) | (a + rem));
        }

        // Geometry of the staged rows and of the 2x2 tiling of rows^T rows
        template <int NK, bool WITH_J>
        struct RowGeom
        {
            static constexpr int D1 = WITH_J ? 6 * NK + 1 : 1;          // row length: [r | J]
            static constexpr int D1E = (D1 + 1) & ~1;                    // padded with one zero column to an even length
            static constexpr int PITCH = (D1E / 2) % 2 == 1 ? D1E : D1E + 2; // even pitch with odd PITCH/2: row writes 2-way at worst
            static constexpr int T = D1E / 2;                            // tiles per dimension
            static constexpr int NT = T * (T + 1) / 2;                   // upper-triangular tiles
            static constexpr int MT = WITH_J ? (NT + 31) / 32 : 1;       // tiles owned by one lane
            static constexpr int E = WITH_J ? packed_len(NK) : 1;        // packed upper triangle
        };

        // Persistent sweep: what a pass needs beyond its TrackParams (see SweepCtl)
        struct PersistArgs
        {
            SweepCtl *ctl;
            unsigned int target;     // ctl->done value that makes the records of this pass valid
            const EvalStage *stage;  // frame times / segment table for the candidate's sample records
            unsigned long long *t_release; // where the pass's last block stamps the globaltimer when it releases the pass
            const double *sample_u;  // shared memory: sample_u of every exposure sample (nullptr: computed on the spot)
            double *knots_s;         // shared memory, 2 x 112 doubles: [t(48) | R(64)] of the knots a step starts from / of its candidate
        };

        // spin (thread 0) until ctl->done reaches target; false: aborted or timed out (the caller leaves the kernel)
        __device__ __forceinline__ bool wait_pass(SweepCtl *ctl, unsigned int target)
        {
            __shared__ int go_s;
            if (threadIdx.x == 0)
            {
                int go = 1;
                const unsigned long long t0 = global_timer_ns();
                while ((int)(ld_acquire_gpu(&ctl->done) - target) < 0)
                {
                    if (*reinterpret_cast<volatile int *>(&ctl->abort) != 0)
                    {
                        go = 0;
                        break;
                    }
                    if (global_timer_ns() - t0 > 4000000000ull)
                    {
                        *reinterpret_cast<volatile int *>(&ctl->abort) = 1;
                        go = 0;
                        break;
                    }
                    __nanosleep(32);
                }
                go_s = go;
            }
            __syncthreads();
            const bool go = go_s != 0;
            __syncthreads();
            return go;
        }

        // spin (thread 0) until the copy engine has delivered the level's points (TrackParams::ready_flag); false: timed out
        __device__ __forceinline__ bool wait_ready(const unsigned int *flag, unsigned int epoch, SweepCtl *ctl)
        {
            __shared__ int ready_s;
            if (threadIdx.x == 0)
            {
                int go = 1;
                const unsigned long long t0 = global_timer_ns();
                while (ld_acquire_sys_u32(flag) != epoch)
                {
                    if (*reinterpret_cast<volatile int *>(&ctl->abort) != 0 || global_timer_ns() - t0 > 4000000000ull)
                    {
                        *reinterpret_cast<volatile int *>(&ctl->abort) = 1;
                        go = 0;
                        break;
                    }
                    __nanosleep(64);
                }
                ready_s = go;
            }
            __syncthreads();
            const bool go = ready_s != 0;
            __syncthreads();
            return go;
        }

        // the candidate's sample records, computed by the threads of ONE block (the pass's last block) right after the solve:
        // same arithmetic as pose_kernel, one thread per (frame, sample)
        template <int K>
        __device__ __noinline__ void pose_records_block(const EvalStage *st, const double *kt, const double *kR, int with_jacobian,
                                                        float *samples, double *mid, int *seg_end, const double *u_pre = nullptr)
        {
            const int total = st->N * st->F;
            for (int g = threadIdx.x; g < total; g += blockDim.x)
                pose_one<K>(st, kt, kR, g, with_jacobian, samples, mid, seg_end, nullptr, u_pre);
        }

        // One pass over one pyramid level.  K: knots per segment, NK: knots in the window (NK - K + 1 segments touched),
        // WITH_J: Hessian pass or cost only, PACKED: keyframe texels available, WARPS: warps per block.
        // PERSIST = false: the body of track_kernel (one launch per pass).  PERSIST = true: one pass of sweep_kernel — all
        // blocks stay resident, the pass starts when SweepCtl::done reaches pa.target and its last block releases done =
        // target + 1.  Returns false when the sweep was aborted.
        // What the LAST block of a pass does once every block has announced its partials: the fixed-order sum, the exchange with the
        // other ranks, the solve / candidate / pose records (Hessian pass) or record / commit (cost pass), the publication to the
        // host and — inside a persistent sweep — the release of the pass.  Kept out of line so that this cold, fp64-heavy code does
        // not compete for registers with the sample loops of track_pass.
#ifndef MBAVO_FINISH_NOINLINE
#define MBAVO_FINISH_NOINLINE 0 // measured: out of line, the sweep kernel loses 40 us per C3 sweep (ptxas then schedules the sample loops differently)
#endif
#if MBAVO_FINISH_NOINLINE
#define MBAVO_FINISH_ATTR __noinline__
#else
#define MBAVO_FINISH_ATTR __forceinline__
#endif
        template <int K, int NK, bool WITH_J, int WARPS, bool PERSIST>
        __device__ MBAVO_FINISH_ATTR bool pass_finish(const TrackParams &prm, unsigned char *smem_raw, const PersistArgs pa)
        {
            using G = RowGeom<NK, WITH_J>;
            constexpr int kWarpsPerBlock = WARPS, kThreads = kWarpsPerBlock * 32;
            constexpr int E = G::E;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            (void)lane;
            double *red_s = reinterpret_cast<double *>(smem_raw);
            const int num_blocks = gridDim.x * gridDim.y;
            const double inv_num_residuals = prm.inv_num_residuals;
            MBAVO_STAMP(8);
            // last block: sum the partials of all blocks in a fixed order.  GRP adjacent lanes share one element: lane `part`
            // sums the blocks b = part, part + GRP, ... (32 loads in flight), the GRP partial sums are combined by a fixed
            // xor-shuffle tree.  Deterministic: the order depends only on the grid size.
            __threadfence();
            // (persistent Hessian pass) the knots the step will start from: the load travels under the sum of the partials
            constexpr bool kKnotsInSmem = PERSIST && WITH_J && MBAVO_FAST_FINISH;
            double knot_pre = 0.0;
            if constexpr (kKnotsInSmem)
            {
                if (prm.gn.state && threadIdx.x < 112)
                    knot_pre = __ldcg(prm.gn.state->cur_t + threadIdx.x); // cur_t (48) and cur_R (64) are adjacent in GnState
            }
            // (third cut) the record-buffer selector for the candidate's records (Hessian pass); what the commit of a cost pass reads —
            // the level's cost, model decrease and status, the knots and the candidate, four elements per lane of warp 0
            constexpr bool kPre3 = PERSIST && MBAVO_FINISH3;
            int cur_buf_pre = 0, status_pre = 0;
            double cost_pre = 0.0, model_pre = 0.0, cand_pre[4] = {0.0, 0.0, 0.0, 0.0}, cur_pre[4] = {0.0, 0.0, 0.0, 0.0};
            if constexpr (kPre3)
            {
                if (prm.gn.state)
                {
                    const GnState *st = prm.gn.state;
                    if constexpr (WITH_J)
                        cur_buf_pre = __ldcg(&st->cur_buf);
                    else if (threadIdx.x < 32)
                    {
                        cost_pre = __ldcg(&st->cost), model_pre = __ldcg(&st->model), status_pre = __ldcg(&st->status);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                        {
                            const int e = (int)threadIdx.x + 32 * u;
                            const int idx = e < 3 * prm.gn.n_knots ? e : 48 + (e - 3 * prm.gn.n_knots);
                            if (e < 7 * prm.gn.n_knots)
                                cand_pre[u] = __ldcg(st->cand_t + idx), cur_pre[u] = __ldcg(st->cur_t + idx);
                        }
                    }
                }
            }
            double *fin_s = red_s; // this rank's vector, scaled by 1 / num_residuals (global when sharded)
            constexpr int EP = (E + 1) & ~1; // row pitch of the per-block partials (even: rows are 16-byte aligned)
            if constexpr (MBAVO_WIDE_SUM && E > 1)
            {
                // Hessian pass: a thread owns a PAIR of elements (one 128-bit load per block row) and one of PARTS interleaved subsets
                // of the rows, with up to 24 loads in flight — the 148 x 190 partials of a C3 pass are ONE round trip to L2 per thread
                // instead of four dependent ones; the PARTS partial sums are then combined in part order through shared memory.
                // Deterministic: the order depends only on the grid size and the block shape.
                constexpr int COLS = EP / 2;
                constexpr int PARTS = kThreads / COLS >= 8 ? 8 : (kThreads / COLS >= 1 ? kThreads / COLS : 1);
                constexpr int U = 24;
                double *part_s = fin_s + EP; // [PARTS][EP]; aliases the solve scratch, which is touched only after this sum
                for (int c0 = 0; c0 < COLS; c0 += kThreads / PARTS)
                {
                    const int col = c0 + threadIdx.x % (kThreads / PARTS), part = threadIdx.x / (kThreads / PARTS);
                    if (col < COLS && part < PARTS)
                    {
                        const double2 *src = reinterpret_cast<const double2 *>(prm.block_partials) + col;
                        double2 acc = make_double2(0.0, 0.0);
                        int b = part;
                        for (; b + (U - 1) * PARTS < num_blocks; b += U * PARTS)
                        {
                            double2 v[U];
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                v[u] = __ldcg(src + (size_t)(b + u * PARTS) * COLS);
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                acc.x += v[u].x, acc.y += v[u].y;
                        }
                        for (; b + 3 * PARTS < num_blocks; b += 4 * PARTS)
                        {
                            double2 v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                v[u] = __ldcg(src + (size_t)(b + u * PARTS) * COLS);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                acc.x += v[u].x, acc.y += v[u].y;
                        }
                        for (; b < num_blocks; b += PARTS)
                        {
                            const double2 v = __ldcg(src + (size_t)b * COLS);
                            acc.x += v.x, acc.y += v.y;
                        }
                        part_s[part * EP + 2 * col] = acc.x, part_s[part * EP + 2 * col + 1] = acc.y;
                    }
                }
                __syncthreads();
                for (int e = threadIdx.x; e < E; e += kThreads)
                {
                    double s = part_s[e];
#pragma unroll
                    for (int q = 1; q < PARTS; ++q)
                        s += part_s[q * EP + e];
                    fin_s[e] = s * inv_num_residuals;
                }
            }
            else
            {
            constexpr int GRP = E >= kThreads ? 1 : (kThreads / E >= 32 ? 32 : (kThreads / E >= 16 ? 16 : (kThreads / E >= 8 ? 8 : (kThreads / E >= 4 ? 4 : (kThreads / E >= 2 ? 2 : 1)))));
            for (int e0 = 0; e0 < E; e0 += kThreads / GRP)
            {
                const int e = e0 + threadIdx.x / GRP, part = threadIdx.x % GRP;
                double s = 0.0;
                if (e < E)
                {
                    const double *src = prm.block_partials + e;
                    int b = part;
                    for (; b + 31 * GRP < num_blocks; b += 32 * GRP)
                    {
                        double v[32];
#pragma unroll
                        for (int u = 0; u < 32; ++u)
                            v[u] = __ldcg(src + (size_t)(b + u * GRP) * EP);
#pragma unroll
                        for (int u = 0; u < 32; ++u)
                            s += v[u];
                    }
                    for (; b + 7 * GRP < num_blocks; b += 8 * GRP)
                    {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[u] = __ldcg(src + (size_t)(b + u * GRP) * EP);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            s += v[u];
                    }
                    for (; b < num_blocks; b += GRP)
                        s += __ldcg(src + (size_t)b * EP);
                }
#pragma unroll
                for (int o = 1; o < GRP; o <<= 1)
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                if (e < E && part == 0)
                    fin_s[e] = s * inv_num_residuals;
            }
            }
            if constexpr (kKnotsInSmem)
            {
                if (threadIdx.x < 112)
                    pa.knots_s[threadIdx.x] = knot_pre;
            }
            __syncthreads();
            MBAVO_STAMP(9);
            const ShardParams &sh = prm.shard;
            if (sh.world > 1)
            {
                // one-shot all-reduce over NVLink: every element goes into the slot [parity][my rank] of EVERY rank's mailbox as one
                // 16-byte store of two self-validating words; then the same thread polls the W copies of its element in its own
                // mailbox and sums them in rank order (identical bits on every rank).  No fence, no flag: one NVLink traversal.
                const int par = (int)(sh.seq & 1ull);
                const unsigned long long tagv = publish_tag(sh.seq), tag = tagv << 32;
                for (int e = threadIdx.x; e < E; e += blockDim.x)
                {
                    const unsigned long long b = (unsigned long long)__double_as_longlong(fin_s[e]);
                    const unsigned long long w0 = tag | (b & 0xffffffffull), w1 = tag | (b >> 32);
                    for (int r = 0; r < sh.world; ++r)
                        st_sys_v2(&sh.peer[r]->slot[par][sh.rank][e], w0, w1);
                }
                const Mailbox *mine = sh.peer[sh.rank];
                for (int e = threadIdx.x; e < E; e += blockDim.x)
                {
                    double s = 0.0;
                    bool ok = true;
                    for (int r = 0; r < sh.world && ok; ++r)
                    {
                        ulonglong2 w = ld_sys_v2(&mine->slot[par][r][e]);
                        if ((w.x >> 32) != tagv || (w.y >> 32) != tagv)
                        {
                            const unsigned long long t0 = global_timer_ns();
                            do
                            {
                                __nanosleep(20);
                                w = ld_sys_v2(&mine->slot[par][r][e]);
                                if (global_timer_ns() - t0 > 4000000000ull) // a peer that never arrives must not hang the GPU
                                {
                                    ok = false;
                                    break;
                                }
                            } while ((w.x >> 32) != tagv || (w.y >> 32) != tagv);
                        }
                        s += __longlong_as_double((long long)((w.x & 0xffffffffull) | (w.y << 32)));
                    }
                    fin_s[e] = ok ? s : __longlong_as_double(0x7ff8000000000000ll); // NaN: a peer timed out
                }
                __syncthreads();
            }
            const GnParams &gp = prm.gn;
            if (gp.state)
            {
                // device-resident Gauss-Newton sweep: solve / record on the spot, publish the level's scalars
                GnState *st = gp.state;
                if constexpr (WITH_J)
                {
                    double *A = fin_s + ((E + 1) & ~1), *w = A + 72 * NK * NK;
#ifndef MBAVO_BLOCK_SOLVE
#define MBAVO_BLOCK_SOLVE 1
#endif
                    if constexpr (PERSIST && MBAVO_BLOCK_SOLVE && NK >= 3 && NK <= 4)
                    {
                        // the whole block builds and factors the damped system; one warp finishes (substitutions, model decrease, candidate)
                        __shared__ int factor_ok_s;
                        gn_build_block<NK>(fin_s, gp.radius, A, w);
                        __syncthreads();
                        ldlt_factor_block<6 * NK>(A, w + 18 * NK + 2, &factor_ok_s);
                        __syncthreads();
                        if (warp == 0)
                            gn_solve_step<NK, true>(fin_s, gp, A, w, prm.phase_times ? prm.phase_times + 16 * (prm.trace_row & 63) : nullptr,
                                                    &factor_ok_s, kKnotsInSmem ? pa.knots_s : nullptr);
                    }
                    else if (warp == 0)
                        gn_solve_step<NK>(fin_s, gp, A, w, prm.phase_times ? prm.phase_times + 16 * (prm.trace_row & 63) : nullptr, nullptr,
                                          kKnotsInSmem ? pa.knots_s : nullptr);
                    if constexpr (PERSIST)
                    {
                        // the candidate's sample records (what the stand-alone pose kernel computes between the two passes of a
                        // level when every pass is its own launch), with Jacobians: the next level may stand on them
                        __syncthreads();
                        // into the record buffer the sweep does NOT stand on (a sweep that never commits keeps cur_buf = 0)
                        const int cb = 1 - (kPre3 ? cur_buf_pre : *reinterpret_cast<volatile int *>(&st->cur_buf));
                        // pose part only: all the cost pass needs.  The Jacobian part, which only a finer level standing on the
                        // committed candidate reads, is computed by the service block WHILE the cost pass runs (below).
                        pose_records_block<K>(pa.stage, kKnotsInSmem ? pa.knots_s + 112 : st->cand_t, kKnotsInSmem ? pa.knots_s + 160 : st->cand_R,
                                              gridDim.x > 1 ? 0 : 1,
                                              const_cast<float *>(prm.samples) + (size_t)cb * prm.samples_stride,
                                              const_cast<double *>(prm.mid) + (size_t)cb * prm.mid_stride,
                                              const_cast<int *>(prm.seg_end) + (size_t)cb * prm.seg_end_stride,
                                              kKnotsInSmem ? pa.sample_u : nullptr);
                        MBAVO_STAMP(15);
                    }
                }
                else if (threadIdx.x < 32)
                {
                    // record the candidate's cost, commit the candidate when it lowered the cost (the finer level then starts from
                    // it), publish the level's scalars (and, after the last level, the knots) to the host.  One warp: the knots are
                    // copied / published a lane per element.  Inside a persistent sweep the pass is released to the other blocks
                    // BEFORE anything is stored to host memory (a fence behind PCIe stores costs microseconds).
                    const int lane32 = threadIdx.x;
                    const double cc = fin_s[0], cost0 = kPre3 ? cost_pre : st->cost, model0 = kPre3 ? model_pre : st->model;
                    const int status0 = kPre3 ? status_pre : st->status;
                    const bool commit = gp.chain && status0 == 0 && cc < cost0;
                    const int nk7 = 7 * gp.n_knots; // cur_t (3n) and cur_R (4n) are adjacent in GnState
                    double *cur = st->cur_t;
                    const double *cand = st->cand_t;
                    static_assert(offsetof(GnState, cur_R) == offsetof(GnState, cur_t) + sizeof(double) * 3 * 16 &&
                                      offsetof(GnState, cand_R) == offsetof(GnState, cand_t) + sizeof(double) * 3 * 16,
                                  "GnState: rotations follow translations");
                    double keep[4]; // the knots the sweep stands on after this level, 4 elements per lane (7 * 16 <= 128)
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                    {
                        const int e = lane32 + 32 * u;
                        const int idx = e < 3 * gp.n_knots ? e : 48 + (e - 3 * gp.n_knots); // element of the padded [t(48) | R(64)] layout
                        keep[u] = 0.0;
                        if (e < nk7)
                        {
                            keep[u] = kPre3 ? (commit ? cand_pre[u] : cur_pre[u]) : (commit ? cand[idx] : cur[idx]);
                            if (commit)
                                cur[idx] = keep[u];
                        }
                    }
                    if (lane32 == 0)
                    {
                        st->cand_cost = cc;
                        if (commit)
                            st->cur_buf ^= 1; // the candidate's sample records are now those of the knots the sweep stands on
                    }
                    if constexpr (PERSIST)
                    {
                        __threadfence();
                        __syncwarp();
                        if (lane32 == 0)
                        {
                            *prm.counter = 0u;
                            *pa.t_release = global_timer_ns();
                            st_release_gpu(&pa.ctl->done, pa.target + 1u);
                        }
                    }
                    if (prm.host_out)
                    {
                        double2 *o = prm.host_out + 4 * gp.slot;
                        if (lane32 == 0)
                        {
                            publish_host(o + 0, cost0, prm.seq), publish_host(o + 1, cc, prm.seq);
                            publish_host(o + 2, (double)status0, prm.seq), publish_host(o + 3, model0, prm.seq);
                        }
                        if (gp.last)
                        {
                            double2 *ok = prm.host_out + 4 * MBAVO_MAX_LEVELS_DEV;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (lane32 + 32 * u < nk7)
                                    publish_host(ok + lane32 + 32 * u, keep[u], prm.seq);
                        }
                    }
                }
                __syncthreads();
            }
            if constexpr (!PERSIST) // (a sweep publishes scalars and knots, never the packed vector)
            {
                for (int e = threadIdx.x; e < E; e += blockDim.x)
                {
                    const double s = fin_s[e];
                    prm.packed_out[e] = s;
                    if (prm.host_out && !gp.state)
                        publish_host(prm.host_out + e, s, prm.seq);
                }
            }
            MBAVO_STAMP(10);
            if constexpr (PERSIST && WITH_J)
            {
                // everything this block wrote (records, sweep state, counter) becomes visible to the blocks that acquire `done`
                // (the cost pass released the sweep above, before its host stores)
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    *prm.counter = 0u;
                    *pa.t_release = global_timer_ns();
                    st_release_gpu(&pa.ctl->done, pa.target + 1u);
                }
            }
            else if constexpr (!PERSIST)
            {
                if (threadIdx.x == 0)
                    *prm.counter = 0u; // re-arm for the next launch
            }
            return true;
        }

        // DBG: also writes the per-stage intermediates of mbavo_debug_dump (patch centres, raw residuals, raw 1 x 6NK rows)
        template <int K, int NK, bool WITH_J, bool PACKED, int WARPS, bool PERSIST, bool DBG = false>
        __device__ __forceinline__ bool track_pass(const TrackParams &prm, unsigned char *smem_raw, const PersistArgs pa)
        {
            using G = RowGeom<NK, WITH_J>;
            constexpr int kWarpsPerBlock = WARPS;
            constexpr int REC = sample_rec_floats(K);
            constexpr int NJ = WITH_J ? NK : 1;
            constexpr int D1 = G::D1, PITCH = G::PITCH, NT = G::NT, MT = G::MT, E = G::E;
            constexpr int EPITCH = (E + 1) & ~1; // row pitch of the per-block partials (pass_finish reads them in 16-byte pairs)

            MBAVO_STAMP(0);
            if constexpr (!PERSIST)
            {
                if (prm.gn.state)
                    cudaTriggerProgrammaticLaunchCompletion(); // the next pose kernel of the sweep may queue up behind this grid
            }
            const LevelDev &lv = prm.lv;
            const int f = blockIdx.y;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int S = lv.S, N = lv.N, TP = prm.TP, PH = prm.PH;
            const int Q = 32 / PH;                           // pixel slots per pass
            // lane -> (slot, phase): slot-fastest (a quarter-warp = 8 pixels of one phase) or phase-fastest (a quarter-warp
            // = 8 consecutive exposure samples of one pixel)
            const bool pf = prm.phase_fast != 0;
            const int slot = pf ? lane / PH : lane & (Q - 1), phase = pf ? lane & (PH - 1) : lane / Q;
            const int bfly_lo = pf ? 1 : Q, bfly_hi = pf ? PH : 32;

            float *samples_s = reinterpret_cast<float *>(smem_raw);                  // N * REC
            int *seg_end_s = reinterpret_cast<int *>(samples_s + N * REC);          // kMaxSegments (+1 pad)
            double *mid_s = reinterpret_cast<double *>(seg_end_s + 8);              // kMidDoubles
            PixelRec *pix_s = reinterpret_cast<PixelRec *>(mid_s + kMidDoubles);    // warps * 32
            int2 *pattern_s = reinterpret_cast<int2 *>(pix_s + kWarpsPerBlock * 32); // S
            float *rho_s = reinterpret_cast<float *>(pattern_s + S);                // warps * max(32, TP*S)
            const int rho_per_warp = max(32, TP * S);
            unsigned short *tile_s = reinterpret_cast<unsigned short *>(rho_s + kWarpsPerBlock * rho_per_warp); // NT (padded)
            float *rows_s = reinterpret_cast<float *>(tile_s + ((NT + 7) & ~7));   // warps * 32 * PITCH
            // fp64 tile accumulators of every warp, [warp][m][ij][lane]: kept out of the register file (they are touched
            // once per 32-pixel chunk) so that the sample loop has the registers
            double *acc_s = reinterpret_cast<double *>(rows_s + kWarpsPerBlock * 32 * PITCH + ((kWarpsPerBlock * 32 * PITCH) & 1));
            double *red_s = reinterpret_cast<double *>(smem_raw);                   // epilogue: warps * E doubles (aliases all)

            for (int e = threadIdx.x; e < S; e += blockDim.x)
                pattern_s[e] = lv.pattern[e];
            if (WITH_J)
            {
                // tile table: tile id -> (ta, tb), ta <= tb, row-major upper triangle of the T x T tile grid
                for (int e = threadIdx.x; e < NT; e += blockDim.x)
                    tile_s[e] = (unsigned short)tile_s_copy(e, G::T);
            }
            // everything above is independent of the pose kernel; what follows reads its output (programmatic dependent
            // launch: this grid may have started before the pose kernel finished)
            MBAVO_STAMP(1);
            if constexpr (PERSIST)
            {
                if (!wait_pass(pa.ctl, pa.target))
                    return false;
                if (prm.ready_flag != nullptr && !wait_ready(prm.ready_flag, prm.ready_epoch, pa.ctl))
                    return false;
            }
            else
                cudaGridDependencySynchronize();
            MBAVO_STAMP(2);
            int buf = prm.buf_select == kBufB ? 1 : 0;
            if (prm.buf_select == kBufCur || prm.buf_select == kBufCand)
            {
                const int cb = __ldcg(&prm.gn.state->cur_buf);
                buf = prm.buf_select == kBufCur ? cb : 1 - cb;
            }
            const float *samples_g = prm.samples + (size_t)buf * prm.samples_stride;
            const int *seg_end_g = prm.seg_end + (size_t)buf * prm.seg_end_stride;
            const double *mid_g = prm.mid + (size_t)buf * prm.mid_stride;
            // (__ldcg: inside a persistent sweep the records were written by another SM during this very kernel)
            for (int e = threadIdx.x; e < N * REC; e += blockDim.x)
                samples_s[e] = __ldcg(samples_g + (size_t)f * N * REC + e);
            if (threadIdx.x < kMaxSegments)
                seg_end_s[threadIdx.x] = __ldcg(seg_end_g + f * kMaxSegments + threadIdx.x);
            if (threadIdx.x < kMidDoubles)
                mid_s[threadIdx.x] = __ldcg(mid_g + f * kMidDoubles + threadIdx.x);
            __syncthreads();

            MBAVO_STAMP(3);
            // Persistent sweep, cost pass: the last block of the grid takes no batches; it completes the candidate's sample
            // records with their Jacobian part (the Hessian pass's last block wrote the pose part only, to release this pass
            // sooner) — off the critical path, and finished before this block takes its ticket, i.e. before the next pass starts.
            const bool service_pass = PERSIST && !WITH_J && gridDim.x > 1 && prm.gn.state != nullptr;
            const bool service_block = service_pass && blockIdx.x == gridDim.x - 1;
            if (service_block && !prm.gn.last && prm.gn.chain)
            {
                pose_records_block<K>(pa.stage, prm.gn.state->cand_t, prm.gn.state->cand_R, kPoseJacobianOnly,
                                      const_cast<float *>(samples_g), nullptr, nullptr);
            }
            const double *mid = mid_s;
            const float2 fxy = f2((float)lv.fx, (float)lv.fy);
            const float inv_N = 1.0f / (float)N;
            const float huber_a = prm.huber_a;
            const double inv_num_residuals = prm.inv_num_residuals;
            float *my_rows = rows_s + warp * 32 * PITCH;
            float *my_rho = rho_s + warp * rho_per_warp;
            PixelRec *my_pix = pix_s + warp * 32;

            double *my_acc = acc_s + warp * (MT * 4 * 32) + lane;
            if constexpr (WITH_J)
            {
#pragma unroll
                for (int e = 0; e < MT * 4; ++e)
                    my_acc[e * 32] = 0.0;
            }
            double cost_acc = 0.0;

            const int items = TP * S;
            // batch -> warp: one launch per pass sizes its grid to the batches, block-major; the resident grid of a persistent
            // sweep is the whole GPU whatever the level, so batches go block-fastest (a coarse level still reaches every SM)
            const int nblk = (int)gridDim.x - (service_pass ? 1 : 0); // blocks that take batches
            const int wb_first = service_block ? prm.batches_per_frame
                                               : (PERSIST ? warp * nblk + (int)blockIdx.x : (int)blockIdx.x * kWarpsPerBlock + warp);
            for (int wb = wb_first; wb < prm.batches_per_frame; wb += nblk * kWarpsPerBlock)
            {
                const int p0 = wb * TP;
                for (int base = 0; base < items; base += 32)
                {
                    // ---- A: pixel records -----------------------------------------------------------------------------
                    {
                        const int it = base + lane;
                        const bool in_batch = it < items;
                        double2 *centre_out = nullptr;
                        if constexpr (DBG)
                        {
                            if (in_batch && it % S == 0 && p0 + it / S < lv.P && prm.dbg_centres)
                                centre_out = prm.dbg_centres + (size_t)f * lv.P + p0 + it / S;
                        }
                        my_pix[lane] = setup_pixel(lv, mid, f, in_batch ? p0 + it / S : lv.P, in_batch ? it % S : 0, pattern_s, centre_out);
                    }
                    __syncwarp();
                    MBAVO_STAMP(4);
                    // ---- B: exposure samples, PH phases per pixel ---------------------------------------------------
                    const int chunk = min(32, items - base);
                    for (int sub = 0; sub < chunk; sub += Q)
                    {
                        const int q = sub + slot;
                        const float4 r0 = *reinterpret_cast<const float4 *>(my_pix + q);
                        const int4 r1 = *(reinterpret_cast<const int4 *>(my_pix + q) + 1);
                        const bool valid = r1.w & 1;
                        PixelRegs ps;
                        ps.rxy = f2(r0.x, r0.y), ps.D = r0.z, ps.iD = r0.w;
                        ps.X = r1.y, ps.Y = r1.z;
                        ps.lox = -(float)ps.X, ps.hix = (float)(lv.W - 1 - ps.X);
                        ps.loy = -(float)ps.Y, ps.hiy = (float)(lv.H - 1 - ps.Y);
                        ps.fxyiD = mul2(fxy, bc((PACKED && MBAVO_TEXEL == 3) ? 0.5f * ps.iD : ps.iD)); // (patch texels: doubled gradients)

                        float sumI = 0.f;
                        float2 J[NJ][3];
#pragma unroll
                        for (int a = 0; a < NJ; ++a)
                            J[a][0] = J[a][1] = J[a][2] = f2(0.f, 0.f);

                        if (valid)
                        {
                            if constexpr (WITH_J)
                            {
                                int i = phase;
                                SegmentLoop<K, NK, PACKED, 0>::run(samples_s, seg_end_s, i, PH, ps, lv, fxy, sumI, J);
                            }
                            else
                            {
#pragma unroll kCostUnroll
                                for (int i = phase; i < N; i += PH) // cost only: the segment of a sample is irrelevant
                                    sample_step<K, NK, false, PACKED, 0>(SmemRec{samples_s + i * REC}, ps, lv, fxy, sumI, J);
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
                                        J[a][c].x += __shfl_xor_sync(0xffffffffu, J[a][c].x, o);
                                        J[a][c].y += __shfl_xor_sync(0xffffffffu, J[a][c].y, o);
                                    }
                                }
                            }
                        }
                        // residual (…cost.cu:115-121), Huber, weighted row (…cost.cu:185-206)
                        const float icur = __int_as_float(r1.x);
                        const float r = valid ? sumI * inv_N - icur : 0.f;
                        float sw;
                        const float rho = huber(r, huber_a, sw);
                        if constexpr (DBG && WITH_J)
                        {
                            // the pixel's raw residual and Jacobian row (compute_hessian_gradients_cost.cu:115-151), before Huber
                            const int pp = p0 + (base + q) / S;
                            if (phase == 0 && q < chunk && pp < lv.P)
                            {
                                const size_t pix = ((size_t)f * lv.P + pp) * S + (base + q) % S;
                                if (prm.dbg_r)
                                    prm.dbg_r[pix] = r;
                                if (prm.dbg_J)
                                {
                                    float *o = prm.dbg_J + pix * (6 * NK);
#pragma unroll
                                    for (int a = 0; a < NK; ++a)
                                    {
                                        o[3 * a + 0] = valid ? inv_N * J[a][0].x : 0.f, o[3 * a + 1] = valid ? inv_N * J[a][0].y : 0.f;
                                        o[3 * a + 2] = valid ? inv_N * J[a][1].x : 0.f;
                                        o[3 * NK + 3 * a + 0] = valid ? inv_N * J[a][1].y : 0.f;
                                        o[3 * NK + 3 * a + 1] = valid ? inv_N * J[a][2].x : 0.f, o[3 * NK + 3 * a + 2] = valid ? inv_N * J[a][2].y : 0.f;
                                    }
                                }
                            }
                        }
                        if (phase == 0 && q < chunk)
                        {
                            my_rho[base + q] = rho;
                            if constexpr (WITH_J)
                            {
                                const float scale = (r1.w == 1) ? sw : 0.f; // invalid pixels and outliers add nothing (…cost.cu:267)
                                float *row = my_rows + q * PITCH;
                                row[0] = scale * r;
                                const float sj = scale * inv_N;
                                // row = [r | t-block of every knot | w-block of every knot] (merge_hessian_gradient_cost.cpp:52-62)
#pragma unroll
                                for (int a = 0; a < NK; ++a)
                                {
                                    row[1 + 3 * a + 0] = sj * J[a][0].x;
                                    row[1 + 3 * a + 1] = sj * J[a][0].y;
                                    row[1 + 3 * a + 2] = sj * J[a][1].x;
                                    row[1 + 3 * NK + 3 * a + 0] = sj * J[a][1].y;
                                    row[1 + 3 * NK + 3 * a + 1] = sj * J[a][2].x;
                                    row[1 + 3 * NK + 3 * a + 2] = sj * J[a][2].y;
                                }
                                row[D1] = 0.f; // padding column of the 2x2 tiling
                            }
                        }
                    }
                    MBAVO_STAMP(5);
                    if constexpr (WITH_J)
                    {
                        // rows of a short last chunk: zero
                        if (lane >= chunk)
                        {
                            float *row = my_rows + lane * PITCH;
#pragma unroll
                            for (int c = 0; c <= D1; ++c)
                                row[c] = 0.f;
                        }
                        __syncwarp();
                        // ---- C: rows^T rows (…cost.cu:214-230) in 2x2 tiles: (a0 a1)^T (b0 b1) summed over the 32 rows
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                        {
                            const int id = lane + 32 * m;
                            if (id < NT)
                            {
                                const unsigned int t = tile_s[id];
                                const float *pa = my_rows + 2 * (t >> 8), *pb = my_rows + 2 * (t & 0xff);
                                float2 s0 = f2(0.f, 0.f), s1 = f2(0.f, 0.f);
#pragma unroll 8
                                for (int qq = 0; qq < 32; ++qq)
                                {
                                    const float2 va = *reinterpret_cast<const float2 *>(pa + qq * PITCH);
                                    const float2 vb = *reinterpret_cast<const float2 *>(pb + qq * PITCH);
                                    s0 = fma2(bc(va.x), vb, s0);
                                    s1 = fma2(bc(va.y), vb, s1);
                                }
                                double *a = my_acc + m * 4 * 32;
                                a[0] += (double)s0.x, a[32] += (double)s0.y, a[64] += (double)s1.x, a[96] += (double)s1.y;
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

            MBAVO_STAMP(6);
            // ---- epilogue: warp -> block -> grid, all in fixed order (deterministic) --------------------------------
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);

            // the accumulators leave shared memory before red_s (which aliases everything) is written
            double acc[MT][4];
            if constexpr (WITH_J)
            {
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int ij = 0; ij < 4; ++ij)
                        acc[m][ij] = my_acc[(m * 4 + ij) * 32];
            }
            __syncthreads();
            if constexpr (WITH_J)
            {
                // tile element (i, j) is matrix element (r, c) = (2 ta + i, 2 tb + j); packed index of the upper triangle
#pragma unroll
                for (int m = 0; m < MT; ++m)
                {
                    const int id = lane + 32 * m;
                    if (id < NT)
                    {
                        const unsigned int t = tile_s_copy(id, G::T);
#pragma unroll
                        for (int ij = 0; ij < 4; ++ij)
                        {
                            const int r = 2 * (int)(t >> 8) + (ij >> 1), c = 2 * (int)(t & 0xff) + (ij & 1);
                            if (r <= c && c < D1)
                                red_s[warp * E + r * D1 - r * (r - 1) / 2 + (c - r)] = acc[m][ij];
                        }
                    }
                }
                __syncwarp();
            }
            if (lane == 0)
                red_s[warp * E] = cost_acc; // packed[0] is the cost, not sum r^2 (…cost.cu:232-238)
            __syncthreads();
            const int block_linear = blockIdx.y * gridDim.x + blockIdx.x;
            const int num_blocks = gridDim.x * gridDim.y;
            for (int e = threadIdx.x; e < E; e += blockDim.x)
            {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kWarpsPerBlock; ++w)
                    s += red_s[w * E + e];
                prm.block_partials[(size_t)block_linear * EPITCH + e] = s;
            }
            if (EPITCH != E && threadIdx.x == 0)
                prm.block_partials[(size_t)block_linear * EPITCH + E] = 0.0; // padding element of the row (read as half of a pair)
            MBAVO_STAMP(7);
            __threadfence();
            __shared__ unsigned int ticket_s;
            __syncthreads();
            if (threadIdx.x == 0)
                ticket_s = atomicAdd(prm.counter, 1u);
            __syncthreads();
            if (ticket_s != (unsigned int)(num_blocks - 1))
                return true;
            return pass_finish<K, NK, WITH_J, WARPS, PERSIST>(prm, smem_raw, pa);
        }

        template <int K, int NK, bool WITH_J, bool PACKED, bool BIG>
        __global__ void __launch_bounds__(track_warps(WITH_J, NK, BIG) * 32, WITH_J ? ((!BIG && NK <= 3) ? 2 : 1) : MBAVO_MINB_C)
            track_kernel(const __grid_constant__ TrackParams prm)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            track_pass<K, NK, WITH_J, PACKED, track_warps(WITH_J, NK, BIG), false>(prm, smem_raw, PersistArgs{});
        }

        // Hessian pass that also dumps its per-stage intermediates (mbavo_debug_dump): test infrastructure of the product kernel,
        // same code path with DBG = true; small block shape.
        template <int K, int NK, bool PACKED>
        __global__ void __launch_bounds__(track_warps(true, NK, false) * 32, 1) track_debug_kernel(const __grid_constant__ TrackParams prm)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            track_pass<K, NK, true, PACKED, track_warps(true, NK, false), false, true>(prm, smem_raw, PersistArgs{});
        }
        template <int K, int NK, bool PACKED>
        cudaError_t launch_debug_one(const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream)
        {
            cudaError_t e = cudaFuncSetAttribute(track_debug_kernel<K, NK, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess)
                return e;
            track_debug_kernel<K, NK, PACKED><<<grid, track_warps(true, NK, false) * 32, smem, stream>>>(prm);
            return cudaGetLastError();
        }

        // The whole coarse-to-fine sweep in one launch: one block per SM, all passes inside (see SweepCtl).  The sample records
        // of the knots the sweep starts from come from the pose kernel launched right before it.
        template <int K, int NK, bool PACKED>
        __global__ void __launch_bounds__(track_warps(true, NK, true) * 32, 1)
            sweep_kernel(const __grid_constant__ SweepParams sp, const __grid_constant__ EvalStage stage)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            constexpr int WARPS = track_warps(true, NK, true);
            // the frame times / segment table in shared memory: the pose records of a candidate are computed by an out-of-line
            // function that reads them through a pointer (generic loads from the parameter space are slow and it is latency that counts)
            __shared__ EvalStage stage_s;
            {
                const int *src = reinterpret_cast<const int *>(&stage);
                int *dst = reinterpret_cast<int *>(&stage_s);
                for (int e = threadIdx.x; e < (int)(sizeof(EvalStage) / sizeof(int)); e += blockDim.x)
                    dst[e] = src[e];
            }
            __syncthreads();
            // MBAVO_FAST_FINISH: position of every exposure sample inside its spline segment — fixed for the whole sweep, so the two
            // divisions per sample leave the serial section between a Hessian pass and its cost pass — and room for the knots
            __shared__ double sample_u_s[64];
            __shared__ double knots_s[2 * 112];
            const bool have_u = MBAVO_FAST_FINISH && stage_s.N * stage_s.F <= 64;
            if (have_u && (int)threadIdx.x < stage_s.N * stage_s.F)
                sample_u_s[threadIdx.x] = sample_u(&stage_s, threadIdx.x);
            __syncthreads();
            cudaGridDependencySynchronize(); // the pose kernel's records (and the sweep state it initialised)
            if (blockIdx.x == 0 && threadIdx.x == 0)
                sp.pass_times[0] = global_timer_ns();
            for (int li = 0; li < sp.n_levels; ++li)
            {
                const PersistArgs ph{sp.ctl, sp.base + 2u * (unsigned int)li, &stage_s, sp.pass_times + 1 + 2 * li, have_u ? sample_u_s : nullptr, knots_s};
                if (!track_pass<K, NK, true, PACKED, WARPS, true>(sp.pass[2 * li], smem_raw, ph))
                    return;
                const PersistArgs pc{sp.ctl, sp.base + 2u * (unsigned int)li + 1u, &stage_s, sp.pass_times + 2 + 2 * li, have_u ? sample_u_s : nullptr, knots_s};
                if (!track_pass<K, NK, false, PACKED, WARPS, true>(sp.pass[2 * li + 1], smem_raw, pc))
                    return;
            }
        }

        template <int K, int NK, bool WITH_J, bool PACKED, bool BIG>
        cudaError_t launch_one(const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream, bool dependent)
        {
            static unsigned long long configured = 0; // per instantiation and per device (attribute of the device function)
            int dev = 0;
            cudaGetDevice(&dev);
            if (!(configured >> dev & 1ull))
            {
                cudaError_t e = cudaFuncSetAttribute(track_kernel<K, NK, WITH_J, PACKED, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     200 * 1024);
                if (e != cudaSuccess)
                    return e;
                configured |= 1ull << dev;
            }
            // Programmatic dependent launch: the kernel may start while the pose kernel before it in the stream is still
            // running; it waits (griddepcontrol.wait) right before it reads what the pose kernel wrote.
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = grid, cfg.blockDim = dim3(track_warps(WITH_J, NK, BIG) * 32, 1, 1), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr, cfg.numAttrs = dependent ? 1 : 0;
            return cudaLaunchKernelEx(&cfg, track_kernel<K, NK, WITH_J, PACKED, BIG>, prm);
        }

        // sweep_kernel: one block per SM, cooperative (all blocks co-resident: they wait for each other between the passes),
        // programmatically chained behind the pose kernel when the driver accepts both attributes together
        template <int K, int NK, bool PACKED>
        cudaError_t launch_sweep_one(const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream, bool dependent,
                                     int *query_occupancy)
        {
            static unsigned long long configured = 0;
            static int pdl_ok = 1; // cooperative + programmatic serialisation accepted together
            int dev = 0;
            cudaGetDevice(&dev);
            if (!(configured >> dev & 1ull))
            {
                cudaError_t e = cudaFuncSetAttribute(sweep_kernel<K, NK, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e != cudaSuccess)
                    return e;
                configured |= 1ull << dev;
            }
            constexpr int threads = track_warps(true, NK, true) * 32;
            if (query_occupancy)
            {
                int n = 0;
                cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sweep_kernel<K, NK, PACKED>, threads, smem);
                *query_occupancy = n;
                return e;
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(num_sms, 1, 1), cfg.blockDim = dim3(threads, 1, 1), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            if (dependent && pdl_ok)
            {
                cfg.numAttrs = 2;
                cudaError_t e = cudaLaunchKernelEx(&cfg, sweep_kernel<K, NK, PACKED>, sp, stage);
                if (e == cudaSuccess)
                    return e;
                cudaGetLastError();
                pdl_ok = 0;
            }
            cfg.numAttrs = 1;
            return cudaLaunchKernelEx(&cfg, sweep_kernel<K, NK, PACKED>, sp, stage);
        }

        // Occupancy-derived grid width for one instantiation
        template <int K, int NK, bool WITH_J, bool PACKED, bool BIG>
        int blocks_per_sm(size_t smem)
        {
            int n = 0;
            cudaFuncSetAttribute(track_kernel<K, NK, WITH_J, PACKED, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, track_kernel<K, NK, WITH_J, PACKED, BIG>, track_warps(WITH_J, NK, BIG) * 32, smem);
            return n > 0 ? n : 1;
        }

        // One (K, NK, WITH_J): texel / direct-gather variant x block shape (the cost-only pass has one shape)
        template <int K, int NK, bool WITH_J>
        cudaError_t dispatch_variant(bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                     int *query_occupancy, bool dependent)
        {
#define MBAVO_VARIANT(P_, B_)                                                             \
    if (packed == P_ && big == B_)                                                        \
    {                                                                                     \
        if (query_occupancy)                                                              \
        {                                                                                 \
            *query_occupancy = blocks_per_sm<K, NK, WITH_J, P_, B_>(smem);                \
            return cudaSuccess;                                                           \
        }                                                                                 \
        return launch_one<K, NK, WITH_J, P_, B_>(prm, grid, smem, stream, dependent);     \
    }
            if constexpr (WITH_J)
            {
                MBAVO_VARIANT(true, true)
                MBAVO_VARIANT(false, true)
            }
            big = false;
            MBAVO_VARIANT(true, false)
            MBAVO_VARIANT(false, false)
#undef MBAVO_VARIANT
            return cudaErrorInvalidValue;
        }
    } // namespace
} // namespace mbavo
#endif
