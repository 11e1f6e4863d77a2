// Fused blur-aware photometric tracking kernel for one pyramid level.
//
// Replaces, in ONE launch, the reference's kernel chain (src/ba_tracker):
//   kernel_compute_local_patches_xy              compute_local_patches_xy.cu:9-50
//   kernel_compute_pixel_jacobian_residual       compute_hessian_gradients_cost.cu:23-156   (+ compute_pixel_intensity.h:25-209)
//   kernel_compute_patch_cost_gradient_hessian   compute_hessian_gradients_cost.cu:165-239
//   kernel_compute_frame_cost_gradient_hessian   compute_hessian_gradients_cost.cu:247-283
// without the P*S*N*d fp64 scratch, the (d+1) block barriers per pixel and the E barriers per patch.
//
// Mapping.  A warp owns a batch of TP whole host-map points of one frame (TP*S = 32 lanes for the 8-pixel pattern);
// a lane owns one residual pixel and walks the N exposure samples sequentially in registers (no cross-thread
// reduction for the exposure average).  Per sample: plane-induced warp (fp32), 4-tap bilinear gather of I and of the
// interleaved gradient, dI/dt (1x3) and dI/dtheta (1x3, right perturbation of the pose rotation), chained to the
// control knots through the per-sample spline blocks produced by pose_kernel.  After the loop: residual, Huber,
// row = sqrt(w) [r, J]; the 32 rows of the warp are staged in shared memory and the packed upper triangle of
// row^T row is accumulated with one element set per lane (fp32 over the 32 rows, fp64 across batches).
// Epilogue: deterministic block reduction -> per-block partials -> last block sums all partials in block order.
//
// Precision: per-sample arithmetic fp32 (the reference's bilinear taps/weights are fp32 too, compute_pixel_intensity.h:43-68),
// patch centre fp64 (its truncation picks the live pixel, …cost.cu:69-70), every sum across pixels fp64.
#include "mbavo_device.h"

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

        struct PixelState
        {
            float rx, ry;     // ray of the live pixel, (X - cx)/fx, (Y - cy)/fy (z = 1; the result is scale-invariant)
            float D;          // plane depth of the point
            float iD;         // 1 / (D + 1e-8): projection with P_z == D        compute_pixel_intensity.h:137
            int X, Y;         // live pixel (also the origin of the reference coordinate, see sample_step)
            float icur;       // live-image intensity at the pixel
            bool valid;       // pixel inside the live image and point slot in range
        };

        // Per-item setup: patch centre (fp64), integer live pixel, ray.  compute_local_patches_xy.cu:26-49,
        // compute_hessian_gradients_cost.cu:63-78.
        __device__ __forceinline__ PixelState setup_pixel(const LevelDev &lv, const double *__restrict__ mid, int f, int p,
                                                          int j, const int2 *__restrict__ pattern_s)
        {
            PixelState ps;
            ps.valid = false;
            ps.rx = ps.ry = 0.f;
            ps.D = 1.f;
            ps.iD = 1.f;
            ps.X = ps.Y = 0;
            ps.icur = 0.f;
            if (p >= lv.P)
                return ps;
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
            ps.valid = true;
            ps.X = X, ps.Y = Y;
            ps.rx = (float)(((double)X - lv.cx) / lv.fx);
            ps.ry = (float)(((double)Y - lv.cy) / lv.fy);
            ps.D = (float)z;
            ps.iD = (float)(1.0 / (z + 1e-8));
            ps.icur = u8_to_float(__ldg(lv.cur_I[f] + (size_t)Y * lv.W + X));
            return ps;
        }

        // One exposure sample of one pixel.  K knots per segment; OFF = segment offset inside the knot window.
        // Accumulates sumI and, if WITH_J, the 1 x 6NK row (Jt: translation block, Jw: rotation block).
        //
        // Reference coordinate.  compute_pixel_intensity.h:108-144 evaluates u = fx P_x / (D + 1e-8) + c_x in fp64 and
        // rounds only the fractional part to fp32.  A plain fp32 evaluation of u loses ~W * 2^-24 px (4e-5 px at VGA),
        // which is what limits the parity of H, g and the step.  Here the coordinate is evaluated RELATIVE TO THE LIVE
        // PIXEL X:  with A = (R - I) ray (small), m = ray + A, tau = t_z / (D + 1e-8),
        //     u - X = fx [ (A_x - r_x A_z - tau m_x) / m_z + t_x / (D + 1e-8) ]
        // in which every term is of the size of the blur, so fp32 keeps ~1e-6 px; the integer tap is X + floor(u - X).
        template <int K, int NK, bool WITH_J, int OFF>
        __device__ __forceinline__ void sample_step(const float *__restrict__ rec, const PixelState &ps, const LevelDev &lv,
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
            const float il = __frcp_rn(m2);                     // 1 / lambda
            const float tau = a2.w * ps.iD;
            const float du = fxf * fmaf(fmaf(-tau, m0, fmaf(-ps.rx, A2, A0)), il, a2.y * ps.iD);
            const float dv = fyf * fmaf(fmaf(-tau, m1, fmaf(-ps.ry, A2, A1)), il, a2.z * ps.iD);
            // inside [0, W-1] x [0, H-1] (compute_pixel_intensity.h:35); an invalid sample contributes nothing while the
            // divisor stays N (…cost.cu:107-110)
            const float xf = floorf(du), yf = floorf(dv);
            const float dx = du - xf, dy = dv - yf; // exact
            const int xi = ps.X + (int)xf, yi = ps.Y + (int)yf;
            if (!(xi >= 0 && yi >= 0 && (xi < lv.W - 1 || (xi == lv.W - 1 && dx == 0.f)) &&
                  (yi < lv.H - 1 || (yi == lv.H - 1 && dy == 0.f)) && du == du && dv == dv))
                return;

            // bilinear_interpolation, compute_pixel_intensity.h:40-68
            const float dxdy = dx * dy;
            const float w00 = 1.0f - dx - dy + dxdy, w01 = dx - dxdy, w10 = dy - dxdy, w11 = dxdy;
            // the +1 taps carry weight 0 on the last column / row; clamp them instead of reading out of bounds
            const int x1 = min(xi + 1, lv.W - 1), y1 = min(yi + 1, lv.H - 1);
            const int i00 = yi * lv.W + xi, i01 = yi * lv.W + x1, i10 = y1 * lv.W + xi, i11 = y1 * lv.W + x1;
            const float I00 = u8_to_float(__ldg(lv.ref_I + i00)), I01 = u8_to_float(__ldg(lv.ref_I + i01));
            const float I10 = u8_to_float(__ldg(lv.ref_I + i10)), I11 = u8_to_float(__ldg(lv.ref_I + i11));
            sumI += w11 * I11 + w10 * I10 + w01 * I01 + w00 * I00;

            if (WITH_J)
            {
                const float2 g00 = __ldg(lv.ref_dIxy + i00), g01 = __ldg(lv.ref_dIxy + i01);
                const float2 g10 = __ldg(lv.ref_dIxy + i10), g11 = __ldg(lv.ref_dIxy + i11);
                const float gx = w11 * g11.x + w10 * g10.x + w01 * g01.x + w00 * g00.x;
                const float gy = w11 * g11.y + w10 * g10.y + w01 * g01.y + w00 * g00.y;
                const float s = (ps.D - a2.w) * il;                 // (D - t_z) / lambda       compute_pixel_intensity.h:128
                // dI/dt = dI/dP (I - m e_z^T / lambda)                                   compute_pixel_intensity.h:196-202
                const float gtx = gx * (fxf * ps.iD), gty = gy * (fyf * ps.iD);
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

        template <int K, int NK, bool WITH_J, int OFF>
        struct SegmentLoop
        {
            __device__ __forceinline__ static void run(const float *__restrict__ samples_s, const int *__restrict__ seg_end_s,
                                                       int &i, const PixelState &ps, const LevelDev &lv, float fxf, float fyf,
                                                       float &sumI, float (&Jt)[WITH_J ? NK : 1][3],
                                                       float (&Jw)[WITH_J ? NK : 1][3])
            {
                constexpr int REC = sample_rec_floats(K);
                const int end = seg_end_s[OFF];
                for (; i < end; ++i)
                    sample_step<K, NK, WITH_J, OFF>(samples_s + i * REC, ps, lv, fxf, fyf, sumI, Jt, Jw);
                if constexpr (OFF + 1 <= NK - K)
                    SegmentLoop<K, NK, WITH_J, (OFF + 1 <= NK - K ? OFF + 1 : OFF)>::run(samples_s, seg_end_s, i, ps, lv, fxf,
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

        // K: knots per segment, NK: knots in the window (NK - K + 1 segments touched), WITH_J: Hessian pass or cost only.
        template <int K, int NK, bool WITH_J>
        __global__ void __launch_bounds__(kThreads) track_kernel(const TrackParams prm)
        {
            constexpr int REC = sample_rec_floats(K);
            constexpr int D1 = WITH_J ? 6 * NK + 1 : 1;     // row length: [r | J]
            constexpr int D1P = D1 | 1;                      // odd row pitch: conflict-free row writes
            constexpr int E = WITH_J ? packed_len(NK) : 1;   // packed upper triangle
            constexpr int ME = (E + 31) / 32;                // elements owned by one lane

            const LevelDev &lv = prm.lv;
            const int f = blockIdx.y;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int S = lv.S, N = lv.N, TP = prm.TP;

            extern __shared__ __align__(16) unsigned char smem_raw[];
            float *samples_s = reinterpret_cast<float *>(smem_raw);                  // N * REC
            int *seg_end_s = reinterpret_cast<int *>(samples_s + N * REC);          // kMaxSegments (+1 pad)
            int2 *pattern_s = reinterpret_cast<int2 *>(seg_end_s + 8);              // S
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
            const float huber_a = prm.stage->huber_a;
            const double inv_num_residuals = prm.stage->inv_num_residuals;
            float *my_rows = rows_s + warp * 32 * D1P;
            float *my_rho = rho_s + warp * rho_per_warp;

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
                    const int it = base + lane;
                    const bool in_batch = it < items;
                    const int p = p0 + (in_batch ? it / S : 0), j = in_batch ? it % S : 0;
                    PixelState ps = setup_pixel(lv, mid, f, in_batch ? p : lv.P, j, pattern_s);

                    float sumI = 0.f;
                    float Jt[WITH_J ? NK : 1][3], Jw[WITH_J ? NK : 1][3];
#pragma unroll
                    for (int a = 0; a < (WITH_J ? NK : 1); ++a)
                        Jt[a][0] = Jt[a][1] = Jt[a][2] = Jw[a][0] = Jw[a][1] = Jw[a][2] = 0.f;

                    if (ps.valid)
                    {
                        if constexpr (WITH_J)
                        {
                            int i = 0;
                            SegmentLoop<K, NK, true, 0>::run(samples_s, seg_end_s, i, ps, lv, fxf, fyf, sumI, Jt, Jw);
                        }
                        else
                        {
                            for (int i = 0; i < N; ++i) // cost only: the segment of a sample is irrelevant
                                sample_step<K, NK, false, 0>(samples_s + i * REC, ps, lv, fxf, fyf, sumI, Jt, Jw);
                        }
                    }

                    // residual (…cost.cu:115-121), Huber, weighted row (…cost.cu:185-206)
                    const float r = ps.valid ? sumI * inv_N - ps.icur : 0.f;
                    float sw;
                    const float rho = huber(r, huber_a, sw);
                    if (in_batch)
                        my_rho[it] = rho;
                    if (WITH_J)
                    {
                        const bool flagged = (p < lv.P) && lv.flags[p] == 1;
                        const float scale = (ps.valid && !flagged) ? sw : 0.f; // outliers are skipped in the sums (…cost.cu:267)
                        float *row = my_rows + lane * D1P;
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
                        __syncwarp();
                        // packed upper triangle of rows^T rows (…cost.cu:214-230): lane owns elements lane, lane+32, …
#pragma unroll
                        for (int m = 0; m < ME; ++m)
                        {
                            const int e = lane + 32 * m;
                            if (e < E)
                            {
                                const int a = pair_s[e] >> 8, b = pair_s[e] & 0xff;
                                float sacc = 0.f;
#pragma unroll 8
                                for (int q = 0; q < 32; ++q)
                                    sacc = fmaf(my_rows[q * D1P + a], my_rows[q * D1P + b], sacc);
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
                prm.packed_out[e] = s * inv_num_residuals;
            }
            if (threadIdx.x == 0)
                *prm.counter = 0u; // re-arm for the next launch
        }

        template <int K, int NK, bool WITH_J>
        cudaError_t launch_one(const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream)
        {
            static unsigned long long configured = 0; // per instantiation and per device (attribute of the device function)
            int dev = 0;
            cudaGetDevice(&dev);
            if (!(configured >> dev & 1ull))
            {
                cudaError_t e = cudaFuncSetAttribute(track_kernel<K, NK, WITH_J>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     200 * 1024);
                if (e != cudaSuccess)
                    return e;
                configured |= 1ull << dev;
            }
            track_kernel<K, NK, WITH_J><<<grid, kThreads, smem, stream>>>(prm);
            return cudaGetLastError();
        }
    } // namespace

    size_t track_kernel_smem_bytes(int K, int NK, bool with_j, int N, int S, int TP)
    {
        const int REC = sample_rec_floats(K);
        const int D1 = with_j ? 6 * NK + 1 : 1, D1P = D1 | 1;
        const int E = with_j ? packed_len(NK) : 1;
        const int rho_per_warp = max(32, TP * S);
        size_t main_bytes = (size_t)N * REC * 4 + 8 * 4 + (size_t)S * 8 + (size_t)kWarpsPerBlock * rho_per_warp * 4 +
                            (size_t)((E + 7) & ~7) * 2 + (size_t)kWarpsPerBlock * 32 * D1P * 4;
        size_t red_bytes = (size_t)kWarpsPerBlock * E * 8;
        return (main_bytes > red_bytes ? main_bytes : red_bytes) + 16;
    }

    // Occupancy-derived grid width for one instantiation
    template <int K, int NK, bool WITH_J>
    static int blocks_per_sm(size_t smem)
    {
        int n = 0;
        cudaFuncSetAttribute(track_kernel<K, NK, WITH_J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, track_kernel<K, NK, WITH_J>, kThreads, smem);
        return n > 0 ? n : 1;
    }

#define MBAVO_DISPATCH(K_, NK_)                                                                     \
    if (K == K_ && NK == NK_)                                                                       \
    {                                                                                               \
        if (query_occupancy)                                                                        \
        {                                                                                           \
            *query_occupancy = blocks_per_sm<K_, NK_, true>(smem);                                  \
            return cudaSuccess;                                                                     \
        }                                                                                           \
        return launch_one<K_, NK_, true>(prm, grid, smem, stream);                                  \
    }

    // with_j: Hessian pass (templated on the window) or cost-only pass (one instantiation per K)
    cudaError_t launch_track_kernel(int K, int NK, bool with_j, const TrackParams &prm, dim3 grid, size_t smem,
                                    cudaStream_t stream, int *query_occupancy)
    {
        if (!with_j)
        {
            if (K == 2)
            {
                if (query_occupancy)
                {
                    *query_occupancy = blocks_per_sm<2, 2, false>(smem);
                    return cudaSuccess;
                }
                return launch_one<2, 2, false>(prm, grid, smem, stream);
            }
            if (K == 4)
            {
                if (query_occupancy)
                {
                    *query_occupancy = blocks_per_sm<4, 4, false>(smem);
                    return cudaSuccess;
                }
                return launch_one<4, 4, false>(prm, grid, smem, stream);
            }
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
