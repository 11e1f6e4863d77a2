// EXPERIMENT (measured, see profiles/r2_tma_variant.md): the cost-only pass with TMA-staged keyframe tiles.
//
// The north star asks for "one fused kernel per pyramid level that TMA-stages image tiles into shared memory".  The product
// kernels gather the keyframe through L1 from packed texels instead (track_kernel.cuh); this unit is the staged alternative, built
// to be MEASURED against them on the same inputs, with the same arithmetic and the same result:
//   * a warp owns batches of 4 host-map points (lane = pixel), as in track_pass;
//   * per point one 2-D tile of the 8-bit keyframe image, box_w x box_h bytes centred on the point's keyframe pixel, is fetched
//     by cp.async.bulk.tensor.2d (one elected lane per point issues it; completion on a per-warp mbarrier), double-buffered:
//     the tiles of the warp's next batch are in flight while it walks the exposure samples of the current one;
//   * a sample whose 2 x 2 bilinear footprint lies inside the tile takes its four taps from shared memory, any other sample
//     falls back to the global quad-texel gather of the product kernel — so the result does not depend on the box size.
// Reference: bilinear_interpolation / compute_pixel_intensity (src/ba_tracker/compute_pixel_intensity.h:25-72, 91-144),
// kernel_compute_pixel_jacobian_residual cost-only branch (compute_hessian_gradients_cost.cu:23-121), Huber (…:188-199).
#include "track_kernel.cuh"

#include <cuda.h>

namespace mbavo
{
    namespace
    {
        constexpr int kTmaWarps = 8;

        __device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
        __device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
        }
        __device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        }
        __device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
        {
            asm volatile("{\n"
                         ".reg .pred p;\n"
                         "WAIT_%=:\n"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                         "@p bra DONE_%=;\n"
                         "bra WAIT_%=;\n"
                         "DONE_%=:\n"
                         "}" ::"r"(smem_u32(bar)),
                         "r"(parity)
                         : "memory");
        }
        __device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar)
        {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                         "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                         : "memory");
        }

        // Cost-only pass, K knots per segment (only the pose part of the sample records is read).
        template <int K>
        __global__ void __launch_bounds__(kTmaWarps * 32, 1)
            track_cost_tma_kernel(const __grid_constant__ TrackParams prm, const __grid_constant__ CUtensorMap tmap, int box_w, int box_h,
                                  double *__restrict__ cost_out, unsigned long long *__restrict__ counters)
        {
            constexpr int REC = sample_rec_floats(K);
            const LevelDev &lv = prm.lv;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int S = lv.S, N = lv.N, TP = prm.TP;
            const int tile_bytes = box_w * box_h; // per point; a multiple of 128 (host checks)

            extern __shared__ __align__(16) unsigned char smem_raw[];
            // TMA destinations must be 128-byte aligned (the launch reserves the slack)
            unsigned char *tiles_s = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // [warp][buf][TP][tile_bytes]
            float *samples_s = reinterpret_cast<float *>(tiles_s + (size_t)kTmaWarps * 2 * TP * tile_bytes); // N * REC
            double *mid_s = reinterpret_cast<double *>(samples_s + ((N * REC + 3) & ~3));
            PixelRec *pix_s = reinterpret_cast<PixelRec *>(mid_s + kMidDoubles);                 // warps * 32
            int2 *pattern_s = reinterpret_cast<int2 *>(pix_s + kTmaWarps * 32);
            int2 *origin_s = pattern_s + ((S + 1) & ~1);                                         // [warp][buf][TP]
            unsigned long long *bars = reinterpret_cast<unsigned long long *>(origin_s + kTmaWarps * 2 * TP); // [warp][buf]
            float *rho_s = reinterpret_cast<float *>(bars + kTmaWarps * 2);                       // [warp][32]
            __shared__ double red_s[kTmaWarps];

            for (int e = threadIdx.x; e < S; e += blockDim.x)
                pattern_s[e] = lv.pattern[e];
            for (int e = threadIdx.x; e < N * REC; e += blockDim.x)
                samples_s[e] = __ldcg(prm.samples + e);
            if (threadIdx.x < kMidDoubles)
                mid_s[threadIdx.x] = __ldcg(prm.mid + threadIdx.x);
            if (lane == 0)
            {
                mbar_init(bars + warp * 2, 1);
                mbar_init(bars + warp * 2 + 1, 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            __syncthreads();

            const float2 fxy = f2((float)lv.fx, (float)lv.fy);
            const float inv_N = 1.0f / (float)N;
            unsigned char *my_tiles = tiles_s + (size_t)warp * 2 * TP * tile_bytes;
            int2 *my_origin = origin_s + warp * 2 * TP;
            PixelRec *my_pix = pix_s + warp * 32;
            float *my_rho = rho_s + warp * 32;
            double cost_acc = 0.0;
            unsigned long long n_tile = 0, n_fallback = 0;

            // issue the tiles of batch wb into buffer b: lane t < TP fetches the tile of point wb * TP + t
            auto prefetch = [&](int wb, int b)
            {
                // the buffer was read through the generic proxy: order those reads before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    mbar_expect_tx(bars + warp * 2 + b, (unsigned int)(TP * tile_bytes));
                __syncwarp();
                if (lane < TP)
                {
                    const int p = wb * TP + lane;
                    int2 o = make_int2(0, 0);
                    if (p < lv.P)
                    {
                        const double *xy = reinterpret_cast<const double *>(lv.xy + (size_t)p * lv.xy_stride + lv.xy_offset);
                        // 16-byte aligned inner coordinate keeps every row of the box one or two whole 32-byte sectors
                        o.x = (((int)xy[0] - box_w / 2) + 8) & ~15;
                        o.y = (int)xy[1] - box_h / 2;
                    }
                    my_origin[b * TP + lane] = o;
                    tma_load_2d(my_tiles + (size_t)(b * TP + lane) * tile_bytes, &tmap, o.x, o.y, bars + warp * 2 + b);
                }
            };

            const int stride = gridDim.x * kTmaWarps;
            int wb = blockIdx.x * kTmaWarps + warp;
            unsigned int phase_bits = 0; // parity of the next completion to wait for, per buffer
            int it = 0;
            if (wb < prm.batches_per_frame)
                prefetch(wb, 0);
            for (; wb < prm.batches_per_frame; wb += stride, ++it)
            {
                const int b = it & 1;
                if (wb + stride < prm.batches_per_frame)
                    prefetch(wb + stride, b ^ 1);
                // phase A (fp64 patch centre, live pixel) overlaps the tile fetch
                {
                    const int p = wb * TP + lane / S;
                    my_pix[lane] = setup_pixel(lv, mid_s, 0, (lane < TP * S) ? p : lv.P, lane % S, pattern_s);
                }
                __syncwarp();
                const float4 r0 = *reinterpret_cast<const float4 *>(my_pix + lane);
                const int4 r1 = *(reinterpret_cast<const int4 *>(my_pix + lane) + 1);
                const bool valid = r1.w & 1;
                PixelRegs ps;
                ps.rxy = f2(r0.x, r0.y), ps.D = r0.z, ps.iD = r0.w;
                ps.X = r1.y, ps.Y = r1.z;
                ps.lox = -(float)ps.X, ps.hix = (float)(lv.W - 1 - ps.X);
                ps.loy = -(float)ps.Y, ps.hiy = (float)(lv.H - 1 - ps.Y);
                ps.fxyiD = mul2(fxy, bc(ps.iD));
                const int t = min(lane / S, TP - 1);
                mbar_wait(bars + warp * 2 + b, (phase_bits >> b) & 1u);
                phase_bits ^= 1u << b;
                const int2 org = my_origin[b * TP + t];
                const unsigned char *tile = my_tiles + (size_t)(b * TP + t) * tile_bytes;

                float sumI = 0.f;
                if (valid)
                {
#pragma unroll 4
                    for (int i = 0; i < N; ++i)
                    {
                        const float *rec = samples_s + i * REC;
                        const float4 g0 = *reinterpret_cast<const float4 *>(rec), g1 = *reinterpret_cast<const float4 *>(rec + 4),
                                     g2 = *reinterpret_cast<const float4 *>(rec + 8);
                        // same geometry as sample_step (track_kernel.cuh): reference coordinate relative to the live pixel
                        const float2 A01 = f2(fmaf(g0.x, ps.rxy.x, fmaf(g0.y, ps.rxy.y, g1.z)), fmaf(g0.z, ps.rxy.x, fmaf(g0.w, ps.rxy.y, g1.w)));
                        const float A2 = fmaf(g1.x, ps.rxy.x, fmaf(g1.y, ps.rxy.y, g2.x));
                        const float2 m01 = add2(ps.rxy, A01);
                        const float il = rcp_approx(1.0f + A2);
                        const float tau = g2.y * ps.iD;
                        const float2 num = fma2(bc(-tau), m01, fma2(bc(-A2), ps.rxy, A01));
                        const float2 duv = mul2(fxy, fma2(num, bc(il), mul2(f2(g2.z, g2.w), bc(ps.iD))));
                        const bool ok = duv.x >= ps.lox && duv.x <= ps.hix && duv.y >= ps.loy && duv.y <= ps.hiy;
                        int xo, yo;
                        const float xf = floor_to_int(duv.x, xo), yf = floor_to_int(duv.y, yo);
                        const float dx = duv.x - xf, dy = duv.y - yf;
                        const int xi = ok ? ps.X + xo : 0, yi = ok ? ps.Y + yo : 0;
                        const float dxdy = dx * dy;
                        const float w00 = ok ? 1.0f - dx - dy + dxdy : 0.f, w01 = ok ? dx - dxdy : 0.f, w10 = ok ? dy - dxdy : 0.f,
                                    w11 = ok ? dxdy : 0.f;
                        const int tx = xi - org.x, ty = yi - org.y;
                        float I00, I01, I10, I11;
                        // inside the tile with its right / lower neighbour (the image's last column / row never is: those taps
                        // carry weight 0 and the quad texel clamps them)
                        if (tx >= 0 && ty >= 0 && tx < box_w - 1 && ty < box_h - 1 && xi < lv.W - 1 && yi < lv.H - 1)
                        {
                            const unsigned char *q = tile + ty * box_w + tx;
                            I00 = u8_to_float(q[0]), I01 = u8_to_float(q[1]), I10 = u8_to_float(q[box_w]), I11 = u8_to_float(q[box_w + 1]);
                            ++n_tile;
                        }
                        else
                        {
                            const unsigned int tq = __ldg(lv.ref_quad + (size_t)yi * lv.W + xi);
                            I00 = byte_to_float<0>(tq), I01 = byte_to_float<1>(tq), I10 = byte_to_float<2>(tq), I11 = byte_to_float<3>(tq);
                            n_fallback += ok ? 1 : 0;
                        }
                        sumI += blend4(w00, w01, w10, w11, I00, I01, I10, I11);
                    }
                }
                const float icur = __int_as_float(r1.x);
                const float r = valid ? sumI * inv_N - icur : 0.f;
                float sw;
                my_rho[lane] = (lane < TP * S) ? huber(r, prm.huber_a, sw) : 0.f;
                __syncwarp();
                if (lane < TP && wb * TP + lane < lv.P && lv.flags[wb * TP + lane] != 1)
                {
                    double sp = 0.0;
                    for (int jj = 0; jj < S; ++jj)
                        sp += (double)my_rho[lane * S + jj];
                    cost_acc += sp;
                }
                __syncwarp();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);
                n_tile += __shfl_xor_sync(0xffffffffu, n_tile, o);
                n_fallback += __shfl_xor_sync(0xffffffffu, n_fallback, o);
            }
            if (lane == 0)
            {
                red_s[warp] = cost_acc;
                atomicAdd(counters, n_tile);
                atomicAdd(counters + 1, n_fallback);
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                double s = 0.0;
                for (int w = 0; w < kTmaWarps; ++w)
                    s += red_s[w];
                atomicAdd(cost_out, s * prm.inv_num_residuals);
            }
        }
    } // namespace

    size_t track_cost_tma_smem_bytes(int K, int N, int S, int TP, int box_w, int box_h)
    {
        const int REC = sample_rec_floats(K);
        size_t b = (size_t)kTmaWarps * 2 * TP * box_w * box_h;
        b += (size_t)((N * REC + 3) & ~3) * 4 + kMidDoubles * 8 + (size_t)kTmaWarps * 32 * sizeof(PixelRec) + (size_t)((S + 1) & ~1) * 8 +
             (size_t)kTmaWarps * 2 * TP * 8 + (size_t)kTmaWarps * 2 * 8 + (size_t)kTmaWarps * 32 * 4;
        return b + 128;
    }

    // tmap: a CUtensorMap (128 bytes, built by the caller with cuTensorMapEncodeTiled over the level's 8-bit keyframe image)
    cudaError_t launch_track_cost_tma(int K, const TrackParams &prm, const void *tmap, int box_w, int box_h, int grid, size_t smem,
                                      double *cost_out, unsigned long long *counters, cudaStream_t stream)
    {
        CUtensorMap map;
        memcpy(&map, tmap, sizeof map);
        if (K == 2)
        {
            cudaError_t e = cudaFuncSetAttribute(track_cost_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            if (e != cudaSuccess)
                return e;
            track_cost_tma_kernel<2><<<grid, kTmaWarps * 32, smem, stream>>>(prm, map, box_w, box_h, cost_out, counters);
        }
        else
        {
            cudaError_t e = cudaFuncSetAttribute(track_cost_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            if (e != cudaSuccess)
                return e;
            track_cost_tma_kernel<4><<<grid, kTmaWarps * 32, smem, stream>>>(prm, map, box_w, box_h, cost_out, counters);
        }
        return cudaGetLastError();
    }
} // namespace mbavo
