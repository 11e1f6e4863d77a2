// Tracking-kernel plumbing shared by the instantiation units: keyframe texel packing, shared-memory sizing, dispatch.
#include "mbavo_device.h"

#include <cuda_fp16.h>

namespace mbavo
{
    cudaError_t track_dispatch_k2_lo(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                     cudaStream_t stream, int *query_occupancy, bool dependent);
    cudaError_t track_dispatch_k2_hi(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                     cudaStream_t stream, int *query_occupancy, bool dependent);
    cudaError_t track_dispatch_k4_lo(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                     cudaStream_t stream, int *query_occupancy, bool dependent);
    cudaError_t track_dispatch_k4_hi(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                     cudaStream_t stream, int *query_occupancy, bool dependent);

    namespace
    {
        // the 16-byte row pair of a pair texel (see LevelDev): (gx gy)(x,y) | (gx gy)(x+1,y) | I(x,y) I(x+1,y) | 0
        __device__ __forceinline__ uint4 row_pair(float2 g0, float2 g1, unsigned int b0, unsigned int b1)
        {
            uint4 t;
            t.x = (unsigned int)__half_as_ushort(__float2half_rn(g0.x)) | ((unsigned int)__half_as_ushort(__float2half_rn(g0.y)) << 16);
            t.y = (unsigned int)__half_as_ushort(__float2half_rn(g1.x)) | ((unsigned int)__half_as_ushort(__float2half_rn(g1.y)) << 16);
            t.z = (unsigned int)__half_as_ushort(__float2half_rn((float)b0)) | ((unsigned int)__half_as_ushort(__float2half_rn((float)b1)) << 16);
            t.w = 0u;
            return t;
        }
        // free byte of a patch texel: the byte whose doubled central difference against `fixed` is sign * 2 g.  false when no
        // byte reproduces g exactly (g not a whole number of half grey levels, or out of a byte's reach)
        __device__ __forceinline__ bool free_byte(unsigned int fixed, float g, float sign, unsigned int &byte)
        {
            const float v = fminf(fmaxf(rintf((float)fixed + sign * 2.0f * g), 0.f), 255.f); // NaN -> 0
            byte = (unsigned int)v;
            return 0.5f * sign * (v - (float)fixed) == g;
        }
        // The keyframe texel of pixel (x, y) from the four taps of its bilinear footprint (bytes b.., gradients g..; x+1 / y+1
        // already clamped by the caller).  Returns false when the texel cannot hold the gradients exactly.
        __device__ __forceinline__ bool make_pair_texel(unsigned int b00, unsigned int b01, unsigned int b10, unsigned int b11, float2 g00,
                                                        float2 g01, float2 g10, float2 g11, PairTexel &out)
        {
#if MBAVO_TEXEL == 3
            // rows y-1 .. y+2 of the 4 x 4 neighbourhood, one word per row, columns x-1 .. x+2 from the low byte up; the middle
            // 2 x 2 are the intensities, the 8 bytes next to them are chosen so that every tap's gradient is the doubled central
            // difference the kernel takes (real neighbours for Gradient.h's gradient image; the zero gradient of a border pixel
            // comes out as a copy of the opposite byte); the corners are unused
            unsigned int r1c0, r1c3, r2c0, r2c3, r0c1, r0c2, r3c1, r3c2;
            bool ok = free_byte(b01, g00.x, -1.f, r1c0);
            ok &= free_byte(b00, g01.x, 1.f, r1c3);
            ok &= free_byte(b11, g10.x, -1.f, r2c0);
            ok &= free_byte(b10, g11.x, 1.f, r2c3);
            ok &= free_byte(b10, g00.y, -1.f, r0c1);
            ok &= free_byte(b11, g01.y, -1.f, r0c2);
            ok &= free_byte(b00, g10.y, 1.f, r3c1);
            ok &= free_byte(b01, g11.y, 1.f, r3c2);
            out.x = (r0c1 << 8) | (r0c2 << 16);
            out.y = r1c0 | (b00 << 8) | (b01 << 16) | (r1c3 << 24);
            out.z = r2c0 | (b10 << 8) | (b11 << 16) | (r2c3 << 24);
            out.w = (r3c1 << 8) | (r3c2 << 16);
            return ok;
#else
            const __half hx = __float2half_rn(g00.x), hy = __float2half_rn(g00.y);
            // bitwise round trip (also rejects NaN and values that overflow to inf); every pixel vouches for its own gradient
            const bool ok = __float_as_uint(__half2float(hx)) == __float_as_uint(g00.x) && __float_as_uint(__half2float(hy)) == __float_as_uint(g00.y);
            out = row_pair(g00, g01, b00, b01);
            return ok;
#endif
        }

        // Keyframe texels (see LevelDev).  One thread per pixel; *inexact counts the pixels whose texel cannot hold the gradients.
        __global__ void pack_kernel(const unsigned char *__restrict__ I, const float2 *__restrict__ g, int H, int W,
                                    PairTexel *__restrict__ pair, unsigned int *__restrict__ quad, int *__restrict__ inexact)
        {
            const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
            if (x >= W)
                return;
            const int x1 = min(x + 1, W - 1), y1 = min(y + 1, H - 1);
            const int i00 = y * W + x, i01 = y * W + x1, i10 = y1 * W + x, i11 = y1 * W + x1;
            const unsigned int b00 = I[i00], b01 = I[i01], b10 = I[i10], b11 = I[i11];
            quad[i00] = b00 | (b01 << 8) | (b10 << 16) | (b11 << 24);
            if (!make_pair_texel(b00, b01, b10, b11, g[i00], g[i01], g[i10], g[i11], pair[i00]))
                atomicAdd(inexact, 1);
        }

        // ImagePyramid<T>::computePyramid, one level (src/core/measurements/ImagePyramid.h:76-95): T(0.25 * (float sum of the
        // 2x2 block)) — the sum of four bytes is exact in float and 0.25 * it is exact in double, so the truncating cast is
        // the integer (a + b + c + d) >> 2.
        __global__ void pyr_down_kernel(const unsigned char *__restrict__ src, int Ws, unsigned char *__restrict__ dst, int Hd,
                                        int Wd)
        {
            const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
            if (x >= Wd)
                return;
            const unsigned char *p = src + (size_t)2 * y * Ws + 2 * x;
            dst[(size_t)y * Wd + x] = (unsigned char)(((unsigned int)p[0] + p[1] + p[Ws] + p[Ws + 1]) >> 2);
        }

        // compute_image_gradients (src/core/image_proc/Gradient.h:17-75) fused with the texel packing: 0.5 * central
        // differences (exact in fp32 and in fp16), zero on the 1-pixel border.  grad (float2, may be null) receives the
        // reference layout for the direct-gather kernels.
        __device__ __forceinline__ float2 central_gradient(const unsigned char *__restrict__ I, int H, int W, int x, int y)
        {
            if (x == 0 || y == 0 || x == W - 1 || y == H - 1)
                return make_float2(0.f, 0.f);
            const unsigned char *c = I + (size_t)y * W + x;
            return make_float2(0.5f * ((float)c[1] - (float)c[-1]), 0.5f * ((float)c[W] - (float)c[-W]));
        }
        __global__ void pack_image_kernel(const unsigned char *__restrict__ I, int H, int W, PairTexel *__restrict__ pair,
                                          unsigned int *__restrict__ quad, float2 *__restrict__ grad)
        {
            const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
            if (x >= W)
                return;
            const int x1 = min(x + 1, W - 1), y1 = min(y + 1, H - 1);
            const int i00 = y * W + x;
            const unsigned int b00 = I[i00], b01 = I[y * W + x1], b10 = I[y1 * W + x], b11 = I[y1 * W + x1];
            const float2 g0 = central_gradient(I, H, W, x, y), g1 = central_gradient(I, H, W, x1, y);
            if (grad)
                grad[i00] = g0;
            if (pair)
            {
                quad[i00] = b00 | (b01 << 8) | (b10 << 16) | (b11 << 24);
                make_pair_texel(b00, b01, b10, b11, g0, g1, central_gradient(I, H, W, x, y1), central_gradient(I, H, W, x1, y1), pair[i00]);
            }
        }
        // one pyramid step of a new frame (PyrStepParams): the block finds its job, then runs the body of pack_image_kernel /
        // pyr_down_kernel / a byte clear on its 256 outputs
        __global__ void __launch_bounds__(256) pyr_step_kernel(const __grid_constant__ PyrStepParams p)
        {
            int j = 0;
            while (j + 1 < p.n_jobs && (int)blockIdx.x >= p.job[j + 1].block_begin)
                ++j;
            const PyrJob &job = p.job[j];
            const int i = ((int)blockIdx.x - job.block_begin) * 256 + (int)threadIdx.x;
            if (job.kind == 2)
            {
                if (i < job.Wd)
                    job.dst[i] = 0;
            }
            else if (job.kind == 1)
            {
                if (i < job.Hd * job.Wd)
                {
                    const int x = i % job.Wd, y = i / job.Wd;
                    const unsigned char *q = job.src + (size_t)2 * y * job.Ws + 2 * x;
                    job.dst[i] = (unsigned char)(((unsigned int)q[0] + q[1] + q[job.Ws] + q[job.Ws + 1]) >> 2);
                }
            }
            else if (i < job.Hs * job.Ws)
            {
                const int H = job.Hs, W = job.Ws, x = i % W, y = i / W;
                const unsigned char *I = job.src;
                const int x1 = min(x + 1, W - 1), y1 = min(y + 1, H - 1);
                const unsigned int b00 = I[i], b01 = I[y * W + x1], b10 = I[y1 * W + x], b11 = I[y1 * W + x1];
                const float2 g0 = central_gradient(I, H, W, x, y), g1 = central_gradient(I, H, W, x1, y);
                if (job.grad)
                    job.grad[i] = g0;
                if (job.pair)
                {
                    job.quad[i] = b00 | (b01 << 8) | (b10 << 16) | (b11 << 24);
                    make_pair_texel(b00, b01, b10, b11, g0, g1, central_gradient(I, H, W, x, y1), central_gradient(I, H, W, x1, y1), job.pair[i]);
                }
            }
        }
    } // namespace

    // jobs[].block_begin is filled here; returns cudaErrorInvalidValue for an empty step
    cudaError_t launch_pyr_step_kernel(PyrStepParams &p, cudaStream_t stream)
    {
        int blocks = 0;
        for (int j = 0; j < p.n_jobs; ++j)
        {
            PyrJob &job = p.job[j];
            const long long n = job.kind == 0 ? (long long)job.Hs * job.Ws : (job.kind == 1 ? (long long)job.Hd * job.Wd : (long long)job.Wd);
            job.block_begin = blocks;
            blocks += (int)((n + 255) / 256);
        }
        if (blocks == 0)
            return cudaErrorInvalidValue;
        pyr_step_kernel<<<blocks, 256, 0, stream>>>(p);
        return cudaGetLastError();
    }

    cudaError_t launch_pyr_down_kernel(const unsigned char *src, int Ws, unsigned char *dst, int Hd, int Wd, cudaStream_t stream)
    {
        const dim3 block(128, 1, 1), grid((Wd + 127) / 128, Hd, 1);
        pyr_down_kernel<<<grid, block, 0, stream>>>(src, Ws, dst, Hd, Wd);
        return cudaGetLastError();
    }

    cudaError_t launch_pack_image_kernel(const unsigned char *I, int H, int W, PairTexel *pair, unsigned int *quad, float *grad,
                                         cudaStream_t stream)
    {
        const dim3 block(128, 1, 1), grid((W + 127) / 128, H, 1);
        pack_image_kernel<<<grid, block, 0, stream>>>(I, H, W, pair, quad, reinterpret_cast<float2 *>(grad));
        return cudaGetLastError();
    }

    cudaError_t sweep_dispatch_k2(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                  bool dependent, int *query_occupancy);
    cudaError_t sweep_dispatch_k2_hi(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                     bool dependent, int *query_occupancy);
    cudaError_t sweep_dispatch_k4(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                  bool dependent, int *query_occupancy);

    // persistent sweep kernel: instantiated for the texel path and the windows listed here; cudaErrorNotSupported otherwise
    // (the caller then runs the sweep pass by pass)
    cudaError_t launch_sweep_kernel(int K, int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem,
                                    cudaStream_t stream, bool dependent, int *query_occupancy)
    {
        if (K == 2 && NK <= 3)
            return sweep_dispatch_k2(NK, sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        if (K == 2)
            return sweep_dispatch_k2_hi(NK, sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        if (K == 4)
            return sweep_dispatch_k4(NK, sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        return cudaErrorNotSupported;
    }

    size_t track_pass_smem_bytes(int K, int NK, bool with_j, int kWarpsPerBlock, int N, int S, int TP);
    size_t track_kernel_smem_bytes(int K, int NK, bool with_j, bool big, int N, int S, int TP)
    {
        return track_pass_smem_bytes(K, NK, with_j, track_warps(with_j, NK, big), N, S, TP);
    }
    // both passes of a level inside the persistent sweep kernel (block shape of the big Hessian pass)
    size_t sweep_kernel_smem_bytes(int K, int NK, int N, int S, int TP)
    {
        const int w = track_warps(true, NK, true);
        const size_t a = track_pass_smem_bytes(K, NK, true, w, N, S, TP), b = track_pass_smem_bytes(K, NK, false, w, N, S, TP);
        return a > b ? a : b;
    }

    size_t track_pass_smem_bytes(int K, int NK, bool with_j, int kWarpsPerBlock, int N, int S, int TP)
    {
        const int REC = sample_rec_floats(K);
        const int D1 = with_j ? 6 * NK + 1 : 1, D1E = (D1 + 1) & ~1;
        const int PITCH = (D1E / 2) % 2 == 1 ? D1E : D1E + 2, T = D1E / 2, NT = T * (T + 1) / 2;
        const int E = with_j ? packed_len(NK) : 1;
        const int rho_per_warp = max(32, TP * S);
        const int kThreads = kWarpsPerBlock * 32;
        size_t main_bytes = (size_t)N * REC * 4 + 8 * 4 + kMidDoubles * 8 + (size_t)kWarpsPerBlock * 32 * 32 /* PixelRec */ +
                            (size_t)S * 8 +
                            (size_t)kWarpsPerBlock * rho_per_warp * 4 + (size_t)((NT + 7) & ~7) * 2 +
                            (size_t)kWarpsPerBlock * 32 * PITCH * 4;
        const int MT = with_j ? (NT + 31) / 32 : 1;
        main_bytes += 8 + (with_j ? (size_t)kWarpsPerBlock * MT * 4 * 32 * 8 : 0); // fp64 tile accumulators
        size_t red_bytes = (size_t)(kWarpsPerBlock > (kThreads / E + 1) ? kWarpsPerBlock : kThreads / E + 1) * E * 8;
        // the device-resident Gauss-Newton step of the last block: packed vector + dense window system + 4 vectors
        const size_t solve_bytes = with_j ? ((size_t)((E + 1) & ~1) + 108 * NK * NK + 18 * NK + 16) * 8 : 0; // + the unit-lower factor of ldlt_solve_rows
        size_t need = main_bytes > red_bytes ? main_bytes : red_bytes;
        need = need > solve_bytes ? need : solve_bytes;
        return need + 16;
    }

    cudaError_t launch_pack_kernel(const unsigned char *I, const float *dIxy, int H, int W, PairTexel *pair, unsigned int *quad,
                                   int *inexact, cudaStream_t stream)
    {
        const dim3 block(128, 1, 1), grid((W + 127) / 128, H, 1);
        pack_kernel<<<grid, block, 0, stream>>>(I, reinterpret_cast<const float2 *>(dIxy), H, W, pair, quad, inexact);
        return cudaGetLastError();
    }

    // with_j: Hessian pass (templated on the window) or cost-only pass (one instantiation per K).  dependent: launch with
    // programmatic stream serialisation (the previous kernel in the stream is the pose kernel).
    cudaError_t launch_track_kernel(int K, int NK, bool with_j, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                    cudaStream_t stream, int *query_occupancy, bool dependent)
    {
        const bool packed = prm.lv.ref_pair != nullptr;
        if (K == 2)
            return (!with_j || NK <= 3) ? track_dispatch_k2_lo(NK, with_j, packed, big, prm, grid, smem, stream, query_occupancy, dependent)
                                        : track_dispatch_k2_hi(NK, with_j, packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (K == 4)
            return (!with_j || NK <= 5) ? track_dispatch_k4_lo(NK, with_j, packed, big, prm, grid, smem, stream, query_occupancy, dependent)
                                        : track_dispatch_k4_hi(NK, with_j, packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
