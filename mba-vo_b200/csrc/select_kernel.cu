// Semi-dense host-map point selection on the GPU — the keyframe-side caller of the hot path:
//   FeatureDetectorSemiDense::detect          src/core/feature_detectors/FeatureDetectorSemiDense.cpp:16-59
//   FeatureDetectorBase::gridSelection        src/core/feature_detectors/FeatureDetectorBase.cpp:49-92
//   gradient magnitude                        src/core/image_proc/Gradient.h:33-72
//   depth look-up of tmpProcessKeyframe       src/ba_tracker/blur_aware_direct_tracker.cpp:389-409
//
// Byte / index work, bit-exact with the reference: the halved central differences of 8-bit pixels, their squares and the sum
// of the squares are exact in fp32 and the square root is correctly rounded, so the magnitude has one possible value; the
// strongest pixel of a cell is the maximum of the 64-bit key (magnitude bits, ~row-major index), which is the first strongest
// pixel of the reference's row-major scan; the selected points leave in cell order through an ordered compaction.
//
// Two launches for the whole pyramid: cells_kernel (one block per grid cell of every level) and compact_kernel (one block per
// level).  The pyramid images (0.4 MB at VGA) are L2-resident from the pyramid construction that precedes the call.
#include "mbavo_device.h"

namespace mbavo
{
    namespace
    {
        constexpr int kCellThreads = 128;

        __device__ __forceinline__ float grad_mag(const unsigned char *__restrict__ I, int H, int W, int x, int y)
        {
            if (x == 0 || y == 0 || x == W - 1 || y == H - 1)
                return 0.0f; // Gradient.h:37-49
            const unsigned char *p = I + (size_t)y * W + x;
            const float dx = 0.5f * ((float)__ldg(p + 1) - (float)__ldg(p - 1));
            const float dy = 0.5f * ((float)__ldg(p + W) - (float)__ldg(p - W));
            return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))); // every step exact but the correctly rounded root
        }

        __global__ void __launch_bounds__(kCellThreads) cells_kernel(const SelectParams prm)
        {
            int lv = 0;
            while (lv + 1 < prm.n_levels && (int)blockIdx.x >= prm.lv[lv + 1].cell_base)
                ++lv;
            const SelectLevel &L = prm.lv[lv];
            const int cell = (int)blockIdx.x - L.cell_base;
            const int cy = cell / L.ncw, cx = cell % L.ncw;
            const int y0 = cy * L.ch, x0 = cx * L.cw;
            const int y1 = min(y0 + L.ch, L.H), x1 = min(x0 + L.cw, L.W);
            const int w = max(x1 - x0, 0), h = max(y1 - y0, 0);
            unsigned long long best = 0ull;
            for (int i = threadIdx.x; i < w * h; i += kCellThreads)
            {
                const int y = y0 + i / w, x = x0 + i % w;
                const float m = grad_mag(L.I, L.H, L.W, x, y);
                if (m > prm.score_threshold) // FeatureDetectorSemiDense.cpp:33
                {
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)(y * L.W + x));
                    best = key > best ? key : best; // m >= 0: its bit pattern orders like its value
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            __shared__ unsigned long long warp_best[kCellThreads / 32];
            if ((threadIdx.x & 31) == 0)
                warp_best[threadIdx.x >> 5] = best;
            __syncthreads();
            if (threadIdx.x == 0)
            {
#pragma unroll
                for (int i = 1; i < kCellThreads / 32; ++i)
                    best = warp_best[i] > best ? warp_best[i] : best;
                int4 rec = make_int4(0, 0, 0, 0);
                const float m = __uint_as_float((unsigned int)(best >> 32));
                // an untouched cell keeps OpenCV's zero response and is dropped; so is a response below 1e-6 (Base.cpp:85-88)
                if (best != 0ull && !((double)m < 1e-6))
                {
                    const unsigned int lin = 0xffffffffu - (unsigned int)(best & 0xffffffffull);
                    const int y = (int)(lin / (unsigned int)L.W), x = (int)(lin % (unsigned int)L.W);
                    // (int)(pt.x * 2^lv + 0.5) of an integer pt.x (tracker.cpp:397-398)
                    const float z = __ldg(prm.depth_z + (size_t)(y << lv) * prm.W0 + (x << lv));
                    if (!((double)z < 1e-2)) // :401-404 (a NaN depth passes, as in the reference)
                        rec = make_int4(x, y, __float_as_int(z), 1);
                }
                prm.cell_rec[blockIdx.x] = rec;
            }
        }

        // ordered compaction of one level's cell records: the points leave in cell order (Base.cpp:83-90)
        __global__ void __launch_bounds__(1024) compact_kernel(const SelectParams prm)
        {
            const SelectLevel &L = prm.lv[blockIdx.x];
            const int4 *rec = prm.cell_rec + L.cell_base;
            const int n_cells = L.ncw * L.nch;
            __shared__ int warp_count[32];
            __shared__ int base_s;
            if (threadIdx.x == 0)
                base_s = 0;
            __syncthreads();
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int c0 = 0; c0 < n_cells; c0 += 1024)
            {
                const int c = c0 + threadIdx.x;
                int4 r = make_int4(0, 0, 0, 0);
                if (c < n_cells)
                    r = rec[c];
                const unsigned int vote = __ballot_sync(0xffffffffu, r.w != 0);
                if (lane == 0)
                    warp_count[warp] = __popc(vote);
                __syncthreads();
                int before = base_s, total = 0;
                for (int i = 0; i < 32; ++i)
                {
                    const int n = warp_count[i];
                    before += i < warp ? n : 0;
                    total += n;
                }
                const int slot = before + __popc(vote & ((1u << lane) - 1u));
                if (r.w != 0 && slot < prm.capacity)
                {
                    L.xy[slot] = make_double2((double)r.x, (double)r.y);
                    L.z[slot] = (double)__int_as_float(r.z);
                }
                __syncthreads();
                if (threadIdx.x == 0)
                    base_s += total;
                __syncthreads();
            }
            if (threadIdx.x == 0)
                prm.count[blockIdx.x] = base_s;
        }
    } // namespace

    cudaError_t launch_select_kernels(const SelectParams &prm, int total_cells, cudaStream_t stream)
    {
        cells_kernel<<<total_cells, kCellThreads, 0, stream>>>(prm);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
        compact_kernel<<<prm.n_levels, 1024, 0, stream>>>(prm);
        return cudaGetLastError();
    }
} // namespace mbavo
