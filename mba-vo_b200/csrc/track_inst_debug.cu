// mbavo_debug_dump: the Hessian pass with DBG = true (track_kernel.cuh) for the windows the stage tests use — k = 2 with 2 or 3
// knots, k = 4 with 4 — on the texel path.  Test infrastructure of the product kernel: same code, extra stores.
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t launch_track_debug_kernel(int K, int NK, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream)
    {
        const bool packed = prm.lv.ref_pair != nullptr;
        if (K == 2 && NK == 2)
            return packed ? launch_debug_one<2, 2, true>(prm, grid, smem, stream) : launch_debug_one<2, 2, false>(prm, grid, smem, stream);
        if (K == 2 && NK == 3 && packed)
            return launch_debug_one<2, 3, true>(prm, grid, smem, stream);
        if (K == 4 && NK == 4 && packed)
            return launch_debug_one<4, 4, true>(prm, grid, smem, stream);
        return cudaErrorNotSupported;
    }
} // namespace mbavo
