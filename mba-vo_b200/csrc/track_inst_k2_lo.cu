// Instantiations of the tracking kernel: spline order k = 2, knot windows [2, 3] + the cost-only pass (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k2_lo(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                  int *query_occupancy, bool dependent)
    {
        if (!with_j)
            return dispatch_variant<2, 2, false>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 2)
            return dispatch_variant<2, 2, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 3)
            return dispatch_variant<2, 3, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
