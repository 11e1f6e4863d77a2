// Instantiations of the tracking kernel: spline order k = 4, knot windows [4, 5] + the cost-only pass (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k4_lo(int NK, bool with_j, bool packed, bool crec, const TrackParams &prm, const void *table, dim3 grid,
                                  size_t smem, cudaStream_t stream, int *query_occupancy)
    {
        if (!with_j)
            return dispatch_variant<4, 4, false>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        if (NK == 4)
            return dispatch_variant<4, 4, true>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        if (NK == 5)
            return dispatch_variant<4, 5, true>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
