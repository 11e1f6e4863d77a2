// Instantiations of the tracking kernel: spline order k = 2, knot windows [2, 3] + the cost-only pass (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k2_lo(int NK, bool with_j, bool packed, bool crec, const TrackParams &prm, const void *table, dim3 grid,
                                  size_t smem, cudaStream_t stream, int *query_occupancy)
    {
        if (!with_j)
            return dispatch_variant<2, 2, false>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        if (NK == 2)
            return dispatch_variant<2, 2, true>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        if (NK == 3)
            return dispatch_variant<2, 3, true>(packed, crec, prm, table, grid, smem, stream, query_occupancy);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
