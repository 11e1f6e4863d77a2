// Instantiations of the tracking kernel: spline order k = 4, knot windows [4, 5] + the cost-only pass (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k4_lo(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                  int *query_occupancy, bool dependent)
    {
        if (!with_j)
            return dispatch_variant<4, 4, false>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 4)
            return dispatch_variant<4, 4, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 5)
            return dispatch_variant<4, 5, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
