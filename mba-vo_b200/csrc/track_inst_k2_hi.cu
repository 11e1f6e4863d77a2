// Instantiations of the tracking kernel: spline order k = 2, knot windows [4, 5, 6] (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k2_hi(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                  int *query_occupancy, bool dependent)
    {
        if (!with_j)
            return cudaErrorInvalidValue;
        if (NK == 4)
            return dispatch_variant<2, 4, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 5)
            return dispatch_variant<2, 5, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 6)
            return dispatch_variant<2, 6, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
