// Instantiations of the tracking kernel: spline order k = 4, knot windows [6, 7] (see track_kernel.cuh).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t track_dispatch_k4_hi(int NK, bool with_j, bool packed, bool big, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream,
                                  int *query_occupancy, bool dependent)
    {
        if (!with_j)
            return cudaErrorInvalidValue;
        if (NK == 6)
            return dispatch_variant<4, 6, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        if (NK == 7)
            return dispatch_variant<4, 7, true>(packed, big, prm, grid, smem, stream, query_occupancy, dependent);
        return cudaErrorInvalidValue;
    }
} // namespace mbavo
