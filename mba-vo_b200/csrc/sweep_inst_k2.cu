// Instantiations of the persistent sweep kernel (track_kernel.cuh): spline order k = 2, knot windows 2 and 3 — the tracker's own
// mode (two knots) and an exposure that straddles a knot.
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t sweep_dispatch_k2(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                  bool dependent, int *query_occupancy)
    {
        if (NK == 2)
            return launch_sweep_one<2, 2, true>(sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        if (NK == 3)
            return launch_sweep_one<2, 3, true>(sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        return cudaErrorNotSupported;
    }
} // namespace mbavo
