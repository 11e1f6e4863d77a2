// Instantiations of the persistent sweep kernel (track_kernel.cuh): spline order k = 2, knot windows 4 and 5.
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t sweep_dispatch_k2_hi(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                     bool dependent, int *query_occupancy)
    {
        if (NK == 4)
            return launch_sweep_one<2, 4, true>(sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        if (NK == 5)
            return launch_sweep_one<2, 5, true>(sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        return cudaErrorNotSupported;
    }
} // namespace mbavo
