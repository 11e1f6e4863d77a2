// Instantiation of the persistent sweep kernel (track_kernel.cuh): cubic spline (k = 4), one segment (window of 4 knots).
#include "track_kernel.cuh"

namespace mbavo
{
    cudaError_t sweep_dispatch_k4(int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem, cudaStream_t stream,
                                  bool dependent, int *query_occupancy)
    {
        if (NK == 4)
            return launch_sweep_one<4, 4, true>(sp, stage, num_sms, smem, stream, dependent, query_occupancy);
        return cudaErrorNotSupported;
    }
} // namespace mbavo
