// fast_gauss_r20.cu -- radius 20 instantiation of the two-kernel Gaussian (one translation unit per radius so that
// they compile in parallel; see fast_gauss_impl.cuh).  Radius <= 15: rotating accumulators, unrolled 2R+1 times.
// Radius >= 16: shift formulation, rolled loops.
#include "gauss_shift_impl.cuh"

namespace gip {
cudaError_t gauss_run_r20(const Job& job, cudaStream_t stream) { return run_radius<20, true>(job, stream); }
}  // namespace gip
