// fast_gauss_r15.cu -- radius 15 instantiation of the two-kernel Gaussian (one translation unit per radius so that
// they compile in parallel; see fast_gauss_impl.cuh).  rotate = rotating accumulators, unrolled 2R+1 times (radius <= 15);
// shift = partial sums move through the FMA destination, rolled loops (radius >= 5).
#include "fast_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_run_r15(const Job& job, cudaStream_t stream) { return run_radius<15, false>(job, stream); }
cudaError_t gauss_shift_r15(const Job& job, cudaStream_t stream) { return run_radius<15, true>(job, stream); }
}  // namespace gip
