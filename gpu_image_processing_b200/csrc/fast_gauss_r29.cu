// fast_gauss_r29.cu -- radius 29 instantiation of the two-kernel Gaussian (one translation unit per radius so that
// they compile in parallel; see fast_gauss_impl.cuh).  rotate = rotating accumulators, unrolled 2R+1 times (radius <= 15);
// shift = partial sums move through the FMA destination, rolled loops (radius >= 5).
#include "fast_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_shift_r29(const Job& job, cudaStream_t stream) { return run_radius<29, true>(job, stream); }
}  // namespace gip
