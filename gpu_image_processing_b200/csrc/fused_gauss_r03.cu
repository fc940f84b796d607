// fused_gauss_r03.cu -- the single-kernel Gaussian of fused_gauss_impl.cuh for radius 3 (one translation unit per radius: they compile in parallel)
#include "fused_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_fused_r03(const Job& job, cudaStream_t stream, bool* handled) { return run_fused_radius<3>(job, stream, handled); }
}  // namespace gip
