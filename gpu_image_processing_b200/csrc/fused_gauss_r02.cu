// fused_gauss_r02.cu -- the single-kernel Gaussian of fused_gauss_impl.cuh for radius 2 (one translation unit per radius: they compile in parallel)
#include "fused_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_fused_r02(const Job& job, cudaStream_t stream, bool* handled) { return run_fused_radius<2>(job, stream, handled); }
}  // namespace gip
