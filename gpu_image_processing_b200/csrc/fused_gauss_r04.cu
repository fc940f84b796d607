// fused_gauss_r04.cu -- the single-kernel Gaussian of fused_gauss_impl.cuh for radius 4 (one translation unit per radius: they compile in parallel)
#include "fused_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_fused_r04(const Job& job, cudaStream_t stream, bool* handled) { return run_fused_radius<4>(job, stream, handled); }
}  // namespace gip
