// fused_gauss_r01.cu -- the single-kernel Gaussian of fused_gauss_impl.cuh for radius 1 (one translation unit per radius: they compile in parallel)
#include "fused_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_fused_r01(const Job& job, cudaStream_t stream, bool* handled) { return run_fused_radius<1>(job, stream, handled); }
}  // namespace gip
