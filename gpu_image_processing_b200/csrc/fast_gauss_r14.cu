// fast_gauss_r14.cu -- radius 14 instantiation of the Gaussian kernels (one translation unit per radius
// so that the 45 H-pass and 15 V-pass kernels compile in parallel; see fast_gauss_impl.cuh).
#include "fast_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_run_r14(const Job& job, cudaStream_t stream) { return run_radius<14>(job, stream); }
}  // namespace gip
