// fast_gauss_r09.cu -- radius 9 instantiation of the Gaussian kernels (one translation unit per radius
// so that the 45 H-pass and 15 V-pass kernels compile in parallel; see fast_gauss_impl.cuh).
#include "fast_gauss_impl.cuh"

namespace gip {
cudaError_t gauss_run_r09(const Job& job, cudaStream_t stream) { return run_radius<9>(job, stream); }
}  // namespace gip
