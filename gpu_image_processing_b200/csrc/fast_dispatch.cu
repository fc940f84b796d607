// fast_dispatch.cu -- chooses a fused sm_100a kernel for a job (filled in as kernels land).
#include "common.cuh"

namespace gip {

cudaError_t launch_fast_box(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t launch_fast_gauss(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t launch_fast_sobel(const Job& job, cudaStream_t stream, bool* handled);

cudaError_t launch_fast(FilterKind kind, const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (kind == kBox) return launch_fast_box(job, stream, handled);
    if (kind == kSobel) return launch_fast_sobel(job, stream, handled);
    if (kind == kGaussian) return launch_fast_gauss(job, stream, handled);
    return cudaSuccess;
}

}  // namespace gip
