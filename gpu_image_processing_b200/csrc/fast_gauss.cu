// fast_gauss.cu -- dispatch of the fused Gaussian path by radius (kernels: fast_gauss_impl.cuh).
#include <cstdlib>
#include "common.cuh"

namespace gip {

cudaError_t gauss_run_r01(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r02(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r03(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r04(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r05(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r06(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r07(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r08(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r09(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r10(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r11(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r12(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r13(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r14(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r15(const Job& job, cudaStream_t stream);

cudaError_t gauss_fused_r01(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r02(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r03(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r04(const Job& job, cudaStream_t stream, bool* handled);

cudaError_t launch_fast_gauss(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int r = job.radius;
    if (r < 1 || r > 15) return cudaSuccess;
    const int64_t pitch = job.src.pitch;
    if (job.src.band_y1 - job.src.band_y0 > 0x3fffffff || job.height > 0x3fffffff || pitch > 0x7fffffff) return cudaSuccess;
    if (num_sms() <= 0) return cudaErrorInvalidDevice;
    cudaError_t err;
    // radius <= 4 on 16-byte aligned rows: one fused kernel, no scratch image (fused_gauss_impl.cuh)
    static const int no_fused = [] { const char* e = getenv("GIP_GAUSS_NO_FUSED"); return e ? atoi(e) : 0; }();   // A/B runs
    const bool aligned16 = (pitch % 16 == 0) && (job.src.image_stride % 16 == 0) && ((uintptr_t)job.src.band % 16 == 0) &&
                           ((uintptr_t)job.out % 16 == 0) && (!job.src.above || (uintptr_t)job.src.above % 16 == 0) &&
                           (!job.src.below || (uintptr_t)job.src.below % 16 == 0);
    if (r <= 4 && aligned16 && !no_fused) {
        switch (r) {
            case 1: err = gauss_fused_r01(job, stream, handled); break;
            case 2: err = gauss_fused_r02(job, stream, handled); break;
            case 3: err = gauss_fused_r03(job, stream, handled); break;
            default: err = gauss_fused_r04(job, stream, handled); break;
        }
        if (err != cudaSuccess || *handled) return err;
    }
    switch (r) {
        case 1: err = gauss_run_r01(job, stream); break;
        case 2: err = gauss_run_r02(job, stream); break;
        case 3: err = gauss_run_r03(job, stream); break;
        case 4: err = gauss_run_r04(job, stream); break;
        case 5: err = gauss_run_r05(job, stream); break;
        case 6: err = gauss_run_r06(job, stream); break;
        case 7: err = gauss_run_r07(job, stream); break;
        case 8: err = gauss_run_r08(job, stream); break;
        case 9: err = gauss_run_r09(job, stream); break;
        case 10: err = gauss_run_r10(job, stream); break;
        case 11: err = gauss_run_r11(job, stream); break;
        case 12: err = gauss_run_r12(job, stream); break;
        case 13: err = gauss_run_r13(job, stream); break;
        case 14: err = gauss_run_r14(job, stream); break;
        case 15: err = gauss_run_r15(job, stream); break;
        default: return cudaSuccess;
    }
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
