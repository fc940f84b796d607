// fast_gauss.cu -- dispatch of the fused Gaussian path by radius (kernels: fast_gauss_impl.cuh).
#include <cstdlib>
#include "common.cuh"

namespace gip {

constexpr int kShiftMinDefault = 16;

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
cudaError_t gauss_shift_r05(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r06(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r07(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r08(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r09(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r10(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r11(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r12(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r13(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r14(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r15(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r16(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r17(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r18(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r19(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r20(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r21(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r22(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r23(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r24(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r25(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r26(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r27(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r28(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r29(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r30(const Job& job, cudaStream_t stream);
cudaError_t gauss_shift_r31(const Job& job, cudaStream_t stream);

cudaError_t gauss_fused_r01(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r02(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r03(const Job& job, cudaStream_t stream, bool* handled);
cudaError_t gauss_fused_r04(const Job& job, cudaStream_t stream, bool* handled);

cudaError_t launch_fast_gauss(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int r = job.radius;
    if (r < 1 || r > 31) return cudaSuccess;
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
    // radius >= shift_min: shift formulation (every radius above 15; GIP_GAUSS_SHIFT_MIN moves the switch-over for A/B runs)
    static const int shift_min = [] { const char* e = getenv("GIP_GAUSS_SHIFT_MIN"); const int v = e ? atoi(e) : 0; return v >= 5 ? v : kShiftMinDefault; }();
    typedef cudaError_t (*RunFn)(const Job&, cudaStream_t);
    static const RunFn rotate[16] = {nullptr, gauss_run_r01, gauss_run_r02, gauss_run_r03, gauss_run_r04, gauss_run_r05, gauss_run_r06,
                                     gauss_run_r07, gauss_run_r08, gauss_run_r09, gauss_run_r10, gauss_run_r11, gauss_run_r12,
                                     gauss_run_r13, gauss_run_r14, gauss_run_r15};
    static const RunFn shift[32] = {nullptr, nullptr, nullptr, nullptr, nullptr, gauss_shift_r05, gauss_shift_r06, gauss_shift_r07,
                                    gauss_shift_r08, gauss_shift_r09, gauss_shift_r10, gauss_shift_r11, gauss_shift_r12, gauss_shift_r13,
                                    gauss_shift_r14, gauss_shift_r15, gauss_shift_r16, gauss_shift_r17, gauss_shift_r18, gauss_shift_r19,
                                    gauss_shift_r20, gauss_shift_r21, gauss_shift_r22, gauss_shift_r23, gauss_shift_r24, gauss_shift_r25,
                                    gauss_shift_r26, gauss_shift_r27, gauss_shift_r28, gauss_shift_r29, gauss_shift_r30, gauss_shift_r31};
    err = (r > 15 || r >= shift_min) ? shift[r](job, stream) : rotate[r](job, stream);
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
