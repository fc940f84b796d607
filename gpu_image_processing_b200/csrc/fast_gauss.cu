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
cudaError_t gauss_run_r16(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r17(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r18(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r19(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r20(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r21(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r22(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r23(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r24(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r25(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r26(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r27(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r28(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r29(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r30(const Job& job, cudaStream_t stream);
cudaError_t gauss_run_r31(const Job& job, cudaStream_t stream);

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
    // radius <= 4: one fused kernel, no scratch image, rows at any byte alignment (fused_gauss_impl.cuh)
    static const int no_fused = [] { const char* e = getenv("GIP_GAUSS_NO_FUSED"); return e ? atoi(e) : 0; }();   // A/B runs
    if (r <= 4 && !no_fused) {
        switch (r) {
            case 1: err = gauss_fused_r01(job, stream, handled); break;
            case 2: err = gauss_fused_r02(job, stream, handled); break;
            case 3: err = gauss_fused_r03(job, stream, handled); break;
            default: err = gauss_fused_r04(job, stream, handled); break;
        }
        if (err != cudaSuccess || *handled) return err;
    }
    // radius <= 15: rotating accumulators; 16..31: shift formulation.  Measured on the B200 (8K RGB): the rotating form is
    // faster up to 14 (r = 8: 177 vs 200 us, r = 12: 239 vs 250 us), equal at 15 (293 us), and does not fit the register
    // file beyond.
    typedef cudaError_t (*RunFn)(const Job&, cudaStream_t);
    static const RunFn run[32] = {nullptr,
                                  gauss_run_r01, gauss_run_r02, gauss_run_r03, gauss_run_r04, gauss_run_r05, gauss_run_r06,
                                  gauss_run_r07, gauss_run_r08, gauss_run_r09, gauss_run_r10, gauss_run_r11, gauss_run_r12,
                                  gauss_run_r13, gauss_run_r14, gauss_run_r15, gauss_run_r16, gauss_run_r17, gauss_run_r18,
                                  gauss_run_r19, gauss_run_r20, gauss_run_r21, gauss_run_r22, gauss_run_r23, gauss_run_r24,
                                  gauss_run_r25, gauss_run_r26, gauss_run_r27, gauss_run_r28, gauss_run_r29, gauss_run_r30,
                                  gauss_run_r31};
    err = run[r](job, stream);
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
