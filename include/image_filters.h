// image_filters.h -- drop-in replacement for the reference's cuda_lib/include/image_filters.h.
//
// A translation unit that includes this header and links libgip_b200.so instead of the
// reference's libgpu_image_filters.a sees the same three entry points with the same C++
// linkage (the reference header has no extern "C", so the symbols are the mangled
//   _Z12gaussianBlurPhS_iiifi17OptimizationLevelP18PerformanceMetrics   (image_filters.h:46-56)
//   _Z7boxBlurPhS_iiii17OptimizationLevelP18PerformanceMetrics          (image_filters.h:72-81)
//   _Z18sobelEdgeDetectionPhS_iii17OptimizationLevelP18PerformanceMetrics (image_filters.h:103-111)
// ), the same PerformanceMetrics layout (image_filters.h:17-21) and the same
// OptimizationLevel values (image_filters.h:24-29).  The reference's bindings.cpp and tests/*.cu
// therefore compile and link against this header unchanged.
//
// Semantics kept from the reference (cuda_lib/src/image_filters.cu):
//   * d_input / d_output are device pointers owned by the caller, width*height*channels bytes,
//     row-major, interleaved channels (:95).  d_input is never written.
//   * the call is synchronous: on return d_output is complete (:894).
//   * cudaErrorNotSupported for a level the filter does not implement (:693-696, :958-961, :1615-1618).
//   * metrics may be null (:911); time_ms is CUDA-event kernel time, bandwidth_gbps is
//     4*W*H*C/2^30 per second for the blurs (:905-906) and 2*W*H*C/2^30 for Sobel (:1711-1712).
// Deliberate differences: arguments are validated (cudaErrorInvalidValue for null pointers,
// non-positive sizes, channels not in {1,3,4}, radius<0, sigma<=0); sizes are 64-bit inside,
// so images above 2 GiB work; nothing is printed unless GIP_VERBOSE=1; the Gaussian weights are
// kernel parameters, so concurrent calls do not race on a global __constant__ (:13-15).
#ifndef IMAGE_FILTERS_H
#define IMAGE_FILTERS_H

#include <cuda_runtime.h>

struct PerformanceMetrics {
    float time_ms;         // kernel time, milliseconds (CUDA events)
    float bandwidth_gbps;  // reference convention, see above
    float fps;             // 1000 / time_ms
};

enum OptimizationLevel {
    NAIVE = 1,
    SHARED_MEMORY = 2,
    TEXTURE_MEMORY = 3,
    ADVANCED = 4
};

// Separable two-pass Gaussian, clamp-to-edge, u8-rounded intermediate.  level: NAIVE or
// TEXTURE_MEMORY (both produce the reference's level-1 == level-2 values).
cudaError_t gaussianBlur(unsigned char* d_input, unsigned char* d_output,
                         int width, int height, int channels,
                         float sigma, int kernelRadius,
                         OptimizationLevel level, PerformanceMetrics* metrics);

// Separable two-pass box blur.  level: NAIVE or SHARED_MEMORY (same values; unlike the
// reference's level 2, radii above 16 are correct).
cudaError_t boxBlur(unsigned char* d_input, unsigned char* d_output,
                    int width, int height, int channels,
                    int kernelRadius,
                    OptimizationLevel level, PerformanceMetrics* metrics);

// 3x3 Sobel gradient magnitude with fused grayscale conversion; border pixels are 0 and the
// edge value is replicated into every channel.  level NAIVE keeps gray in float (reference
// level-1 values); SHARED_MEMORY rounds gray to u8 first (reference level-2 values).
cudaError_t sobelEdgeDetection(unsigned char* d_input, unsigned char* d_output,
                               int width, int height, int channels,
                               OptimizationLevel level, PerformanceMetrics* metrics);

#endif  // IMAGE_FILTERS_H
