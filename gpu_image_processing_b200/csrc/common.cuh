// common.cuh -- shared declarations of the B200 filter library (host + device).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gip {

constexpr int kMaxFusedRadius = 31;          // 2r+1 <= 63 weights travel as kernel parameters
constexpr int kMaxTaps = 2 * kMaxFusedRadius + 1;

enum FilterKind : int { kGaussian = 0, kBox = 1, kSobel = 2 };

// Where the rows of one image live.  A whole image is the degenerate case
// (band_y0 = 0, band_y1 = height, above = below = nullptr).  For a row band, rows above /
// below the band are read through `above` / `below`, which may be peer-GPU memory (NVLink P2P).
struct RowSrc {
    const uint8_t* band;    // image row band_y0
    const uint8_t* above;   // image row above_y0 (rows [above_y0, band_y0)), or nullptr
    const uint8_t* below;   // image row band_y1, or nullptr
    int64_t band_y0, band_y1, above_y0;
    int64_t pitch;          // bytes per row (width * channels)
    int64_t image_stride;   // bytes between images of a batch (band pointer only)

    // y must already be clamped to [0, height).
    __host__ __device__ __forceinline__ const uint8_t* row(int64_t y, int64_t img) const {
        if (y < band_y0) return above + (y - above_y0) * pitch;
        if (y >= band_y1) return below + (y - band_y1) * pitch;
        return band + img * image_stride + (y - band_y0) * pitch;
    }
};

struct Job {
    RowSrc src;
    uint8_t* out;           // output row band_y0 of image 0
    int64_t width, height;  // full image size in pixels
    int channels;
    int64_t batch;
    int radius;             // blurs
    int sobel_u8_gray;      // Sobel: 1 = round gray to u8 before the stencil (reference level 2)
    float weights[kMaxTaps + 1];   // Gaussian taps, index radius+i; unused entries are 0
};

// Gaussian taps for the general path when radius > kMaxFusedRadius live in global memory.
struct WideWeights { const float* d_weights; };

cudaError_t launch_general(FilterKind kind, const Job& job, const float* d_wide_weights,
                           cudaStream_t stream);
cudaError_t launch_fast(FilterKind kind, const Job& job, cudaStream_t stream, bool* handled);

void count_launch(int n = 1);
// Stream-ordered scratch from the library's own memory pool of the current device (freed with cudaFreeAsync).  The pool
// keeps what it has been given until gip_release_cache(); the device's default pool -- shared with any other
// cudaMallocAsync user in the process -- is left alone.
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t stream);
int num_sms();   // SM count of the current device (cached per device)

__host__ __device__ __forceinline__ int64_t clamp64(int64_t v, int64_t lo, int64_t hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

}  // namespace gip
