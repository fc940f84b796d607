// general_kernels.cu -- the general path: any radius, any alignment, any size.
//
// Two-pass blurs through a u8 scratch image and a one-pass Sobel, one thread per output byte
// (blurs) or pixel (Sobel).  It exists for the shapes the fused kernels do not take
// (radius > 31) and as the in-library cross-check of the fused kernels (gip_set_path(1)).
// Arithmetic follows /root/reference/cuda_lib/src/image_filters.cu with every rounding step
// written as an explicit intrinsic, in the order the reference's SASS performs them:
//   Gaussian tap   sum = fma(pixel, w, sum)                        (:98, :139)
//   Gaussian store (uchar)(sum + 0.5f)                             (:102, :142)
//   box store      (uchar)fma(sum, 1.0f/k, 0.5f)                   (:394, :429)
//   gray           fma(B,.114f, fma(R,.299f, G*.587f))             (:1245)
//   Sobel          row-major single-rounded adds, fma(gx,gx,gy*gy), sqrtf, fminf, +0.5f (:1246-1305)
#include "common.cuh"

namespace gip {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float tap_weight(const Job& job, const float* wide, int idx) {
    return wide ? wide[idx] : job.weights[idx];
}

// Horizontal pass over image rows [ty0, ty1) into the scratch image `tmp`
// (row ty0 of image `img` at tmp + img*trows*pitch).
template <bool kBox>
__global__ void __launch_bounds__(kThreads)
gip_blur_h_general(const __grid_constant__ Job job, const float* __restrict__ wide,
                   uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1, int64_t img0, int64_t nimg) {
    const int64_t pitch = job.src.pitch;
    const int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= pitch) return;
    const int C = job.channels, r = job.radius;
    const int64_t x = b / C;
    const int ch = (int)(b - x * C);
    const int64_t trows = ty1 - ty0;
    const float inv = __fdiv_rn(1.0f, (float)(2 * r + 1));
    for (int64_t i = blockIdx.z; i < nimg; i += gridDim.z) {
        for (int64_t y = ty0 + blockIdx.y; y < ty1; y += gridDim.y) {
            const uint8_t* row = job.src.row(y, img0 + i);
            float sum = 0.0f;
            for (int t = -r; t <= r; t++) {
                const int64_t nx = clamp64(x + t, 0, job.width - 1);
                const float p = (float)row[nx * C + ch];
                if (kBox) sum = __fadd_rn(sum, p);
                else      sum = __fmaf_rn(p, tap_weight(job, wide, r + t), sum);
            }
            const float v = kBox ? __fmaf_rn(sum, inv, 0.5f) : __fadd_rn(sum, 0.5f);
            tmp[(i * trows + (y - ty0)) * pitch + b] = (uint8_t)v;
        }
    }
}

// Vertical pass: output rows [band_y0, band_y1) from the scratch rows [ty0, ty1).
template <bool kBox>
__global__ void __launch_bounds__(kThreads)
gip_blur_v_general(const __grid_constant__ Job job, const float* __restrict__ wide,
                   const uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1, int64_t img0, int64_t nimg) {
    const int64_t pitch = job.src.pitch;
    const int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= pitch) return;
    const int r = job.radius;
    const int64_t trows = ty1 - ty0;
    const int64_t y0 = job.src.band_y0, y1 = job.src.band_y1;
    const float inv = __fdiv_rn(1.0f, (float)(2 * r + 1));
    for (int64_t i = blockIdx.z; i < nimg; i += gridDim.z) {
        const uint8_t* timg = tmp + i * trows * pitch;
        uint8_t* oimg = job.out + (img0 + i) * job.src.image_stride;
        for (int64_t y = y0 + blockIdx.y; y < y1; y += gridDim.y) {
            float sum = 0.0f;
            for (int t = -r; t <= r; t++) {
                const int64_t ny = clamp64(y + t, 0, job.height - 1);
                const float p = (float)timg[(ny - ty0) * pitch + b];
                if (kBox) sum = __fadd_rn(sum, p);
                else      sum = __fmaf_rn(p, tap_weight(job, wide, r + t), sum);
            }
            const float v = kBox ? __fmaf_rn(sum, inv, 0.5f) : __fadd_rn(sum, 0.5f);
            oimg[(y - y0) * pitch + b] = (uint8_t)v;
        }
    }
}

__device__ __forceinline__ float gray_of(const uint8_t* px, int C, int u8_gray) {
    if (C == 1) return (float)px[0];
    float g = __fmul_rn((float)px[1], 0.587f);
    g = __fmaf_rn((float)px[0], 0.299f, g);
    g = __fmaf_rn((float)px[2], 0.114f, g);
    if (u8_gray) g = (float)(uint8_t)__fadd_rn(g, 0.5f);
    return g;
}

__global__ void __launch_bounds__(kThreads)
gip_sobel_general(const __grid_constant__ Job job) {
    const int64_t W = job.width, H = job.height;
    const int64_t x = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (x >= W) return;
    const int C = job.channels;
    const int64_t y0 = job.src.band_y0, y1 = job.src.band_y1;
    for (int64_t i = blockIdx.z; i < job.batch; i += gridDim.z) {
        uint8_t* oimg = job.out + i * job.src.image_stride;
        for (int64_t y = y0 + blockIdx.y; y < y1; y += gridDim.y) {
            uint8_t* o = oimg + (y - y0) * job.src.pitch + x * C;
            uint8_t e = 0;
            if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
                float g[3][3];
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    const uint8_t* row = job.src.row(y + dy - 1, i);
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
                        g[dy][dx] = gray_of(row + (x + dx - 1) * C, C, job.sobel_u8_gray);
                }
                float gx = __fsub_rn(0.0f, g[0][0]);
                float gy = gx;
                gy = __fsub_rn(gy, __fadd_rn(g[0][1], g[0][1]));
                gx = __fadd_rn(gx, g[0][2]);
                gy = __fsub_rn(gy, g[0][2]);
                gx = __fsub_rn(gx, __fadd_rn(g[1][0], g[1][0]));
                gx = __fadd_rn(gx, __fadd_rn(g[1][2], g[1][2]));
                gx = __fsub_rn(gx, g[2][0]);
                gy = __fadd_rn(gy, g[2][0]);
                gy = __fadd_rn(gy, __fadd_rn(g[2][1], g[2][1]));
                gx = __fadd_rn(gx, g[2][2]);
                gy = __fadd_rn(gy, g[2][2]);
                float m = __fsqrt_rn(__fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
                m = fminf(m, 255.0f);
                e = (uint8_t)__fadd_rn(m, 0.5f);
            }
            for (int c = 0; c < C; c++) o[c] = e;
        }
    }
}

unsigned grid_y(int64_t rows) { return (unsigned)(rows < 32768 ? (rows > 0 ? rows : 1) : 32768); }
unsigned grid_z(int64_t n) { return (unsigned)(n < 1024 ? (n > 0 ? n : 1) : 1024); }

}  // namespace

cudaError_t launch_general(FilterKind kind, const Job& job, const float* d_wide, cudaStream_t stream) {
    const int64_t pitch = job.src.pitch;
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (kind == kSobel) {
        dim3 grid((unsigned)((job.width + kThreads - 1) / kThreads), grid_y(rows), grid_z(job.batch));
        gip_sobel_general<<<grid, kThreads, 0, stream>>>(job);
        count_launch();
        return cudaGetLastError();
    }
    const int r = job.radius;
    const int64_t ty0 = clamp64(job.src.band_y0 - r, 0, job.height);
    const int64_t ty1 = clamp64(job.src.band_y1 + r, 0, job.height);
    const int64_t trows = ty1 - ty0;
    // bound the scratch image: process the batch in chunks of whole images (<= 1 GiB of scratch)
    int64_t chunk = (int64_t(1) << 30) / (trows * pitch);
    if (chunk < 1) chunk = 1;
    if (chunk > job.batch) chunk = job.batch;
    uint8_t* tmp = nullptr;
    cudaError_t err = scratch_alloc((void**)&tmp, (size_t)(chunk * trows * pitch), stream);
    if (err != cudaSuccess) return err;
    const unsigned gx = (unsigned)((pitch + kThreads - 1) / kThreads);
    for (int64_t img0 = 0; img0 < job.batch && err == cudaSuccess; img0 += chunk) {
        const int64_t n = (job.batch - img0 < chunk) ? job.batch - img0 : chunk;
        dim3 gh(gx, grid_y(trows), grid_z(n)), gv(gx, grid_y(rows), grid_z(n));
        if (kind == kBox) {
            gip_blur_h_general<true><<<gh, kThreads, 0, stream>>>(job, d_wide, tmp, ty0, ty1, img0, n);
            gip_blur_v_general<true><<<gv, kThreads, 0, stream>>>(job, d_wide, tmp, ty0, ty1, img0, n);
        } else {
            gip_blur_h_general<false><<<gh, kThreads, 0, stream>>>(job, d_wide, tmp, ty0, ty1, img0, n);
            gip_blur_v_general<false><<<gv, kThreads, 0, stream>>>(job, d_wide, tmp, ty0, ty1, img0, n);
        }
        count_launch(2);
        err = cudaGetLastError();
    }
    cudaError_t ferr = cudaFreeAsync(tmp, stream);
    return err != cudaSuccess ? err : ferr;
}

}  // namespace gip
