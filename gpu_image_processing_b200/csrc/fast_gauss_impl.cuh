// fast_gauss_impl.cuh -- separable Gaussian blur for sm_100a, radius 1..15, weights as kernel parameters.
//
// Replaces gaussianBlur{Horizontal,Vertical}{Naive,Level2}
// (/root/reference/cuda_lib/src/image_filters.cu:64-144, :159-347).  Two kernels, like the reference's
// two passes, with the reference's u8-rounded intermediate image (:102) between them; the work inside a
// pass is what changes.  The reference converts and multiplies every neighbour again for every output
// (2r+1 byte loads + I2F + FFMA per output byte).  Here every input byte is loaded and converted ONCE
// and scattered into the 2r+1 outputs it belongs to, which live in a rotating set of register
// accumulators (static register indices by unrolling 2r+1 steps; R = radius is a template parameter):
//   out[o] accumulates  fma(in[o-r], w[0], 0), fma(in[o-r+1], w[1], .), ... fma(in[o+r], w[2r], .)
// in exactly the reference's tap order (:86-99), so the float32 sums are bit-identical.
//   H pass  (gip_gauss_h)  a CTA stages a 64-row x 256-pixel tile (+ r pixels of halo, clamp-to-edge)
//           in shared memory with cp.async.  A thread owns one channel of a row PAIR and marches
//           along x: the two rows are the two lanes of FFMA2, so one issue slot does two taps.  Lanes of
//           a warp are 32 different row pairs (odd shared-memory pitch: conflict-free byte loads).
//           Rounded bytes go to an output tile in shared memory and leave with coalesced 32-bit stores.
//   V pass  (gip_gauss_v)  a thread owns a 4-byte column group and marches down a band of rows straight
//           from global memory (coalesced 32-bit loads); adjacent bytes are the FFMA2 lanes.
// Rounding: (uchar)(sum + 0.5f) (:102, :142) == low mantissa byte of RZ((sum + 0.5f) + 2^23).
// u8 -> float: PRMT into the mantissa of 2^23, minus 2^23 (exact), two values per FADD2.
// The algorithm is FP32-issue bound, not HBM bound, from radius 3 up (4r+2 FMAs per byte); DESIGN.md
// carries the instruction roofline next to the HBM one.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kMaxFastRadius = 15;
constexpr int kTileRows = 64;           // H pass: 32 row pairs (rows l and l+32 are one lane's pair)
constexpr int kTileBytesMax = 1024;     // H pass: tile width in bytes (256 RGBA / 256 RGB / 1024 gray pixels)

// A thread marches over a segment of kSegPixels pixels (+ 2R of warm-up).  Short segments for small radii double
// the number of warps per tile (occupancy) at a warm-up overhead of 2R / kSegPixels.
template <int C, int R> struct HCfg {
    static constexpr int kSegPixels = (R <= 4) ? 64 : 128;
    static constexpr int kTilePixels = (C == 1) ? 1024 : 256;
    static constexpr int kSegs = kTilePixels / kSegPixels;
    static constexpr int kThreads = 32 * C * kSegs;
};

__device__ __forceinline__ uint64_t to_float_pair(uint32_t lo_byte, uint32_t hi_byte) {
    // bytes (already zero-extended) -> floats: OR into the mantissa of 2^23, subtract 2^23
    return add_rn_x2(pack_f2(lo_byte | 0x4B000000u, hi_byte | 0x4B000000u), splat_f2(-8388608.0f));
}
__device__ __forceinline__ uint64_t round_pair(uint64_t acc) {
    return add_rz_x2(add_rn_x2(acc, splat_f2(0.5f)), splat_f2(8388608.0f));
}

// ------------------------------------------------------------------------------------------------
// H pass.  Image rows [ty0, ty1) of every image of the chunk -> scratch image `tmp` (same pitch).
// ------------------------------------------------------------------------------------------------
template <int R, int C, bool kVec>
__global__ void __launch_bounds__(HCfg<C, R>::kThreads, 1)
gip_gauss_h(const __grid_constant__ Job job, uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1,
            int64_t img0, int tiles_x, int tiles_y, int in_pitch, int out_pitch, int64_t tpitch) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int R2 = 2 * R + 1;
    constexpr int TW = HCfg<C, R>::kTilePixels;
    constexpr int kSegPixels = HCfg<C, R>::kSegPixels;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t W = job.width, pitch = job.src.pitch;
    unsigned t = blockIdx.x;
    const int tx = (int)(t % (unsigned)tiles_x); t /= (unsigned)tiles_x;
    const int ty = (int)(t % (unsigned)tiles_y);
    const int64_t img = img0 + t / (unsigned)tiles_y;
    const int64_t row0 = ty0 + (int64_t)ty * kTileRows;
    const int nrows = (int)((ty1 - row0) < kTileRows ? (ty1 - row0) : kTileRows);
    const int64_t gx0 = (int64_t)tx * TW;                       // first output pixel of the tile
    const int64_t b0 = (gx0 - R) * C;                           // image-row byte position of tile byte 0
    const int tile_bytes = (TW + 2 * R) * C;
    const int skew = kVec ? (int)(((b0 % 4) + 4) % 4) : 0;      // keeps 4-byte chunks aligned on both sides
    uint8_t* in_tile = smem;
    uint8_t* out_tile = smem + (size_t)kTileRows * in_pitch;

    // ---- stage: bytes [lo, hi) of each row are real image bytes, the rest is clamp-to-edge replication
    const int64_t lo = b0 < 0 ? 0 : b0;
    int64_t hi = b0 + tile_bytes; if (hi > pitch) hi = pitch;
    if (kVec) {
        // 4-byte cp.async: the odd-word row pitch that makes the byte loads conflict-free rules out 16-byte chunks
        const int64_t cs = lo & ~int64_t(3);
        int64_t ce = (hi + 3) & ~int64_t(3); if (ce > pitch) ce = pitch;
        const int nchunk = ce > cs ? (int)((ce - cs) >> 2) : 0;
        const uint32_t in_s = smem_addr(in_tile);
        for (int rr = warp; rr < nrows; rr += HCfg<C, R>::kThreads / 32) {
            const uint8_t* src = job.src.row(row0 + rr, img) + cs;
            const uint32_t dst = in_s + (uint32_t)(rr * in_pitch + skew + (int)(cs - b0));
            for (int ci = lane; ci < nchunk; ci += 32) cp_async4(dst + 4 * ci, src + 4 * ci);
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else {               // rows that are not 4-byte aligned: a warp per row, plain byte copies
        const int n = hi > lo ? (int)(hi - lo) : 0;
        for (int rr = warp; rr < nrows; rr += HCfg<C, R>::kThreads / 32) {
            const uint8_t* src = job.src.row(row0 + rr, img) + lo;
            uint8_t* dst = in_tile + rr * in_pitch + (int)(lo - b0);
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
    __syncthreads();
    {   // replicate the edge pixel into the halo outside the image (left of pixel 0, right of pixel W-1)
        const int nl = b0 < 0 ? (int)(-b0) : 0;
        const int nr = (b0 + tile_bytes > pitch) ? (int)(b0 + tile_bytes - pitch) : 0;
        const int first_r = tile_bytes - nr;
        if (nl + nr > 0) {
            for (int rr = warp; rr < nrows; rr += HCfg<C, R>::kThreads / 32) {
                uint8_t* rowp = in_tile + rr * in_pitch + skew;
                for (int i = lane; i < nl; i += 32) rowp[i] = rowp[nl + (i % C)];                 // b0 is a multiple of C
                for (int k = lane; k < nr; k += 32) rowp[first_r + k] = rowp[first_r - C + (k % C)];
            }
        }
    }
    __syncthreads();

    // ---- march: this thread = (row pair, channel, segment)
    {
        const int ch = warp % C, seg = warp / C;
        const int ra = lane, rb = lane + 32;
        const bool has_a = ra < nrows, has_b = rb < nrows;
        const uint8_t* pa = in_tile + (has_a ? ra : 0) * in_pitch + skew + seg * kSegPixels * C + ch;
        const uint8_t* pb = in_tile + (has_b ? rb : 0) * in_pitch + skew + seg * kSegPixels * C + ch;
        uint8_t* qa = out_tile + ra * out_pitch + seg * kSegPixels * C + ch;
        uint8_t* qb = out_tile + rb * out_pitch + seg * kSegPixels * C + ch;
        uint64_t acc[R2];
#pragma unroll
        for (int i = 0; i < R2; i++) acc[i] = 0;
        constexpr int kSteps = kSegPixels + 2 * R;       // input pixels seg*128 - R ... seg*128 + 127 + R (tile-relative +R)
        // One step = input pixel s; u = s mod R2 is a compile-time slot after unrolling, so every accumulator
        // index below is a fixed register.  EMIT: output pixel s - 2R is complete after this step.
#define GIP_H_STEP(S, U_, EMIT)                                                                          \
        {                                                                                                \
            const uint64_t v = to_float_pair(pa[(S) * C], pb[(S) * C]);                                  \
            acc[U_] = mul_rn_x2(v, splat_f2(job.weights[0]));                                            \
            _Pragma("unroll")                                                                            \
            for (int k = 1; k < R2; k++)                                                                 \
                acc[((U_) - k + R2) % R2] = fma_rn_x2(v, splat_f2(job.weights[k]), acc[((U_) - k + R2) % R2]); \
            if (EMIT) {                                                                                  \
                const uint64_t z = round_pair(acc[((U_) + 1) % R2]);                                     \
                qa[((S) - 2 * R) * C] = (uint8_t)lo_f2(z);                                               \
                qb[((S) - 2 * R) * C] = (uint8_t)hi_f2(z);                                               \
            }                                                                                            \
        }
        constexpr int kFull = kSteps / R2;               // whole blocks of R2 steps
#pragma unroll
        for (int u = 0; u < R2; u++) GIP_H_STEP(u, u, u >= 2 * R)      // block 0: the first 2R steps produce nothing
        for (int b = 1; b < kFull; b++) {
            const int s0 = b * R2;
#pragma unroll
            for (int u = 0; u < R2; u++) GIP_H_STEP(s0 + u, u, true)
        }
#pragma unroll
        for (int u = 0; u < kSteps - kFull * R2; u++) GIP_H_STEP(kFull * R2 + u, u, true)
#undef GIP_H_STEP
    }
    __syncthreads();

    // ---- copy the output tile to the scratch image (its pitch `tpitch` is a multiple of 16: word stores)
    const int64_t ob0 = gx0 * C;
    int64_t out_bytes = (int64_t)TW * C; if (ob0 + out_bytes > pitch) out_bytes = pitch - ob0;
    uint8_t* tbase = tmp + ((img - img0) * (ty1 - ty0) + (row0 - ty0)) * tpitch + ob0;
    const int nv = (int)((out_bytes + 3) >> 2);
    for (int rr = warp; rr < nrows; rr += HCfg<C, R>::kThreads / 32) {      // a warp per row: coalesced, no index division
        const uint32_t* srow = reinterpret_cast<const uint32_t*>(out_tile + rr * out_pitch);
        uint32_t* grow = reinterpret_cast<uint32_t*>(tbase + (int64_t)rr * tpitch);
        for (int ci = lane; ci < nv; ci += 32) grow[ci] = srow[ci];
    }
}

// ------------------------------------------------------------------------------------------------
// V pass.  Scratch rows -> output rows [band_y0, band_y1).  One thread per 4-byte column group.
// ------------------------------------------------------------------------------------------------
template <int R, bool kAlignedOut>
__global__ void __launch_bounds__(128)
gip_gauss_v(const __grid_constant__ Job job, const uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1,
            int64_t img0, int nbands, int band_rows, int words, int64_t tpitch) {
    constexpr int R2 = 2 * R + 1;
    const int64_t pitch = job.src.pitch;
    const int wi = blockIdx.x * 128 + threadIdx.x;
    if (wi >= words) return;
    const int band = blockIdx.y % nbands;
    const int64_t li = blockIdx.y / nbands;                       // image index inside the chunk
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * band_rows;
    const int64_t Y1 = (Y0 + band_rows < job.src.band_y1) ? Y0 + band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;
    const uint8_t* timg = tmp + li * (ty1 - ty0) * tpitch + 4 * (int64_t)wi;
    const int nbytes = (pitch - 4 * (int64_t)wi >= 4) ? 4 : (int)(pitch - 4 * (int64_t)wi);
    uint8_t* optr = job.out + (img0 + li) * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + 4 * (int64_t)wi;
    const int64_t H = job.height;

    uint64_t acc0[R2], acc1[R2];
#pragma unroll
    for (int i = 0; i < R2; i++) { acc0[i] = 0; acc1[i] = 0; }
    const int nsteps = (int)(Y1 - Y0) + 2 * R;                    // input rows Y0-R .. Y1-1+R (clamped to the image)
    constexpr int M = (8 + R2 - 1) / R2;                          // blocks per super-block: >= 8 rows of loads in flight
    constexpr int U = M * R2;
    // input row of step s is clamp(Y0 - R + s, 0, H - 1): clamp the step index instead (32-bit), one IMAD.WIDE per load
    const int s_lo = (Y0 - R < 0) ? (int)(R - Y0) : 0;
    const int s_hi_img = (int)(H - 1 - (Y0 - R));
    const int s_hi = s_hi_img < nsteps - 1 ? s_hi_img : nsteps - 1;
    const uint8_t* tbase0 = timg + (Y0 - R - ty0) * tpitch;         // row of step 0 (may lie above the image: never dereferenced unclamped)
    auto load = [&](int s) {
        const int sc = s < s_lo ? s_lo : (s > s_hi ? s_hi : s);
        return __ldg(reinterpret_cast<const uint32_t*>(tbase0 + (int64_t)sc * tpitch));
    };
#define GIP_V_STEP(WORD, U_, EMIT)                                                                       \
    {                                                                                                    \
        const uint32_t w_ = (WORD);                                                                      \
        const uint64_t v0 = to_float_pair(w_ & 0xFFu, (w_ >> 8) & 0xFFu);                                \
        const uint64_t v1 = to_float_pair((w_ >> 16) & 0xFFu, w_ >> 24);                                 \
        acc0[U_] = mul_rn_x2(v0, splat_f2(job.weights[0]));                                              \
        acc1[U_] = mul_rn_x2(v1, splat_f2(job.weights[0]));                                              \
        _Pragma("unroll")                                                                                \
        for (int k = 1; k < R2; k++) {                                                                   \
            const uint64_t wk = splat_f2(job.weights[k]);                                                \
            acc0[((U_) - k + R2) % R2] = fma_rn_x2(v0, wk, acc0[((U_) - k + R2) % R2]);                  \
            acc1[((U_) - k + R2) % R2] = fma_rn_x2(v1, wk, acc1[((U_) - k + R2) % R2]);                  \
        }                                                                                                \
        if (EMIT) {                                                                                      \
            const uint64_t z0 = round_pair(acc0[((U_) + 1) % R2]), z1 = round_pair(acc1[((U_) + 1) % R2]); \
            const uint32_t t0 = __byte_perm(lo_f2(z0), hi_f2(z0), 0x4040), t1 = __byte_perm(lo_f2(z1), hi_f2(z1), 0x4040); \
            const uint32_t ow_ = __byte_perm(t0, t1, 0x5410);                                            \
            if (kAlignedOut) stg32_stream(optr, ow_);                                                    \
            else for (int b_ = 0; b_ < nbytes; b_++) optr[b_] = (uint8_t)(ow_ >> (8 * b_));              \
            optr += pitch;                                                                               \
        }                                                                                                \
    }
    uint32_t cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; u++) nxt[u] = load(u);
    for (int s0 = 0; s0 < nsteps; s0 += U) {
#pragma unroll
        for (int u = 0; u < U; u++) { cur[u] = nxt[u]; nxt[u] = load(s0 + U + u); }
        if (s0 >= 2 * R && s0 + U <= nsteps) {
#pragma unroll
            for (int u = 0; u < U; u++) GIP_V_STEP(cur[u], u % R2, true)
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
                if (s0 + u < nsteps) GIP_V_STEP(cur[u], u % R2, s0 + u >= 2 * R)
        }
    }
#undef GIP_V_STEP
}

template <int R, int C>
cudaError_t launch_h(const Job& job, uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0, int64_t nimg,
                     bool vec, cudaStream_t stream) {
    constexpr int TW = HCfg<C, R>::kTilePixels;
    const int tiles_x = (int)((job.width + TW - 1) / TW);
    const int tiles_y = (int)((ty1 - ty0 + kTileRows - 1) / kTileRows);
    int in_pitch = ((TW + 2 * R) * C + 32 + 3) & ~3;            // + skew and chunk-rounding room, whole words
    if (((in_pitch >> 2) & 1) == 0) in_pitch += 4;                // odd number of words: conflict-free across rows
    int out_pitch = (TW * C + 15) & ~15;
    out_pitch += 4;                                               // TW*C/4 is even: +1 word makes the pitch odd
    const size_t smem = (size_t)kTileRows * (in_pitch + out_pitch);
    const int64_t blocks = (int64_t)tiles_x * tiles_y * nimg;
    if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
    cudaError_t e;
    if (vec) {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(gip_gauss_h<R, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; set = true; }
        gip_gauss_h<R, C, true><<<(unsigned)blocks, HCfg<C, R>::kThreads, smem, stream>>>(job, tmp, ty0, ty1, img0, tiles_x, tiles_y, in_pitch, out_pitch, tpitch);
    } else {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(gip_gauss_h<R, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; set = true; }
        gip_gauss_h<R, C, false><<<(unsigned)blocks, HCfg<C, R>::kThreads, smem, stream>>>(job, tmp, ty0, ty1, img0, tiles_x, tiles_y, in_pitch, out_pitch, tpitch);
    }
    count_launch();
    return cudaGetLastError();
}

template <int R>
cudaError_t launch_v(const Job& job, const uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0,
                     int64_t nimg, cudaStream_t stream) {
    const int words = (int)((job.src.pitch + 3) / 4);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    const int64_t col_blocks = (words + 127) / 128;
    // enough bands to fill the machine (each band re-reads 2R scratch rows)
    int64_t want = ((int64_t)num_sms() * 16 + col_blocks * nimg - 1) / (col_blocks * nimg);
    int64_t max_bands = rows / (4 * (2 * R + 1)); if (max_bands < 1) max_bands = 1;
    if (want > max_bands) want = max_bands;
    if (want < 1) want = 1;
    const int nbands = (int)want;
    const int band_rows = (int)((rows + nbands - 1) / nbands);
    if (nbands * nimg > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)col_blocks, (unsigned)(nbands * nimg));
    const bool aligned_out = (job.src.pitch % 4 == 0) && (job.src.image_stride % 4 == 0) && ((uintptr_t)job.out % 4 == 0);
    if (aligned_out) gip_gauss_v<R, true><<<grid, 128, 0, stream>>>(job, tmp, ty0, ty1, img0, nbands, band_rows, words, tpitch);
    else             gip_gauss_v<R, false><<<grid, 128, 0, stream>>>(job, tmp, ty0, ty1, img0, nbands, band_rows, words, tpitch);
    count_launch();
    return cudaGetLastError();
}

template <int R>
cudaError_t run_radius(const Job& job, cudaStream_t stream) {
    const int C = job.channels;
    const int64_t pitch = job.src.pitch;
    const int64_t ty0 = clamp64(job.src.band_y0 - R, 0, job.height);
    const int64_t ty1 = clamp64(job.src.band_y1 + R, 0, job.height);
    const int64_t trows = ty1 - ty0;
    const bool vec = (pitch % 4 == 0) && (job.src.image_stride % 4 == 0) && ((uintptr_t)job.src.band % 4 == 0) &&
                     (!job.src.above || (uintptr_t)job.src.above % 4 == 0) &&
                     (!job.src.below || (uintptr_t)job.src.below % 4 == 0);
    // scratch: whole images of the batch, at most ~1 GiB at a time (and at most 8192 images: grid.y)
    const int64_t tpitch = (pitch + 15) & ~int64_t(15);      // scratch rows are 16-byte aligned whatever the image pitch is
    int64_t chunk = (int64_t(1) << 30) / (trows * tpitch);
    if (chunk < 1) chunk = 1;
    if (chunk > job.batch) chunk = job.batch;
    if (chunk > 2048) chunk = 2048;
    uint8_t* tmp = nullptr;
    cudaError_t err = cudaMallocAsync((void**)&tmp, (size_t)(chunk * trows * tpitch), stream);
    if (err != cudaSuccess) return err;
    for (int64_t img0 = 0; img0 < job.batch && err == cudaSuccess; img0 += chunk) {
        const int64_t n = (job.batch - img0 < chunk) ? job.batch - img0 : chunk;
        if (C == 4)      err = launch_h<R, 4>(job, tmp, tpitch, ty0, ty1, img0, n, vec, stream);
        else if (C == 3) err = launch_h<R, 3>(job, tmp, tpitch, ty0, ty1, img0, n, vec, stream);
        else             err = launch_h<R, 1>(job, tmp, tpitch, ty0, ty1, img0, n, vec, stream);
        if (err == cudaSuccess) err = launch_v<R>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
    }
    cudaError_t ferr = cudaFreeAsync(tmp, stream);
    return err != cudaSuccess ? err : ferr;
}

}  // namespace
}  // namespace gip
