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
//   H pass  (gip_gauss_h)  a CTA stages a 32-row x 256-pixel tile (+ r pixels of halo, clamp-to-edge)
//           in shared memory with cp.async.  A thread owns one channel of one row and marches along
//           x over TWO segments at once: they are the two lanes of FFMA2, so one issue slot does two
//           taps.  Lanes of a warp are the 32 rows (odd shared-memory pitch: conflict-free byte loads).
//           Rounded bytes go to an output tile in shared memory and leave with coalesced 16-byte stores.
//   V pass  (gip_gauss_v)  a thread owns a 4-byte column group and marches down a band of rows straight
//           from global memory (coalesced 32-bit loads); adjacent bytes are the FFMA2 lanes.
// Rounding: (uchar)(sum + 0.5f) (:102, :142) == low mantissa byte of RZ((sum + 0.5f) + 2^23).
// u8 -> float: I2FP.F32.U32 on the zero-extended byte (integer pipe; exact).
// The algorithm is FP32-issue bound, not HBM bound, from radius 3 up (4r+2 FMAs per byte); DESIGN.md
// carries the instruction roofline next to the HBM one.
#pragma once
#include <atomic>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kMaxFastRadius = 31;
constexpr int kMaxRotateRadius = 15;
constexpr int kWChunk = 8;              // shift formulation: head / tail steps per statically shaped chunk    // above: the "shift" formulation (see gip_gauss_h / gip_gauss_wv)
constexpr int kTileRows = 32;           // H pass: lane l owns tile row l; the FFMA2 lanes are two segments of that row
constexpr int kTileBytesMax = 1024;     // H pass: tile width in bytes (256 RGBA / 256 RGB / 1024 gray pixels)

// A thread marches over a segment of kSegPixels pixels (+ 2R of warm-up).  Short segments for small radii double
// the number of warps per tile (occupancy) at a warm-up overhead of 2R / kSegPixels.
template <int C, int R> struct HCfg {
    static constexpr int kSegPixels = (R <= 4) ? 64 : 128;
    static constexpr int kTilePixels = (C == 1) ? 1024 : 256;
    static constexpr int kSegs = kTilePixels / kSegPixels;
    static constexpr int kThreads = 32 * C * kSegs / 2;     // a thread marches two segments at once
};

// Shift formulation (radius 16..31), defined in gauss_shift_impl.cuh; the translation units of radius <= 15 only see
// these declarations (the kShift branches are never instantiated there).
template <int R, int C>
__device__ __forceinline__ void h_shift_march(const Job& job, const uint8_t* pa, const uint8_t* pb, uint8_t* qa, uint8_t* qb);
template <int R>
cudaError_t launch_wv(const Job& job, const uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0,
                      int64_t nimg, cudaStream_t stream);

__device__ __forceinline__ uint64_t to_float_pair(uint32_t lo_byte, uint32_t hi_byte) {
    // bytes (already zero-extended) -> floats, exact
    return pack_f2(u2f_bits(lo_byte), u2f_bits(hi_byte));
}
__device__ __forceinline__ uint64_t round_pair(uint64_t acc) {
    return add_rz_x2(add_rn_x2(acc, splat_f2(0.5f)), splat_f2(8388608.0f));
}

// ------------------------------------------------------------------------------------------------
// H pass.  Image rows [ty0, ty1) of every image of the chunk -> scratch image `tmp` (same pitch).
// ------------------------------------------------------------------------------------------------
// (shift formulation: at least 65536 / (threads x 168) blocks per SM, so that ptxas keeps the two accumulator sets of
// gauss_shift_impl.cuh in one set of registers instead of 237)
template <int R, int C, bool kShift>
__global__ void __launch_bounds__(HCfg<C, R>::kThreads, kShift ? 65536 / (HCfg<C, R>::kThreads * 168) : 1)
gip_gauss_h(const __grid_constant__ Job job, uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1,
            int64_t img0, int tiles_x, int tiles_y, int in_pitch, int out_pitch, int64_t tpitch) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int R2 = 2 * R + 1;
    constexpr int TW = HCfg<C, R>::kTilePixels;
    constexpr int kSegPixels = HCfg<C, R>::kSegPixels;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pitch = job.src.pitch;
    unsigned t = blockIdx.x;
    const int tx = (int)(t % (unsigned)tiles_x); t /= (unsigned)tiles_x;
    const int ty = (int)(t % (unsigned)tiles_y);
    const int64_t img = img0 + t / (unsigned)tiles_y;
    const int64_t row0 = ty0 + (int64_t)ty * kTileRows;
    const int nrows = (int)((ty1 - row0) < kTileRows ? (ty1 - row0) : kTileRows);
    const int64_t gx0 = (int64_t)tx * TW;                       // first output pixel of the tile
    const int64_t b0 = (gx0 - R) * C;                           // image-row byte position of tile byte 0
    const int tile_bytes = (TW + 2 * R) * C;
    uint8_t* in_tile = smem;
    uint8_t* out_tile = smem + (size_t)kTileRows * in_pitch;
    constexpr int kWarps = HCfg<C, R>::kThreads / 32;
    // Tile byte 0 of row rr sits at in_tile + rr*in_pitch + skew(rr), skew(rr) = (global address of that byte) mod 4:
    // whatever the image pitch and base alignment are, 4-byte chunks are then aligned on both sides of cp.async.
    auto row_skew = [&](const uint8_t* rowp) { return (int)(((intptr_t)rowp + b0) & 3); };
    // global pointer of tile row rr: plain pitch arithmetic when the whole tile lies inside the caller's band
    const bool in_band = row0 >= job.src.band_y0 && row0 + nrows <= job.src.band_y1;
    const uint8_t* const band_row0 = job.src.band + img * job.src.image_stride + (row0 - job.src.band_y0) * pitch;
    auto tile_row = [&](int rr) { return in_band ? band_row0 + (int64_t)rr * pitch : job.src.row(row0 + rr, img); };

    // ---- stage: bytes [lo, hi) of each row are real image bytes, the rest is clamp-to-edge replication
    const int64_t lo = b0 < 0 ? 0 : b0;
    int64_t hi = b0 + tile_bytes; if (hi > pitch) hi = pitch;
    {
        const int n = hi > lo ? (int)(hi - lo) : 0;
        const uint32_t in_s = smem_addr(in_tile);
        constexpr int kRowsPerWarp = (kTileRows + kWarps - 1) / kWarps;
        constexpr int kChunkIters = ((TW + 2 * R) * C / 4 + 31) / 32;     // 4-byte chunks of a tile row per lane
        // Whole 4-byte chunks go through cp.async.  The up to 3 bytes before the first aligned chunk and after the
        // last one are plain loads by lanes 0-2 and 4-6: all rows' loads are issued before any is stored, so
        // their latency is paid once, under the asynchronous copies.
        uint8_t edge_byte[kRowsPerWarp];
        int edge_off[kRowsPerWarp];
        // per row: skew, bytes before the first aligned chunk, whole chunks, this lane's edge byte (or -1)
        auto row_layout = [&](const uint8_t* rowp, int& skew, int& head, int& nchunk, int& e) {
            skew = row_skew(rowp);
            head = (int)((-(intptr_t)(rowp + lo)) & 3); if (head > n) head = n;
            nchunk = (n - head) >> 2;
            const int tail0 = head + 4 * nchunk;
            e = lane < 4 ? (lane < head ? lane : -1) : (lane - 4 < n - tail0 && lane < 8 ? tail0 + lane - 4 : -1);
        };
        auto stage_row = [&](int i, int rr, const uint8_t* rowp, int skew, int head, int nchunk, int e) {
            const uint8_t* g = rowp + lo;
            const int off = rr * in_pitch + skew + (int)(lo - b0);
            if (e >= 0) { edge_byte[i] = g[e]; edge_off[i] = off + e; }
            const uint8_t* gp = g + head + 4 * lane;
            const uint32_t dp = in_s + (uint32_t)(off + head + 4 * lane);
            const int mine = nchunk - lane;                        // this lane copies chunks lane, lane+32, ...
#pragma unroll
            for (int it = 0; it < kChunkIters; it++)
                if (mine > 32 * it) cp_async4(dp + 128 * it, gp + 128 * it);
        };
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; i++) edge_off[i] = -1;
        if (in_band && (pitch & 3) == 0) {                         // every row of the tile has the same layout
            int skew, head, nchunk, e;
            row_layout(band_row0, skew, head, nchunk, e);
#pragma unroll
            for (int i = 0; i < kRowsPerWarp; i++) {
                const int rr = warp + i * kWarps;
                if (rr < nrows) stage_row(i, rr, band_row0 + (int64_t)rr * pitch, skew, head, nchunk, e);
            }
        } else {
#pragma unroll
            for (int i = 0; i < kRowsPerWarp; i++) {
                const int rr = warp + i * kWarps;
                if (rr < nrows) {
                    const uint8_t* rowp = tile_row(rr);
                    int skew, head, nchunk, e;
                    row_layout(rowp, skew, head, nchunk, e);
                    stage_row(i, rr, rowp, skew, head, nchunk, e);
                }
            }
        }
        cp_async_commit();
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; i++)
            if (edge_off[i] >= 0) in_tile[edge_off[i]] = edge_byte[i];
        cp_async_wait<0>();
    }
    __syncthreads();
    {   // replicate the edge pixel into the halo outside the image (left of pixel 0, right of pixel W-1)
        const int nl = b0 < 0 ? (int)(-b0) : 0;
        const int nr = (b0 + tile_bytes > pitch) ? (int)(b0 + tile_bytes - pitch) : 0;
        const int first_r = tile_bytes - nr;
        if (nl + nr > 0) {
            for (int rr = warp; rr < nrows; rr += kWarps) {
                uint8_t* rowp = in_tile + rr * in_pitch + row_skew(tile_row(rr));
                for (int i = lane; i < nl; i += 32) rowp[i] = rowp[nl + (i % C)];                 // b0 is a multiple of C
                for (int k = lane; k < nr; k += 32) rowp[first_r + k] = rowp[first_r - C + (k % C)];
            }
        }
    }
    __syncthreads();

    // ---- march: this thread = (row, channel, segment pair)
    {
        constexpr int kHalf = HCfg<C, R>::kSegs / 2;           // segment seg and segment seg + kHalf are the FFMA2 lanes
        const int ch = warp % C, seg = warp / C;
        const int rs = lane < nrows ? lane : 0;                // rows past the tile's end recompute row 0 into rows nobody copies out
        const uint8_t* pa = in_tile + rs * in_pitch + row_skew(tile_row(rs)) + seg * kSegPixels * C + ch;
        const uint8_t* pb = pa + kHalf * kSegPixels * C;
        uint8_t* qa = out_tile + lane * out_pitch + seg * kSegPixels * C + ch;
        uint8_t* qb = qa + kHalf * kSegPixels * C;
        uint64_t acc[R2];
#pragma unroll
        for (int i = 0; i < R2; i++) acc[i] = 0;
        if constexpr (kShift) {
            h_shift_march<R, C>(job, pa, pb, qa, qb);            // gauss_shift_impl.cuh (radius 16..31)
        } else {
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
    }
    __syncthreads();

    // ---- copy the output tile to the scratch image: 16-byte stores (the tile's x origin and the scratch pitch
    // `tpitch` are multiples of 16; a ragged last group spills into the scratch row's padding)
    const int64_t ob0 = gx0 * C;
    int64_t out_bytes = (int64_t)TW * C; if (ob0 + out_bytes > pitch) out_bytes = pitch - ob0;
    uint8_t* tbase = tmp + ((img - img0) * (ty1 - ty0) + (row0 - ty0)) * tpitch + ob0;
    const int nv = (int)((out_bytes + 15) >> 4);
    const uint32_t out_s = smem_addr(out_tile);
    for (int rr = warp; rr < nrows; rr += kWarps) {              // a warp per row: coalesced, no index division
        const uint32_t sp = out_s + (uint32_t)(rr * out_pitch + 16 * lane);
        uint8_t* gp = tbase + (int64_t)rr * tpitch + 16 * lane;
        constexpr int kGroupIters = (TW * C / 16 + 31) / 32;
#pragma unroll
        for (int it = 0; it < kGroupIters; it++) {
            if (nv - lane > 32 * it) {
                uint4 v;
                v.x = lds32(sp + 512 * it); v.y = lds32(sp + 512 * it + 4); v.z = lds32(sp + 512 * it + 8); v.w = lds32(sp + 512 * it + 12);
                *reinterpret_cast<uint4*>(gp + 512 * it) = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// V pass.  Scratch rows -> output rows [band_y0, band_y1).  One thread per 4-byte column group.
// A band is walked in blocks of U = M*R2 input rows (static accumulator slots inside a block):
//   block 0        the first 2R rows only fill the accumulators (static emit pattern),
//   steady blocks  every row emits; the next block's rows are prefetched, without clamping when they
//                  all lie inside the image,
//   tail           fewer than U rows, guarded.  The host sizes bands to k*U - 2R rows, so that only the
//                  last band of an image has a tail.
// kAlignedOut = false: output rows are not 4-byte aligned (odd pitch).  A lane takes the last bytes of its
// left neighbour's word by shuffle and stores the aligned word that straddles both; the first lane of a
// warp, the last one and the partial last column store their leftover bytes one by one.
// ------------------------------------------------------------------------------------------------
template <int R> struct VCfg {
    static constexpr int R2 = 2 * R + 1;
    static constexpr int M = (8 + R2 - 1) / R2;          // blocks per super-block: >= 8 rows of loads in flight
    static constexpr int U = M * R2;
};

template <int R, bool kAlignedOut>
__global__ void __launch_bounds__(128, (R <= 4 && kAlignedOut) ? 5 : 1)
gip_gauss_v(const __grid_constant__ Job job, const uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1,
            int64_t img0, int nbands, int band_rows, int words, int64_t tpitch) {
    constexpr int R2 = VCfg<R>::R2;
    constexpr int U = VCfg<R>::U;
    const int64_t pitch = job.src.pitch;
    const int lane = threadIdx.x & 31;
    // Column of this thread.  Aligned output: 128 consecutive columns per block.  Unaligned output: consecutive warps
    // overlap by one column (lane 0 recomputes the previous warp's last column and never stores), so that every
    // storing lane finds its left neighbour's word in its own warp; a warp then covers 31 columns.
    const int wi_raw = kAlignedOut ? blockIdx.x * 128 + threadIdx.x
                                   : (blockIdx.x * 4 + (threadIdx.x >> 5)) * 31 + lane - 1;
    const bool live = wi_raw >= 0 && wi_raw < words;
    if (kAlignedOut && !live) return;                             // (the unaligned variant keeps whole warps for its shuffles)
    const int wi = wi_raw < 0 ? 0 : (wi_raw < words ? wi_raw : words - 1);   // spare lanes shadow an edge column and never store
    const int band = blockIdx.y % nbands;
    const int64_t li = blockIdx.y / nbands;                       // image index inside the chunk
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * band_rows;
    const int64_t Y1 = (Y0 + band_rows < job.src.band_y1) ? Y0 + band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;
    const uint8_t* timg = tmp + li * (ty1 - ty0) * tpitch + 4 * (int64_t)wi;
    const int nbytes = (pitch - 4 * (int64_t)wi >= 4) ? 4 : (int)(pitch - 4 * (int64_t)wi);
    uint8_t* optr = job.out + (img0 + li) * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + 4 * (int64_t)wi;
    const int64_t H = job.height;
    // Unaligned output only: which stores this thread does for each row misalignment `mis` (the same in every lane of
    // a row).  A full column with a left neighbour stores the aligned word [its address - mis, +4) = the neighbour's
    // last `mis` bytes + its own first 4 - mis.  Bytes nobody's word covers are stored one by one: the first bytes of
    // column 0, the last bytes of the last full column, the partial last column; those lanes sit in at most two
    // warps of a row, every other warp skips the byte path as a whole.
    unsigned mis = (unsigned)((uintptr_t)optr & 3);
    const unsigned mis_step = (unsigned)(pitch & 3);
    unsigned byte_mask = 0, word_mask = 0;        // nibble m of byte_mask: own bytes stored one by one when mis == m
    if (!kAlignedOut && live && lane > 0) {
        const bool full = nbytes == 4;
        const bool next_full = (wi_raw + 1 < words) && (pitch - 4 * (int64_t)(wi_raw + 1) >= 4);
        const unsigned all = (1u << nbytes) - 1;
        if (full) word_mask |= 1; else byte_mask |= all;
        for (int m = 1; m < 4; m++) {
            const unsigned head = (1u << (4 - m)) - 1;        // own bytes that share an aligned word with the left neighbour
            unsigned nib = 0;
            if (full && wi_raw > 0) word_mask |= 1u << m; else nib |= head & all;
            if (!next_full) nib |= all & ~head;              // nobody to the right takes this word's last bytes
            byte_mask |= nib << (4 * m);
        }
    }
    const bool edge_warp = !kAlignedOut && __any_sync(0xffffffffu, byte_mask != 0);

    uint64_t acc0[R2], acc1[R2];
#pragma unroll
    for (int i = 0; i < R2; i++) { acc0[i] = 0; acc1[i] = 0; }
    const int nsteps = (int)(Y1 - Y0) + 2 * R;                    // input rows Y0-R .. Y1-1+R (clamped to the image)
    // input row of step s is clamp(Y0 - R + s, 0, H - 1): clamp the step index instead (32-bit)
    const int s_lo = (Y0 - R < 0) ? (int)(R - Y0) : 0;
    const int s_hi_img = (int)(H - 1 - (Y0 - R));
    const int s_hi = s_hi_img < nsteps - 1 ? s_hi_img : nsteps - 1;
    const uint8_t* tbase0 = timg + (Y0 - R - ty0) * tpitch;         // row of step 0 (may lie above the image: never dereferenced unclamped)
    const unsigned tp32 = (unsigned)tpitch;
    auto load = [&](int s) {
        const int sc = s < s_lo ? s_lo : (s > s_hi ? s_hi : s);
        return __ldg(reinterpret_cast<const uint32_t*>(tbase0 + (int64_t)sc * tpitch));
    };
#define GIP_V_STEP(WORD, U_, EMIT)                                                                       \
    {                                                                                                    \
        const uint32_t w_ = (WORD);                                                                      \
        const uint64_t v0 = to_float_pair(__byte_perm(w_, 0u, 0x4440), __byte_perm(w_, 0u, 0x4441));      \
        const uint64_t v1 = to_float_pair(__byte_perm(w_, 0u, 0x4442), __byte_perm(w_, 0u, 0x4443));      \
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
            if (kAlignedOut) {                                                                           \
                stg32_stream(optr, ow_);                                                                 \
            } else {                                                                                     \
                const uint32_t left_ = __shfl_up_sync(0xffffffffu, ow_, 1);                              \
                if ((word_mask >> mis) & 1) stg32_stream(optr - mis, __funnelshift_l(left_, ow_, 8 * mis)); \
                if (edge_warp) {                                                                         \
                    const unsigned nib_ = (byte_mask >> (4 * mis)) & 15u;                                \
                    if (nib_ & 1) optr[0] = (uint8_t)ow_;                                                \
                    if (nib_ & 2) optr[1] = (uint8_t)(ow_ >> 8);                                         \
                    if (nib_ & 4) optr[2] = (uint8_t)(ow_ >> 16);                                        \
                    if (nib_ & 8) optr[3] = (uint8_t)(ow_ >> 24);                                        \
                }                                                                                        \
                mis = (mis + mis_step) & 3;                                                              \
            }                                                                                            \
            optr += pitch;                                                                               \
        }                                                                                                \
    }
    uint32_t cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; u++) nxt[u] = load(u);
    int s0 = 0;
    if (nsteps >= U) {
#pragma unroll
        for (int u = 0; u < U; u++) { cur[u] = nxt[u]; nxt[u] = load(U + u); }
#pragma unroll
        for (int u = 0; u < U; u++) GIP_V_STEP(cur[u], u % R2, u >= 2 * R)
        for (s0 = U; s0 + U <= nsteps; s0 += U) {
            if (s0 + U >= s_lo && s0 + 2 * U - 1 <= s_hi) {       // the whole prefetched block is inside the image
                const uint8_t* blk = tbase0 + (int64_t)(s0 + U) * tpitch;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    cur[u] = nxt[u];
                    nxt[u] = __ldg(reinterpret_cast<const uint32_t*>(blk + (uint64_t)((unsigned)u * tp32)));
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; u++) { cur[u] = nxt[u]; nxt[u] = load(s0 + U + u); }
            }
#pragma unroll
            for (int u = 0; u < U; u++) GIP_V_STEP(cur[u], u % R2, true)
        }
    }
    if (s0 < nsteps) {                                            // tail of the band (or a band shorter than one block)
#pragma unroll
        for (int u = 0; u < U; u++)
            if (s0 + u < nsteps) GIP_V_STEP(nxt[u], u % R2, s0 + u >= 2 * R)
    }
#undef GIP_V_STEP
}

template <int R, int C, bool kShift>
cudaError_t launch_h(const Job& job, uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0, int64_t nimg,
                     cudaStream_t stream) {
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
    static std::atomic<bool> attr_set[64];                        // per instantiation and per device: the opt-in is a per-device attribute
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_set[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(gip_gauss_h<R, C, kShift>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev].store(true, std::memory_order_release);
    }
    gip_gauss_h<R, C, kShift><<<(unsigned)blocks, HCfg<C, R>::kThreads, smem, stream>>>(job, tmp, ty0, ty1, img0, tiles_x, tiles_y,
                                                                             in_pitch, out_pitch, tpitch);
    count_launch();
    return cudaGetLastError();
}

template <int R>
cudaError_t launch_v(const Job& job, const uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0,
                     int64_t nimg, cudaStream_t stream) {
    constexpr int U = VCfg<R>::U;
    const int words = (int)((job.src.pitch + 3) / 4);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    const bool aligned_out = (job.src.pitch % 4 == 0) && (job.src.image_stride % 4 == 0) && ((uintptr_t)job.out % 4 == 0);
    const int64_t col_blocks = aligned_out ? (words + 127) / 128 : (words + 123) / 124;   // unaligned: 4 warps x 31 columns
    static std::atomic<int> per_sm_cache[64][2];                  // resident blocks per SM of the two variants, per device
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::atomic<int>* per_sm = per_sm_cache[dev];
    if (per_sm[aligned_out] == 0) {
        int n = 0;
        cudaError_t e = aligned_out ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gip_gauss_v<R, true>, 128, 0)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gip_gauss_v<R, false>, 128, 0);
        if (e != cudaSuccess) return e;
        per_sm[aligned_out] = n > 0 ? n : 1;
    }
    // Bands.  A block's time is proportional to its steps (band rows + 2R rows that only fill the accumulators), the
    // launch's to the number of waves of resident blocks: take the band count with the smallest waves x steps.
    // Bands are k*U - 2R rows so that every band but an image's last is whole blocks of U steps.
    const int64_t resident = (int64_t)num_sms() * per_sm[aligned_out];
    int64_t band_rows = rows, best_cost = -1;
    for (int64_t nb = 1; nb <= 512 && nb <= rows; nb++) {
        int64_t br = (rows + nb - 1) / nb;
        br = (br + 2 * R + U - 1) / U * U - 2 * R;
        while (br < 1) br += U;
        if (br > rows) br = rows;
        const int64_t n_actual = (rows + br - 1) / br;
        if (n_actual * nimg > 65535) continue;
        const int64_t blocks = col_blocks * n_actual * nimg;
        const int64_t waves = (blocks + resident - 1) / resident;
        const int64_t cost = waves * (br + 2 * R + 16);          // + a block's fixed start-up, in row steps
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; band_rows = br; }
    }
    const int64_t nbands = (rows + band_rows - 1) / band_rows;
    if (nbands * nimg > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)col_blocks, (unsigned)(nbands * nimg));
    if (aligned_out) gip_gauss_v<R, true><<<grid, 128, 0, stream>>>(job, tmp, ty0, ty1, img0, (int)nbands, (int)band_rows, words, tpitch);
    else             gip_gauss_v<R, false><<<grid, 128, 0, stream>>>(job, tmp, ty0, ty1, img0, (int)nbands, (int)band_rows, words, tpitch);
    count_launch();
    return cudaGetLastError();
}

template <int R, bool kShift>
cudaError_t run_radius(const Job& job, cudaStream_t stream) {
    static_assert(kShift || R <= kMaxRotateRadius, "the rotating-accumulator kernels are unrolled 2R+1 times: small radii only");
    const int C = job.channels;
    const int64_t pitch = job.src.pitch;
    const int64_t ty0 = clamp64(job.src.band_y0 - R, 0, job.height);
    const int64_t ty1 = clamp64(job.src.band_y1 + R, 0, job.height);
    const int64_t trows = ty1 - ty0;
    // scratch: whole images of the batch, at most ~1 GiB at a time (and at most 8192 images: grid.y)
    const int64_t tpitch = (pitch + 15) & ~int64_t(15);      // scratch rows are 16-byte aligned whatever the image pitch is
    int64_t chunk = (int64_t(1) << 30) / (trows * tpitch);
    if (chunk < 1) chunk = 1;
    if (chunk > job.batch) chunk = job.batch;
    if (chunk > 2048) chunk = 2048;
    uint8_t* tmp = nullptr;
    cudaError_t err = scratch_alloc((void**)&tmp, (size_t)(chunk * trows * tpitch), stream);
    if (err != cudaSuccess) return err;
    for (int64_t img0 = 0; img0 < job.batch && err == cudaSuccess; img0 += chunk) {
        const int64_t n = (job.batch - img0 < chunk) ? job.batch - img0 : chunk;
        if (C == 4)      err = launch_h<R, 4, kShift>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
        else if (C == 3) err = launch_h<R, 3, kShift>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
        else             err = launch_h<R, 1, kShift>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
        if (err == cudaSuccess) {
            if constexpr (kShift) err = launch_wv<R>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
            else err = launch_v<R>(job, tmp, tpitch, ty0, ty1, img0, n, stream);
        }
    }
    cudaError_t ferr = cudaFreeAsync(tmp, stream);
    return err != cudaSuccess ? err : ferr;
}

}  // namespace
}  // namespace gip
