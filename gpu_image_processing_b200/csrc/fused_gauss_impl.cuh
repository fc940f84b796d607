// fused_gauss_impl.cuh -- single-kernel separable Gaussian blur for sm_100a, small radii, rows at any byte alignment.
//
// Replaces gaussianBlur{Horizontal,Vertical}{Naive,Level2} AND the d_temp image between them
// (/root/reference/cuda_lib/src/image_filters.cu:64-144, :159-347, :759-761): the horizontally filtered, u8-rounded
// rows (the reference's intermediate, :102) live only in a shared-memory FIFO between the two passes, so every image
// byte is read from HBM once and written once.
//
// A CTA (512 threads, one per SM) owns a column strip of 2048 output bytes and a band of rows and marches down the
// band K = 16 rows per step.  It is warp-specialised:
//   8 producer warps   stage input rows with 16-byte cp.async (two steps ahead, double-buffered) and run the H pass.
//                      A warp filters two rows at once: the rows are the two lanes of the packed float32x2 operations
//                      (FFMA2), so one issue slot does two taps and nothing is ever re-packed.  Lane l owns 64 output
//                      bytes of both rows: it reads its 64 + 2*R*C input bytes with LDS.128, converts each byte once
//                      (PRMT + I2FP, off the FP32 pipe) and accumulates every output in the reference's tap order
//                      (:86-99: i = -R..R, one FMA per tap, first tap a multiply; scatter form: an input is tap t of
//                      output m - tC, so the 2R+1 FMAs it feeds are independent), rounds (:102) and writes the row
//                      into the FIFO (four steps of K rows; three in the any-alignment variant) with STS.128.
//   8 consumer warps   run the V pass one step behind.  A thread owns an 8-byte column group and keeps the 2R+1
//                      partial sums of its columns in registers: a new FIFO row v updates  acc[t] = fma(v, w[t], acc[t-1])
//                      for t = 2R..1, acc[0] = v * w[0]  (the rotation of the accumulators happens in the FMA's
//                      destination: no register moves, no unrolling by 2R+1), the finished acc[2R] is rounded (:142) and
//                      stored with STG.64; a warp writes 256 contiguous bytes.
// Shared-memory rows are padded by one 16-byte chunk per 128 bytes, which makes the lane-strided (64-byte) LDS.128 /
// STS.128 of the H pass conflict-free.
// Arithmetic is bit-identical to the reference: float32 weights from the host (image_filters.cu:25-39), one fmaf per
// tap in tap order, (uchar)(sum + 0.5f) == low mantissa byte of RZ(RN(sum + 0.5f) + 2^23).
// The FP32 pipe is the roofline: 2 * ((2R+1) + 2) packed-lane operations per byte.
//
// kAny = true (odd pitches, unaligned base pointers; the reference's 3239-pixel RGB README shape): the same kernel, plus
//   stage    the 16-byte chunks of GLOBAL memory that lie inside the row are copied (cp.async needs aligned addresses), so
//            a staged row arrives shifted by a = row address mod 16; the producer warp shifts it back in place (two
//            LDS.128, four funnel shifts, one STS.128 per chunk) and lanes fetch the up to 15 bytes at either end of the
//            row that belong to no whole chunk.  Everything after that is the aligned code.
//   store    output rows start at any address: they return through shared memory -- the consumers write them with the
//            aligned kernel's 8-byte store into a block of K + 2R rows behind the FIFO, then every consumer warp stores
//            whole rows with 16-byte stores at the aligned global addresses (flush_row_any, device_utils.cuh).  A strip is
//            1984 bytes (31 lanes) so that the block fits.  Output rows that are 8-byte aligned are stored directly.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kFK = 16;                     // rows per step
constexpr int kFProd = 8, kFCons = 8;       // producer / consumer warps
constexpr int kFThreads = 32 * (kFProd + kFCons);
constexpr int kFStrip = 2048;               // output bytes per strip (32 lanes x 64 bytes)
constexpr int kFLane = 64;
constexpr int kFRingPitch = kFStrip + kFStrip / 8;      // 2304: one pad chunk per 8 chunks

// byte offset of byte `idx` of a row stored in the padded layout
__host__ __device__ constexpr int padded_off(int idx) { return 16 * ((idx >> 4) + (idx >> 7)) + (idx & 15); }

template <int R, int C, bool kAny = false> struct FCfg {
    static constexpr int RC = R * C;
    static constexpr int kPad = (RC + 15) & ~15;                    // bytes staged to the left of the strip
    static constexpr int kDelta = kPad - RC;                        // a lane's first input byte inside its first chunk
    static constexpr int kRowBytes = kPad + kFStrip + kPad;         // staged bytes per row
    static constexpr int kRowChunks = kRowBytes / 16;
    // output bytes per strip.  kAny: 31 lanes' worth (1984 bytes), so that the block of output rows below fits next to the
    // staging buffers and the FIFO
    static constexpr int kUseful = kAny ? 31 * kFLane : kFStrip;
    // kAny: output rows return through shared memory (the consumers write them with the aligned kernel's 8-byte store,
    // then store whole rows to global memory at the aligned addresses, flush_row_any).  One consumer iteration emits at
    // most K + 2R rows (whole blocks of 2R+1 input rows).
    static constexpr int kOutRows = kFK + 2 * R;
    static constexpr int kOutPitch = kUseful;
    static constexpr int kOutBytes = kAny ? kOutRows * kOutPitch + 16 : 0;
    // kAny: only the chunks that the lanes with outputs inside the strip read are staged and shifted
    static constexpr int kNeedChunks = kAny ? (kDelta + kFLane * ((kUseful + kFLane - 1) / kFLane) + 2 * RC + 15) / 16 : kRowChunks;
    static constexpr int kWinBytes = 16 * kNeedChunks;              // staged (logical) bytes per row that are ever read or fixed up
    static constexpr int kRawChunks = kNeedChunks + (kAny ? 1 : 0); // a row shifted by up to 15 bytes spans one more chunk
    static constexpr int kStagePitch = 16 * (kRawChunks + (kRawChunks + 7) / 8);
    static constexpr int kLaneChunks = (kDelta + kFLane + 2 * RC + 15) / 16;   // chunks a lane reads per row
    static constexpr int kCopyIters = (kRawChunks + 31) / 32;
    // FIFO slots (steps of K rows).  The V pass consumes whole blocks of 2R+1 rows, so it may lag a step behind by up to 2R
    // rows and the rows it takes per iteration vary (r = 3: 14, 14, 21, ...): three slots keep producers and consumers in
    // lockstep (a producer's step p waits for the consumers' iteration p - 1), the fourth lets the producers run a step
    // further ahead and absorbs that variation.  The any-alignment variant has no room for it (block of output rows).
#ifndef GIP_FUSED_SLOTS
#define GIP_FUSED_SLOTS 4
#endif
    static constexpr int kSlots = kAny ? 3 : GIP_FUSED_SLOTS;
    static constexpr int kRingRows = kSlots * kFK;
    static constexpr size_t kSmem = (size_t)2 * kFK * kStagePitch + (size_t)kRingRows * kFRingPitch + kOutBytes;
    static_assert(kSmem <= 227 * 1024, "shared memory");
    static_assert(kLaneChunks <= 8, "lane chunk addressing assumes at most 8 chunks");
};

struct FusedTiling {
    int strips, bands, band_rows;
    int decoupled;      // 1: producers and consumers meet on named barriers per ring slot (full / empty), not __syncthreads
    int ret;            // kAny: 1 = output rows return through shared memory and leave with aligned 16-byte stores; 0 = output rows are
                        // 8-byte aligned (only the input needs the shift), the consumers store them themselves
};

// Named barriers 1..NS = FULL[slot], NS+1..2NS = EMPTY[slot] (0 is __syncthreads; NS = 3 or 4 slots, the any-alignment
// variant's consumer-only barriers 7 and 8 come after its 6).  Producers arrive on FULL[s % NS] when the rows
// of step s are in the FIFO and consumers wait there; consumers arrive on EMPTY[s % NS] when every row of step s has been
// consumed (end of their iteration s + 2: they take whole blocks of 2R+1 <= K+1 rows) and producers wait there before
// step s + 3 overwrites the slot.  Every barrier counts all threads of the CTA (arrivals + waiters).
__device__ __forceinline__ void fbar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kFThreads) : "memory"); }
__device__ __forceinline__ void fbar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kFThreads) : "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void stg64_stream(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// byte b (0..3) of word w -> float, exact (zero-extend with PRMT, I2FP on the integer side)
__device__ __forceinline__ uint32_t byte_to_float_bits(uint32_t w, int b) {
    return u2f_bits(__byte_perm(w, 0u, 0x4440 | b));
}
// The same conversion as ONE instruction on the XU pipe (I2F.U8 with a byte selector, 16 lanes/clk/SM): a quarter of the
// rate, but one issue slot instead of two and nothing on the ALU pipe.  Which bytes of a word take which route is a
// compile-time mask per pass (bit b = byte b of every word goes through the XU pipe).
#ifndef GIP_FUSED_XU_H
#define GIP_FUSED_XU_H 0
#endif
#ifndef GIP_FUSED_XU_V
#define GIP_FUSED_XU_V 15    // applied to the aligned kernel at R >= 3 only (see v_block)
#endif
template <int kMask>
__device__ __forceinline__ uint32_t byte_to_float_bits_m(uint32_t w, int b) {
    if ((kMask >> b) & 1) return __float_as_uint((float)((w >> (8 * b)) & 0xffu));
    return byte_to_float_bits(w, b);
}
__device__ __forceinline__ uint64_t round_pair_f(uint64_t acc) {
    return add_rz_x2(add_rn_x2(acc, splat_f2(0.5f)), splat_f2(8388608.0f));
}
// The truncation as two F2I on the XU pipe instead of one packed add on the FP32 pipe (the pipe the taps need): the low
// byte of the result is the same (values below 2^23).  Bit 0 = H pass, bit 1 = V pass.
#ifndef GIP_FUSED_H_SCATTER
#define GIP_FUSED_H_SCATTER 1
#endif
#ifndef GIP_FUSED_F2I
#define GIP_FUSED_F2I 0
#endif
__device__ __forceinline__ uint32_t f2i_rz_bits(uint32_t fbits) {
    uint32_t r;
    asm("{.reg .f32 t; mov.b32 t, %1; cvt.rzi.u32.f32 %0, t;}" : "=r"(r) : "r"(fbits));
    return r;
}
template <bool kXu>
__device__ __forceinline__ uint64_t round_pair_sel(uint64_t acc) {
    if (kXu) {
        const uint64_t t = add_rn_x2(acc, splat_f2(0.5f));
        return pack_f2(f2i_rz_bits(lo_f2(t)), f2i_rz_bits(hi_f2(t)));
    }
    return round_pair_f(acc);
}

// Shift a staged row left by a bytes (1..15), in place: logical chunk c = bytes [a, a + 16) of raw chunks c, c + 1.
// Chunk lane + 32 it sits at padded offset 16 (lane + lane / 8) + 576 it.
template <int NCH>
__device__ __forceinline__ void realign_row(uint32_t row_s, int a, int lane) {
    const int ws = a >> 2;
    const uint32_t bs = 8u * (uint32_t)(a & 3);
    uint32_t pA = row_s + 16u * (uint32_t)(lane + (lane >> 3));
    uint32_t pB = row_s + 16u * (uint32_t)((lane + 1) + ((lane + 1) >> 3));
    // (rolled: the any-alignment kernel is bound by instruction fetch, every copy of this body counts)
#pragma unroll 1
    for (int it = 0; it < (NCH + 31) / 32; it++, pA += 576u, pB += 576u) {
        const bool on = lane + 32 * it < NCH;
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        if (on) {
            const uint4 A = lds128(pA);
            const uint4 B = lds128(pB);
            if (ws == 0) {
                o0 = __funnelshift_r(A.x, A.y, bs); o1 = __funnelshift_r(A.y, A.z, bs); o2 = __funnelshift_r(A.z, A.w, bs); o3 = __funnelshift_r(A.w, B.x, bs);
            } else if (ws == 1) {
                o0 = __funnelshift_r(A.y, A.z, bs); o1 = __funnelshift_r(A.z, A.w, bs); o2 = __funnelshift_r(A.w, B.x, bs); o3 = __funnelshift_r(B.x, B.y, bs);
            } else if (ws == 2) {
                o0 = __funnelshift_r(A.z, A.w, bs); o1 = __funnelshift_r(A.w, B.x, bs); o2 = __funnelshift_r(B.x, B.y, bs); o3 = __funnelshift_r(B.y, B.z, bs);
            } else {
                o0 = __funnelshift_r(A.w, B.x, bs); o1 = __funnelshift_r(B.x, B.y, bs); o2 = __funnelshift_r(B.y, B.z, bs); o3 = __funnelshift_r(B.z, B.w, bs);
            }
        }
        __syncwarp();
        if (on) sts128(pA, o0, o1, o2, o3);
    }
    __syncwarp();
}

template <int R, int C, bool kAny>
__global__ void __launch_bounds__(kFThreads, 1)
gip_gauss_fused(const __grid_constant__ Job job, const __grid_constant__ FusedTiling tl) {
    using Cfg = FCfg<R, C, kAny>;
    constexpr int RC = Cfg::RC, R2 = 2 * R + 1, NS = Cfg::kSlots;
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pitch = job.src.pitch;

    // ---- tile -> (image, band, strip)
    unsigned tile = blockIdx.x;
    const int strip = (int)(tile % (unsigned)tl.strips); tile /= (unsigned)tl.strips;
    const int band = (int)(tile % (unsigned)tl.bands);
    const int64_t img = tile / (unsigned)tl.bands;
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * tl.band_rows;
    const int64_t Y1 = (Y0 + tl.band_rows < job.src.band_y1) ? Y0 + tl.band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;
    const int64_t Ystart = Y0 - R;                   // first input row of the tile
    const int nrows_in = (int)(Y1 - Y0) + 2 * R;
    const int nsteps = (nrows_in + kFK - 1) / kFK;
    const int64_t bxs = (int64_t)strip * Cfg::kUseful;   // first output byte of the strip
    const int64_t A0 = bxs - Cfg::kPad;              // image-row byte position of staged byte 0 (16-byte aligned)
    const uint32_t stage_s = smem_addr(smem);
    const uint32_t ring_s = stage_s + (uint32_t)(2 * kFK * Cfg::kStagePitch);
    const uint32_t obuf_s = ring_s + (uint32_t)(Cfg::kRingRows * kFRingPitch);      // kAny: the block of output rows

    if (warp < kFProd) {
        // ==================================== producer warp: rows 2*warp, 2*warp + 1 of every step ====================
        // copy plan of a row: chunk c = lane + 32 k holds image-row bytes [A0 + 16 c, +16); chunks outside the row are
        // skipped (clamp-to-edge bytes are written by the fix-up below)
        unsigned copy_mask = 0;
#pragma unroll
        for (int k = 0; k < Cfg::kCopyIters; k++) {
            const int c = lane + 32 * k;
            const int64_t g = A0 + 16 * (int64_t)c;
            if (c < Cfg::kRowChunks && g >= 0 && g + 16 <= pitch) copy_mask |= 1u << k;
        }
        const int lane_dst = 16 * (lane + (lane >> 3));          // padded offset of chunk `lane`; chunk lane + 32 k is 576 k further
        const int64_t lane_src = A0 + 16 * lane;
        // clamp-to-edge: staged bytes left of image byte 0 (strip 0) and right of the last image byte
        const int nleft = A0 < 0 ? (int)(-A0) : 0;               // staged indices [0, nleft)
        const int64_t row_end_idx = pitch - A0;                  // staged index of the first byte past the row
        const int nright = (row_end_idx < Cfg::kWinBytes) ? (int)((Cfg::kWinBytes - row_end_idx) < RC ? (Cfg::kWinBytes - row_end_idx) : RC) : 0;
        const bool edge_strip = nleft > 0 || nright > 0;
        // kAny: the last whole chunk of a row can end up to 15 bytes before the row does
        const bool edge_any = edge_strip || (kAny && row_end_idx < Cfg::kWinBytes + 16);

        const int64_t own_lo = job.src.band_y0 > 0 ? job.src.band_y0 : 0;
        const int64_t own_hi = job.src.band_y1 < job.height ? job.src.band_y1 : job.height;
        const int rel_fast_lo = (int)(own_lo + kFK - Ystart);    // row rel and row rel - K both lie in the band's own memory
        const int rel_fast_hi = (int)(own_hi - Ystart);
        const int64_t step_bytes = (int64_t)kFK * pitch;
        const uint8_t* gsrc[2] = {nullptr, nullptr};
        // kAny: gsrc[q] is the row's first byte; the row's address mod 16 of (buffer, q) is kept in a byte of skpack
        uint32_t skpack = 0;
        const int A0_i = (int)A0, pitch_i = (int)pitch;          // row positions fit 31 bits (checked on the host)
        auto stage_rows = [&](int step) {                        // this warp's two rows of `step` into buffer step & 1
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int rel = step * kFK + 2 * warp + q;
                if (rel < nrows_in) {
                    const uint32_t dst = stage_s + (uint32_t)(((step & 1) * kFK + 2 * warp + q) * Cfg::kStagePitch + lane_dst);
                    if (!kAny) {
                        if (rel >= rel_fast_lo && rel < rel_fast_hi && gsrc[q] != nullptr) gsrc[q] += step_bytes;
                        else gsrc[q] = job.src.row(clamp64(Ystart + rel, 0, job.height - 1), img) + lane_src;
#pragma unroll
                        for (int k = 0; k < Cfg::kCopyIters; k++)
                            if ((copy_mask >> k) & 1) cp_async16(dst + 576 * k, gsrc[q] + 512 * k);
                    } else {
                        if (rel >= rel_fast_lo && rel < rel_fast_hi && gsrc[q] != nullptr) gsrc[q] += step_bytes;
                        else gsrc[q] = job.src.row(clamp64(Ystart + rel, 0, job.height - 1), img);
                        const int a = (int)((uintptr_t)gsrc[q] & 15);
                        // Chunks that straddle the row's first / last byte also hold bytes of the row before / after it.
                        // That memory belongs to the same buffer unless the row is the first / last one of its buffer
                        // (band, above or below), so the straddling chunks are copied whole (the foreign bytes are
                        // overwritten by the clamp-to-edge fix-up or never read); only at a buffer's two ends are they
                        // left out and the row's own bytes fetched one by one when the row is consumed (bit 4 / 5).
                        uint32_t unsafe = 0;
                        if (edge_any) {
                            const int64_t y = clamp64(Ystart + rel, 0, job.height - 1);
                            bool first, last;
                            if (y < job.src.band_y0) { first = y == job.src.above_y0; last = y == job.src.band_y0 - 1; }
                            else if (y >= job.src.band_y1) { first = y == job.src.band_y1; last = y + 1 >= (job.src.band_y1 + R < job.height ? job.src.band_y1 + R : job.height); }
                            else { first = y == job.src.band_y0 && img == 0; last = y == job.src.band_y1 - 1 && img == job.batch - 1; }
                            if (pitch_i < 64) first = last = true;
                            unsafe = (first ? 16u : 0u) | (last ? 32u : 0u);
                        }
                        const int sh = 8 * (2 * (step & 1) + q);
                        skpack = (skpack & ~(255u << sh)) | (((uint32_t)a | unsafe) << sh);
                        // raw chunk c = global bytes [row + A0 - a + 16 c, +16)
                        int c_first = 0, c_end = Cfg::kRawChunks;
                        if (edge_any) {
                            if (a - A0_i > 0) c_first = (a - A0_i + ((unsafe & 16u) ? 15 : 0)) >> 4;
                            const int ce = (pitch_i - A0_i + a + ((unsafe & 32u) ? 0 : 15)) >> 4;
                            if (ce < c_end) c_end = ce;
                        }
                        const uint8_t* src = gsrc[q] + (A0_i - a) + 16 * lane;
#pragma unroll
                        for (int k = 0; k < Cfg::kCopyIters; k++) {
                            const int c = lane + 32 * k;
                            if (c >= c_first && c < c_end) cp_async16(dst + 576 * k, src + 512 * k);
                        }
                    }
                }
            }
            cp_async_commit();
        };
        // padded offsets of this lane's chunks: chunk 4*lane + i sits at lane_base + 16 i, plus 16 for i >= 4 in odd lanes
        const int lane_base = 16 * (4 * lane + (lane >> 1));
        const int lane_base_hi = lane_base + 16 * (lane & 1);

        stage_rows(0);
        stage_rows(1);
        for (int step = 0; step <= nsteps; step++) {
            if (tl.decoupled && step == nsteps) break;
            if (step < nsteps) {
                cp_async_wait<1>();                  // the rows of `step` have landed (the copies of step + 1 may be in flight)
                __syncwarp();
                const int rel = step * kFK + 2 * warp;
                const uint32_t rowA = stage_s + (uint32_t)(((step & 1) * kFK + 2 * warp) * Cfg::kStagePitch);
                const uint32_t rowB = rowA + Cfg::kStagePitch;
                if (kAny) {
#pragma unroll 1
                    for (int q = 0; q < 2; q++) {
                        if (rel + q < nrows_in) {
                            const uint32_t row = q ? rowB : rowA;
                            const uint32_t sk = skpack >> (8 * (2 * (step & 1) + q));
                            const int a = (int)(sk & 15u);
                            // first / last row of a buffer, edge strips: the row's bytes that no copied chunk covers
                            int nhead = 0, ntail = 0, head_pos = 0, tail_pos = 0;
                            uint32_t hb = 0, tb = 0;
                            const bool bytes = edge_any && (sk & 48u) != 0;
                            if (bytes) {
                                const uint8_t* rowp = job.src.row(clamp64(Ystart + rel + q, 0, job.height - 1), img);
                                int c_first = 0, c_end = Cfg::kRawChunks;
                                if (a - A0_i > 0) c_first = (a - A0_i + ((sk & 16u) ? 15 : 0)) >> 4;
                                const int ce = (pitch_i - A0_i + a + ((sk & 32u) ? 0 : 15)) >> 4;
                                if (ce < c_end) c_end = ce;
                                if (c_end < c_first) c_end = c_first;
                                const int cov_lo = A0_i - a + 16 * c_first, cov_hi = A0_i - a + 16 * c_end;   // row positions
                                const int win_lo = A0_i > 0 ? A0_i : 0;
                                const int win_hi = (A0_i + Cfg::kWinBytes < pitch_i) ? A0_i + Cfg::kWinBytes : pitch_i;
                                const int h_end = win_hi < cov_lo ? win_hi : cov_lo;
                                tail_pos = win_lo > cov_hi ? win_lo : cov_hi;
                                head_pos = win_lo;
                                nhead = h_end > win_lo ? h_end - win_lo : 0;
                                ntail = win_hi > tail_pos ? win_hi - tail_pos : 0;
                                if (lane < nhead) hb = rowp[head_pos + lane];
                                if (lane < ntail) tb = rowp[tail_pos + lane];
                            }
                            if (a != 0) realign_row<Cfg::kNeedChunks>(row, a, lane);
                            if (bytes) {
                                if (lane < nhead) asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + padded_off(head_pos - A0_i + lane)), "r"(hb) : "memory");
                                if (lane < ntail) asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + padded_off(tail_pos - A0_i + lane)), "r"(tb) : "memory");
                            }
                        }
                    }
                    if (edge_any) __syncwarp();
                }
                if (edge_strip && rel < nrows_in) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint32_t row = q ? rowB : rowA;
                        for (int i = lane; i < nleft; i += 32) {         // image position A0 + i < 0 -> pixel 0, same channel
                            const int ch = (int)(((A0 + i) % C + C) % C);
                            uint32_t v;
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row + padded_off(nleft + ch)));
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + padded_off(i)), "r"(v) : "memory");
                        }
                        for (int i = lane; i < nright; i += 32) {        // image position pitch + i -> last pixel, same channel
                            const int src = (int)row_end_idx - C + (i % C), dst = (int)row_end_idx + i;
                            uint32_t v;
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row + padded_off(src)));
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + padded_off(dst)), "r"(v) : "memory");
                        }
                    }
                    __syncwarp();
                }
                if (tl.decoupled && step >= NS) fbar_sync(NS + 1 + step % NS);      // the consumers are done with step - NS (same slot)
                if (rel < nrows_in) {
                    // ---- H pass of rows rel (low halves of every pair) and rel + 1 (high halves)
                    const uint32_t ringA = ring_s + (uint32_t)(((step % NS) * kFK + 2 * warp) * kFRingPitch + lane_base);
                    const uint32_t ringB = ringA + kFRingPitch;
                    uint32_t rawA[4 * Cfg::kLaneChunks], rawB[4 * Cfg::kLaneChunks];
                    uint64_t f[kFLane + 2 * RC];                 // converted inputs (row A, row B); only 2RC+1 are live at a time
#if GIP_FUSED_H_SCATTER
                    uint64_t hacc[kFLane];                       // partial sums per output byte; only 2RC+1 are live at a time
#endif
                    uint32_t zA[4], zB[4], wA[4], wB[4];
#pragma unroll
                    for (int m = 0; m < kFLane + 2 * RC; m++) {
                        const int bi = Cfg::kDelta + m;          // byte index inside the lane's chunks
                        if (m == 0 || (bi & 15) == 0) {          // first byte of a chunk: load it for both rows
                            const int i = bi >> 4;
                            const uint32_t off = (uint32_t)((i < 4 ? lane_base : lane_base_hi) + 16 * i);
                            const uint4 a = lds128(rowA + off), b = lds128(rowB + off);
                            rawA[4 * i] = a.x; rawA[4 * i + 1] = a.y; rawA[4 * i + 2] = a.z; rawA[4 * i + 3] = a.w;
                            rawB[4 * i] = b.x; rawB[4 * i + 1] = b.y; rawB[4 * i + 2] = b.z; rawB[4 * i + 3] = b.w;
                        }
                        f[m] = pack_f2(byte_to_float_bits_m<GIP_FUSED_XU_H>(rawA[bi >> 2], bi & 3), byte_to_float_bits_m<GIP_FUSED_XU_H>(rawB[bi >> 2], bi & 3));
                        const int j = m - 2 * RC;                // the output byte this input completes
#if GIP_FUSED_H_SCATTER
                        // scatter form: input m is tap t of output m - t C -- 2R+1 independent FMAs per input instead of a chain
                        // of 2R+1 dependent ones per output; every output still takes its taps in the reference's order
#pragma unroll
                        for (int t = 0; t < R2; t++) {
                            const int jo = m - t * C;
                            if (jo >= 0 && jo < kFLane)
                                hacc[jo] = t == 0 ? mul_rn_x2(f[m], splat_f2(job.weights[0])) : fma_rn_x2(f[m], splat_f2(job.weights[t]), hacc[jo]);
                        }
#endif
                        if (j >= 0) {
#if GIP_FUSED_H_SCATTER
                            const uint64_t acc = hacc[j];
#else
                            uint64_t acc = mul_rn_x2(f[j], splat_f2(job.weights[0]));
#pragma unroll
                            for (int t = 1; t < R2; t++) acc = fma_rn_x2(f[j + t * C], splat_f2(job.weights[t]), acc);
#endif
                            const uint64_t z = round_pair_sel<(GIP_FUSED_F2I & 1) != 0>(acc);
                            zA[j & 3] = lo_f2(z); zB[j & 3] = hi_f2(z);
                            if ((j & 3) == 3) {
                                wA[(j >> 2) & 3] = __byte_perm(__byte_perm(zA[0], zA[1], 0x4040), __byte_perm(zA[2], zA[3], 0x4040), 0x5410);
                                wB[(j >> 2) & 3] = __byte_perm(__byte_perm(zB[0], zB[1], 0x4040), __byte_perm(zB[2], zB[3], 0x4040), 0x5410);
                            }
                            if ((j & 15) == 15) {
                                sts128(ringA + 16 * (j >> 4), wA[0], wA[1], wA[2], wA[3]);
                                sts128(ringB + 16 * (j >> 4), wB[0], wB[1], wB[2], wB[3]);
                            }
                        }
                    }
                }
                __syncwarp();                        // every lane has read this buffer: refill it for step + 2
                stage_rows(step + 2);
            }
            if (tl.decoupled) fbar_arrive(1 + step % NS);
            else __syncthreads();
        }
    } else {
        // ==================================== consumer warp ====================================
        // Thread vt owns the 8-byte column group vt of the strip: 4 byte pairs, 2R+1 partial sums each.
        const int vt = tid - 32 * kFProd;
        const int64_t col = bxs + 8 * (int64_t)vt;
        // !kAny: pitch is a multiple of 16, so a group is inside the row or outside it.  kAny: the last group of a row may
        // be partial; its whole 8 bytes go to shared memory and the producers store the row's real length.
        const bool any = col < pitch && 8 * vt < Cfg::kUseful;
        const uint32_t ring_tid = ring_s + (uint32_t)(16 * ((vt >> 1) + (vt >> 4)) + 8 * (vt & 1));
        uint32_t o_s = obuf_s + 8u * (uint32_t)vt;       // kAny: this thread's bytes of the next output row
        uint8_t* optr = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + col;
        // Rows are consumed in blocks of 2R+1: inside a block the accumulator that takes tap k of row u is slot
        // (u - k) mod (2R+1), a compile-time register, and after a block every slot is back where it started -- no register
        // is ever moved.  Whole blocks only (the rows of a step that do not fill a block wait for the next step; the FIFO
        // has three steps of rows for that), except at the end of the tile.
        uint64_t acc[4][R2];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int t = 0; t < R2; t++) acc[q][t] = 0;
        const uint32_t ring_end = ring_tid + (uint32_t)(Cfg::kRingRows * kFRingPitch);
        uint32_t a = ring_tid;                       // FIFO row of input row `done`
        int done = 0;                                // input rows consumed so far (a multiple of 2R+1)
        auto v_block = [&](auto steady_tag) {
            constexpr bool kSteady = decltype(steady_tag)::value;      // every row of the block emits an output row
#pragma unroll
            for (int u = 0; u < R2; u++) {
                const uint2 w = lds64(a);
                a += kFRingPitch; if (a == ring_end) a = ring_tid;
                const uint32_t ww[2] = {w.x, w.y};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    // Measured on B200 (c4 stream / 8K RGB): all V-pass conversions on the XU pipe 21.13 -> 20.60 ms, r = 4 109 -> 106 us,
                    // but r = 1 71 -> 73 us and the any-alignment variant 33.6 -> 34.1 us; H-pass conversions there: no change
                    // or slower; F2I for the truncation: 3-9 % slower.
                    constexpr int kXuV = (!kAny && R >= 3) ? GIP_FUSED_XU_V : 0;
                    const uint64_t v = pack_f2(byte_to_float_bits_m<kXuV>(ww[q >> 1], 2 * (q & 1)), byte_to_float_bits_m<kXuV>(ww[q >> 1], 2 * (q & 1) + 1));
                    acc[q][u] = mul_rn_x2(v, splat_f2(job.weights[0]));
#pragma unroll
                    for (int k = 1; k < R2; k++)
                        acc[q][(u - k + R2) % R2] = fma_rn_x2(v, splat_f2(job.weights[k]), acc[q][(u - k + R2) % R2]);
                }
                const int rel = done + u;
                if (kSteady || (rel >= 2 * R && rel < nrows_in)) {     // this row completes output row Y0 + rel - 2R
                    uint32_t o[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint64_t z0 = round_pair_sel<(GIP_FUSED_F2I & 2) != 0>(acc[2 * h][(u + 1) % R2]), z1 = round_pair_sel<(GIP_FUSED_F2I & 2) != 0>(acc[2 * h + 1][(u + 1) % R2]);
                        o[h] = __byte_perm(__byte_perm(lo_f2(z0), hi_f2(z0), 0x4040), __byte_perm(lo_f2(z1), hi_f2(z1), 0x4040), 0x5410);
                    }
                    if (kAny && tl.ret) {
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(o_s), "r"(o[0]), "r"(o[1]) : "memory");
                        o_s += Cfg::kOutPitch;
                    } else {
                        stg64_stream(optr, o[0], o[1]);
                        optr += pitch;
                    }
                }
            }
            done += R2;
        };
        for (int step = 0; step <= nsteps; step++) {
            if (tl.decoupled) {
                if (step == 0) continue;
                fbar_sync(1 + (step - 1) % NS);      // the rows of step - 1 are in the FIFO
            }
            if (kAny && tl.ret) o_s = obuf_s + 8u * (uint32_t)vt;
            if (any) {
                const int avail = (step * kFK < nrows_in) ? step * kFK : nrows_in;      // rows of the steps before this one
                const int target = (avail == nrows_in) ? nrows_in : avail - avail % R2;
                while (done < target) {
                    if (done >= 2 * R && done + R2 <= nrows_in) v_block(std::true_type{});
                    else v_block(std::false_type{});
                }
            }
            if (kAny && tl.ret) {
                // The iteration's output rows are in shared memory: the consumer warps store them to global memory at the
                // aligned addresses (flush_row_any), a row per warp at a time.  Named barriers 7 and 8 count the consumer
                // threads only.  (Measured with the producers storing instead, as in the box kernel: 35.6 us on the c1 shape;
                // here the producers already carry the input shift and the consumers have the slack.)
                // input rows [target(step - 1), target(step)) were consumed; row rel >= 2R completed output row rel - 2R
                // (the same arithmetic as above, without `done`: threads outside the row never advance it)
                auto v_target = [&](int c) {
                    const int avail = (c * kFK < nrows_in) ? c * kFK : nrows_in;
                    return (avail == nrows_in) ? nrows_in : avail - avail % R2;
                };
                const int hi = v_target(step);
                const int lo = v_target(step - 1) > 2 * R ? v_target(step - 1) : 2 * R;
                asm volatile("bar.sync 7, %0;" ::"n"(32 * kFCons) : "memory");
                int64_t n = pitch - bxs; if (n > Cfg::kUseful) n = Cfg::kUseful;
                for (int i = warp - kFProd; i < hi - lo; i += kFCons) {
                    uint8_t* dst = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0 + (lo - 2 * R) + i) * pitch + bxs;
                    flush_row_any(obuf_s + (uint32_t)(i * Cfg::kOutPitch), dst, (int)n, lane);
                }
                asm volatile("bar.sync 8, %0;" ::"n"(32 * kFCons) : "memory");
            }
            if (!tl.decoupled) __syncthreads();
            else if (step >= 2 && step + NS - 2 <= nsteps - 1) fbar_arrive(NS + 1 + (step - 2) % NS);   // every row of step - 2 is consumed; a producer waits for it at step + NS - 2
        }
    }
}

template <int R, int C, bool kAny>
cudaError_t launch_fused(const Job& job, cudaStream_t stream, bool* handled) {
    using Cfg = FCfg<R, C, kAny>;
    static std::atomic<bool> attr_set[64];           // per instantiation and per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_set[dev]) {
        e = cudaFuncSetAttribute(gip_gauss_fused<R, C, kAny>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    const int sms = num_sms();
    if (sms <= 0) return cudaErrorInvalidDevice;
    FusedTiling tl;
    const int64_t pitch = job.src.pitch;
    tl.strips = (int)((pitch + Cfg::kUseful - 1) / Cfg::kUseful);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    // Row bands: the band count with the smallest (waves of resident CTAs) x (row steps of a tile: its rows, the 2R
    // rows that only fill the V accumulators, and the two-step pipeline fill).
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    int64_t max_bands = rows / 16; if (max_bands < 1) max_bands = 1;
    if (max_bands > 1024) max_bands = 1024;
    int64_t want = 1, best_cost = -1;
    for (int64_t nb = 1; nb <= max_bands; nb++) {
        if (per_band * nb > 0x7fffffff) break;
        const int64_t waves = (per_band * nb + sms - 1) / sms;
        const int64_t cost = waves * ((rows + nb - 1) / nb + 2 * R + 2 * kFK);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; want = nb; }
    }
    tl.bands = (int)want;
    tl.band_rows = (int)((rows + tl.bands - 1) / tl.bands);
    static const int coupled_env = [] { const char* e = getenv("GIP_GAUSS_COUPLED"); return e ? atoi(e) : 0; }();   // A/B runs
    tl.decoupled = coupled_env ? 0 : 1;
    tl.ret = (kAny && (pitch % 8 != 0 || job.src.image_stride % 8 != 0 || (uintptr_t)job.out % 8 != 0)) ? 1 : 0;
    const int64_t tiles = per_band * tl.bands;
    if (tiles > 0x7fffffff) return cudaSuccess;          // two-kernel path
    gip_gauss_fused<R, C, kAny><<<(unsigned)tiles, kFThreads, Cfg::kSmem, stream>>>(job, tl);
    count_launch();
    e = cudaGetLastError();
    *handled = (e == cudaSuccess);
    return e;
}

template <int R>
cudaError_t run_fused_radius(const Job& job, cudaStream_t stream, bool* handled) {
    const int64_t pitch = job.src.pitch;
    const bool aligned16 = (pitch % 16 == 0) && (job.src.image_stride % 16 == 0) && ((uintptr_t)job.src.band % 16 == 0) &&
                           ((uintptr_t)job.out % 16 == 0) && (!job.src.above || (uintptr_t)job.src.above % 16 == 0) &&
                           (!job.src.below || (uintptr_t)job.src.below % 16 == 0);
    if (aligned16) {
        if (job.channels == 4) return launch_fused<R, 4, false>(job, stream, handled);
        if (job.channels == 3) return launch_fused<R, 3, false>(job, stream, handled);
        return launch_fused<R, 1, false>(job, stream, handled);
    }
    static const int no_any = [] { const char* e = getenv("GIP_GAUSS_NO_FUSED_ANY"); return e ? atoi(e) : 0; }();   // A/B runs
    static const int coupled_any = [] { const char* e = getenv("GIP_GAUSS_COUPLED"); return e ? atoi(e) : 0; }();
    if (no_any || coupled_any || pitch > 0x7fff0000) return cudaSuccess;    // two-kernel path (the any-alignment variant needs the named barriers)
    if (job.channels == 4) return launch_fused<R, 4, true>(job, stream, handled);
    if (job.channels == 3) return launch_fused<R, 3, true>(job, stream, handled);
    return launch_fused<R, 1, true>(job, stream, handled);
}

}  // namespace
}  // namespace gip
