// fast_sobel.cu -- fused grayscale + 3x3 Sobel + magnitude for sm_100a, one pass, registers only.
//
// Replaces sobelEdgeDetectionNaive / sobelEdgeDetectionShared
// (/root/reference/cuda_lib/src/image_filters.cu:1152-1315, :1329-1597): the reference reads 27 bytes
// and converts 9 pixels to gray for every output pixel; here every input byte is read once and every
// gray value is computed once.
//
// A warp owns a strip of 240 output pixels and marches down a band of rows.  Lane l holds 8 consecutive
// pixels of a row (24 / 32 / 8 bytes for RGB / RGBA / gray, loaded as 8- or 16-byte vectors: a warp row is
// 768 / 1024 / 256 contiguous bytes); lanes 0 and 31 are halo lanes whose pixels are only read by their
// neighbours.  The gray values of the previous two rows stay in registers; pixel -1 and pixel 8 of a
// lane come from __shfl_up/down.  No shared memory, no block barrier.
// Arithmetic is packed float32x2 (FADD2/FMUL2/FFMA2).  The two halves of a pair are pixels m and m+4
// of the lane, so the pairs (left, centre, right) of an output pair are other whole pairs of the same
// row -- nothing is re-packed except the two pairs that hold a shuffled neighbour.  The sequence of
// roundings is the reference's:
//   u8 -> float     PRMT into the mantissa of 2^23, minus 2^23 (exact)
//   gray            fma(B, .114f, fma(R, .299f, G * .587f))            (:1245, order read off the reference SASS)
//   level 2         gray := (float)(uchar)(gray + 0.5f)                  (:1443-1444)
//   gx, gy          single-rounded adds in row-major tap order          (:1246-1299)
//   magnitude       fma(gx, gx, gy*gy); sqrtf as the reference's inlined sequence
//                   (MUFU.RSQ, x*r, r/2, fma(-s,s,x), fma(d,h,s)); fminf(., 255); +0.5f; truncate (:1303-1305)
//   borders         pixels with x or y on the image edge are 0 in every channel (:1164-1176)
// The edge value is replicated into every channel, alpha included (:1311-1313).
// Warps whose 32 lanes all lie inside the row (every strip but the first and the last one or two of a
// row) run a path without per-word offsets, masks and store predicates.
// Rows at any byte alignment (VB = 1: odd pitches, unaligned base pointers) keep this kernel and this stencil code; a
// warp's row segments come in through a per-warp shared-memory ring filled with 16-byte cp.async from the aligned global
// chunks that cover them, and its output rows leave through a per-warp slab and whole 16-byte stores at the aligned
// global addresses (see ring_mem / oslab_mem in the kernel).
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kLanePixels = 8;
constexpr int kStripPixels = 30 * kLanePixels;       // 30 producing lanes

struct SobelTiling {
    int strips, bands, band_rows;
    long long tiles;
};

template <int C>
struct RowWords { uint32_t w[2 * C]; };              // the 8 pixels of one lane, one row

// Gray values of a row as the stencil wants them.  Float gray (level 1 colour): the pairs themselves.
// Integer gray (level 2 rounds gray to u8, :1443-1444; gray input is u8 already): every partial sum of the
// reference's gx / gy is an integer below 2^24, so the float32 adds are exact in ANY order and the stencil can be
// taken apart: gx = D(top) + 2 D(mid) + D(bottom) with D = right - left, gy = S(bottom) - S(top) with
// S = left + 2 centre + right, D and S computed once per row.  Same bits as the reference, 6 instead of 11 adds.
template <bool kInt> struct GrayRow;
template <> struct GrayRow<false> { uint64_t Q[4], L0, R3; };   // Q[m] = pixels (m, m+4); L0 = (-1, 3); R3 = (4, 8)
template <> struct GrayRow<true> { uint64_t D[4], S[4]; };

template <int VB>
__device__ __forceinline__ void load_vec(uint32_t* dst, const uint8_t* p) {
    if (VB == 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    } else if (VB == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        dst[0] = v.x; dst[1] = v.y;
    } else {
        dst[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    }
}
__device__ __forceinline__ void stg64_stream(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
template <int VB>
__device__ __forceinline__ void store_vec(uint8_t* p, const uint32_t* src) {
    if (VB == 16) stg128_stream(p, make_uint4(src[0], src[1], src[2], src[3]));
    else if (VB == 8) stg64_stream(p, src[0], src[1]);
    else stg32_stream(p, src[0]);
}

// byte `idx` (0..8C-1) of the lane's words as the float bit pattern 2^23 + byte
template <int C>
__device__ __forceinline__ uint32_t biased_byte(const RowWords<C>& r, int idx) {
    return __byte_perm(r.w[idx >> 2], 0x4B000000u, 0x7540 | (idx & 3));
}

// Channel value of pixels (m, m + 4) as a float pair.  Default: PRMT into the mantissa of 2^23 and one packed add of -2^23
// (ALU + FP32 pipe).  Channels in the mask GIP_SOBEL_XU_CH (bit 0 = R) convert with one I2F.U8 (byte selector) on the XU
// pipe instead: no PRMT, no add, a quarter of the rate.
#ifndef GIP_SOBEL_XU_CH
#define GIP_SOBEL_XU_CH 0
#endif
template <int C>
__device__ __forceinline__ uint64_t channel_pair(const RowWords<C>& r, int i0, int i1, int ch, uint64_t kNeg23) {
    if ((GIP_SOBEL_XU_CH >> ch) & 1) {
        const float a = (float)((r.w[i0 >> 2] >> (8 * (i0 & 3))) & 0xffu), b = (float)((r.w[i1 >> 2] >> (8 * (i1 & 3))) & 0xffu);
        return pack_f2(__float_as_uint(a), __float_as_uint(b));
    }
    return add_rn_x2(pack_f2(biased_byte<C>(r, i0), biased_byte<C>(r, i1)), kNeg23);
}

template <int C, bool kU8>
__device__ __forceinline__ GrayRow<kU8 || C == 1> make_gray(const RowWords<C>& r) {
    constexpr bool kInt = kU8 || C == 1;
    const uint64_t kNeg23 = splat_f2(-8388608.0f);
    uint64_t Q[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        if (C == 1) {
            Q[m] = add_rn_x2(pack_f2(biased_byte<C>(r, m), biased_byte<C>(r, m + 4)), kNeg23);
        } else {
            // pixel j: R = byte C*j, G = C*j+1, B = C*j+2
            const uint64_t Rv = channel_pair<C>(r, C * m, C * (m + 4), 0, kNeg23);
            const uint64_t Gv = channel_pair<C>(r, C * m + 1, C * (m + 4) + 1, 1, kNeg23);
            const uint64_t Bv = channel_pair<C>(r, C * m + 2, C * (m + 4) + 2, 2, kNeg23);
            uint64_t g = fma_rn_x2(Bv, splat_f2(0.114f), fma_rn_x2(Rv, splat_f2(0.299f), mul_rn_x2(Gv, splat_f2(0.587f))));
            if (kU8)             // (float)(uchar)(gray + 0.5f): add, truncate on the 2^23 grid, remove the bias
                g = add_rn_x2(add_rz_x2(add_rn_x2(g, splat_f2(0.5f)), splat_f2(8388608.0f)), kNeg23);
            Q[m] = g;
        }
    }
    const uint32_t left = __shfl_up_sync(0xffffffffu, hi_f2(Q[3]), 1);      // pixel 7 of lane-1 == pixel -1
    const uint32_t right = __shfl_down_sync(0xffffffffu, lo_f2(Q[0]), 1);   // pixel 0 of lane+1 == pixel 8
    const uint64_t L0 = pack_f2(left, lo_f2(Q[3]));       // pixels (-1, 3)
    const uint64_t R3 = pack_f2(hi_f2(Q[0]), right);      // pixels ( 4, 8)
    GrayRow<kInt> g;
    if constexpr (kInt) {
        const uint64_t k2 = splat_f2(2.0f);
        g.D[0] = sub_rn_x2(Q[1], L0);   g.S[0] = add_rn_x2(fma_rn_x2(Q[0], k2, L0), Q[1]);
        g.D[1] = sub_rn_x2(Q[2], Q[0]); g.S[1] = add_rn_x2(fma_rn_x2(Q[1], k2, Q[0]), Q[2]);
        g.D[2] = sub_rn_x2(Q[3], Q[1]); g.S[2] = add_rn_x2(fma_rn_x2(Q[2], k2, Q[1]), Q[3]);
        g.D[3] = sub_rn_x2(R3, Q[2]);   g.S[3] = add_rn_x2(fma_rn_x2(Q[3], k2, Q[2]), R3);
    } else {
#pragma unroll
        for (int m = 0; m < 4; m++) g.Q[m] = Q[m];
        g.L0 = L0; g.R3 = R3;
    }
    return g;
}

// sqrt(gx^2 + gy^2) of two pixels, rounded like the reference, as float bit patterns whose low byte is the u8.
__device__ __forceinline__ uint64_t magnitude_pair(uint64_t gx, uint64_t gy) {
    // fma(gx, gx, gy*gy); the 1e-30 only matters when both gradients are zero (keeps rsqrt finite)
    const uint64_t m2 = fma_rn_x2(gx, gx, fma_rn_x2(gy, gy, splat_f2(1e-30f)));
    // sqrtf, the reference's inlined sequence
    const float r0 = rsqrt_approx(__uint_as_float(lo_f2(m2))), r1 = rsqrt_approx(__uint_as_float(hi_f2(m2)));
    const uint64_t r = pack_f2(__float_as_uint(r0), __float_as_uint(r1));
    const uint64_t s0 = mul_rn_x2(m2, r);
    const uint64_t h = mul_rn_x2(r, splat_f2(0.5f));
    const uint64_t ns0 = mul_rn_x2(s0, splat_f2(-1.0f));
    const uint64_t d = fma_rn_x2(ns0, s0, m2);
    const uint64_t s = fma_rn_x2(d, h, s0);
    // (uchar)(fminf(s, 255) + 0.5f) == min(trunc(s + 0.5f), 255): the clamp moves to the integer side (see emit)
    return add_rz_x2(add_rn_x2(s, splat_f2(0.5f)), splat_f2(8388608.0f));
}

// Sobel magnitude of two pixels, as float bit patterns whose low byte is the rounded u8.  The reference adds the taps
// in row-major order with one rounding each (:1246-1299):
//   gx = ((((-TL + TR) - 2 ML) + 2 MR) - BL) + BR        gy = ((((-TL - 2 TC) - TR) + BL) + 2 BC) + BR
// 0 - TL is exact and round-to-nearest is symmetric, so -TL + TR == TR - TL and (-TL - 2 TC) - TR == -((TL + 2 TC) + TR)
// bit for bit: the negations are carried instead of computed (10 operations instead of 11).
__device__ __forceinline__ uint64_t sobel_pair(uint64_t TL, uint64_t TC, uint64_t TR, uint64_t ML, uint64_t MR,
                                               uint64_t BL, uint64_t BC, uint64_t BR) {
    const uint64_t kM2 = splat_f2(-2.0f), kP2 = splat_f2(2.0f);
    uint64_t gx = sub_rn_x2(TR, TL);
    uint64_t ngy = fma_rn_x2(TC, kP2, TL);          // -(gy so far)
    gx = fma_rn_x2(ML, kM2, gx);
    ngy = add_rn_x2(ngy, TR);
    gx = fma_rn_x2(MR, kP2, gx);
    uint64_t gy = sub_rn_x2(BL, ngy);
    gx = sub_rn_x2(gx, BL);
    gy = fma_rn_x2(BC, kP2, gy);
    gx = add_rn_x2(gx, BR);
    gy = add_rn_x2(gy, BR);
    return magnitude_pair(gx, gy);
}
__device__ __forceinline__ uint64_t sobel_pair_int(uint64_t Dt, uint64_t Dm, uint64_t Db, uint64_t St, uint64_t Sb) {
    const uint64_t gx = add_rn_x2(fma_rn_x2(Dm, splat_f2(2.0f), Dt), Db);
    const uint64_t gy = sub_rn_x2(Sb, St);
    return magnitude_pair(gx, gy);
}
__device__ __forceinline__ uint32_t clamp255(uint32_t z) { return min(z, 0x4B0000FFu); }

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// Output rows of colour images leave through a per-warp slab of shared memory: a lane's 24 (RGB) or 32 (RGBA) bytes
// are written to the slab, and the warp stores the slab with consecutive lanes on consecutive 8- / 16-byte chunks, so
// that every store instruction writes whole 32-byte sectors (lane-strided stores write every sector in 3 pieces: the
// round-1 capture showed 9.8 M L2 write sectors for a 3.1 M-sector image).
template <int C, int VB> struct Slab {
    static constexpr bool kUse = (C == 3 && VB == 8) || (C == 4 && VB == 16);
    static constexpr int kBytes = kUse ? 32 * 8 * C : 0;                       // one row of one warp
};

template <int C, bool kU8, int VB>
__global__ void __launch_bounds__(kThreads, 2)
gip_sobel_fused(const __grid_constant__ Job job, const __grid_constant__ SobelTiling tl) {
    constexpr bool kInt = kU8 || C == 1;
    constexpr bool kSlab = Slab<C, VB>::kUse;
    // VB == 1: rows that are not 4-byte aligned (odd pitches, e.g. the reference's 3239-pixel RGB shape).  Loads are
    // aligned words around the lane's bytes, funnel-shifted by the row's misalignment; a lane stores the aligned words
    // that straddle its own bytes and its left neighbour's last ones (by shuffle).  Strips at the row ends go byte by byte.
    constexpr bool kMis = VB == 1;
    constexpr int VBe = kMis ? 4 : VB;
    __shared__ __align__(16) uint8_t slab_mem[kSlab ? 2 * Slab<C, VB>::kBytes * kWarpsPerBlock : 16];
    const uint32_t slab0 = smem_addr(slab_mem) + (uint32_t)((threadIdx.x >> 5) * 2 * Slab<C, VB>::kBytes);
    // Rows at any alignment, strips inside the row: the warp's row segment (32 lanes x 8C bytes) is staged in shared
    // memory with 16-byte cp.async from the aligned global chunks that cover it, kRingAhead rows ahead of the stencil, and a
    // lane reads its own bytes from there (aligned LDS.32 + funnel shift).  The per-lane loads of the register path
    // (seven lane-strided LDG.32 per row: every one touches 24 sectors) kept the L1 pipe busy and the warps waiting.
    constexpr int kRingAhead = 4, kRingSlots = kRingAhead + 1;
    constexpr int kSegChunks = (32 * 8 * C + 16) / 16;          // aligned 16-byte chunks that cover a segment at any alignment
    constexpr int kSlotBytes = 32 * 8 * C + 32;
    __shared__ __align__(16) uint8_t ring_mem[VB == 1 ? kRingSlots * kSlotBytes * kWarpsPerBlock : 16];
    const uint32_t ring0 = smem_addr(ring_mem) + (uint32_t)((threadIdx.x >> 5) * kRingSlots * kSlotBytes);
    // ... and their output rows leave the same way: a lane writes its 8C bytes to a per-warp slab (lane l at 8 + 8C l, so
    // that lane 1's first byte sits on a 16-byte boundary) and the warp stores the 30 producing lanes' bytes with whole
    // 16-byte stores at the aligned global addresses (flush_row_any).  The per-lane aligned-word stores this replaces
    // (2C lane-strided STG.32 per row, each touching 24 sectors) were the L1 pipe's largest load after the input side went
    // through the ring.  (RGBA rows are word-aligned whenever the buffer is: C == 4 keeps the word stores, the ring and two
    // slabs of 1 KB rows would not fit the 48 KB of static shared memory.)
#ifndef GIP_SOBEL_FLUSH
#define GIP_SOBEL_FLUSH 1
#endif
    constexpr bool kFlush = VB == 1 && C != 4 && GIP_SOBEL_FLUSH != 0;
    constexpr int kOSlab = 32 * 8 * C + 32;
    __shared__ __align__(16) uint8_t oslab_mem[kFlush ? 2 * kOSlab * kWarpsPerBlock : 16];
    const uint32_t oslab0 = smem_addr(oslab_mem) + (uint32_t)((threadIdx.x >> 5) * 2 * kOSlab);
    constexpr int NW = 2 * C;                 // words per lane and row
    constexpr int NCH = 8 * C / VBe;          // vector chunks per lane and row
    constexpr int WPC = VBe / 4;              // words per chunk
    const int lane = threadIdx.x & 31;
    long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (tile >= tl.tiles) return;
    const int strip = (int)(tile % tl.strips); tile /= tl.strips;
    const int band = (int)(tile % tl.bands);
    const int64_t img = tile / tl.bands;
    const int64_t W = job.width, H = job.height, pitch = job.src.pitch;
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * tl.band_rows;
    const int64_t Y1 = (Y0 + tl.band_rows < job.src.band_y1) ? Y0 + tl.band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;

    const int64_t x0 = (int64_t)strip * kStripPixels - kLanePixels + kLanePixels * lane;    // first pixel of this lane
    const int64_t boff = x0 * C;
    const bool stores = lane >= 1 && lane <= 30;
    // (misaligned rows: an interior lane also keeps the aligned words around its bytes inside the row)
    const bool lane_inside = kMis ? (boff >= 4 && boff + 8 * C + 4 <= pitch) : (boff >= 0 && boff + 8 * C <= pitch);
    const bool edge = !__all_sync(0xffffffffu, lane_inside);    // first / last strips of a row: offsets, masks, predicates

    // Input rows Y0-1 .. Y1 of the tile, requested in order.  Rows Y0 .. Y1-1 are the band's own memory (one pointer,
    // advanced by `pitch`); only the first and the last row can lie across a seam (image top / bottom clamp, rows owned
    // by the neighbours above / below), so their pointers are computed once.  The load pipeline runs three rows ahead of
    // the stencil: past the last row any output needs it re-reads that row (d_below holds only the rows the stencil
    // needs, not the prefetch).
    const int nrows = (int)(Y1 - Y0);
    const uint8_t* const p_first = job.src.row(clamp64(Y0 - 1, 0, H - 1), img);
    const uint8_t* const p_last = job.src.row(clamp64(Y1, 0, H - 1), img);
    const uint8_t* rp = job.src.band + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch - pitch;   // "row Y0-1" of the band's memory
    int t_next = 0;                          // the next row of the band's own memory, relative to Y0
    // The words of three rows are in flight in registers; the row kL2Ahead rows further down is pulled into L2 at the same
    // time (one PREFETCH per lane and row), so that the register loads find it there instead of in HBM.  On the c4 frame
    // stream a quarter of all stall samples sat on the first use of a loaded row; measured 10.57 -> 9.75 ms per 4096 frames
    // at distance 6 (3: 10.04, 9: 10.04, 14: 10.24; six rows in flight in REGISTERS instead of three was slower: 11.0, and
    // 126 registers).
#ifndef GIP_SOBEL_L2_AHEAD
#define GIP_SOBEL_L2_AHEAD 6
#endif
    constexpr int kL2Ahead = GIP_SOBEL_L2_AHEAD;
    auto next_own_row = [&]() {              // rows Y0, Y0+1, ... ; from Y1 on: the last row, again and again
        rp = (t_next < nrows) ? rp + pitch : p_last;
        if (kL2Ahead > 0 && t_next + kL2Ahead < nrows && lane_inside)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (int64_t)kL2Ahead * pitch + boff));
        t_next++;
        return rp;
    };
    uint8_t* out = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + boff;
    const int y_first = (int)Y0;             // image rows fit 31 bits here (checked on the host)
    const int y_last_interior = (int)(H - 2);

    // magnitudes of output row y (rows T, M, B = y-1, y, y+1) -> z[j] = float bits whose low byte is pixel j
    auto stencil = [&](const GrayRow<kInt>& T, const GrayRow<kInt>& M, const GrayRow<kInt>& Bt, uint32_t (&z)[8]) {
        uint64_t zz[4];
        if constexpr (kInt) {
#pragma unroll
            for (int m = 0; m < 4; m++) zz[m] = sobel_pair_int(T.D[m], M.D[m], Bt.D[m], T.S[m], Bt.S[m]);
        } else {
            zz[0] = sobel_pair(T.L0, T.Q[0], T.Q[1], M.L0, M.Q[1], Bt.L0, Bt.Q[0], Bt.Q[1]);
            zz[1] = sobel_pair(T.Q[0], T.Q[1], T.Q[2], M.Q[0], M.Q[2], Bt.Q[0], Bt.Q[1], Bt.Q[2]);
            zz[2] = sobel_pair(T.Q[1], T.Q[2], T.Q[3], M.Q[1], M.Q[3], Bt.Q[1], Bt.Q[2], Bt.Q[3]);
            zz[3] = sobel_pair(T.Q[2], T.Q[3], T.R3, M.Q[2], M.R3, Bt.Q[2], Bt.Q[3], Bt.R3);
        }
#pragma unroll
        for (int m = 0; m < 4; m++) { z[m] = clamp255(lo_f2(zz[m])); z[m + 4] = clamp255(hi_f2(zz[m])); }
    };
    // the lane's output words: every channel of pixel j is z[j]
    auto pack = [&](const uint32_t (&z)[8], uint32_t (&w)[NW]) {
        if constexpr (C == 1) {
            w[0] = __byte_perm(__byte_perm(z[0], z[1], 0x4040), __byte_perm(z[2], z[3], 0x4040), 0x5410);
            w[1] = __byte_perm(__byte_perm(z[4], z[5], 0x4040), __byte_perm(z[6], z[7], 0x4040), 0x5410);
        } else if constexpr (C == 3) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                w[3 * q] = __byte_perm(z[4 * q], z[4 * q + 1], 0x4000);
                w[3 * q + 1] = __byte_perm(z[4 * q + 1], z[4 * q + 2], 0x4400);
                w[3 * q + 2] = __byte_perm(z[4 * q + 2], z[4 * q + 3], 0x4440);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = __byte_perm(z[j], z[j], 0x0000);
        }
    };

    auto march = [&](auto edge_tag) {
        constexpr bool kEdge = decltype(edge_tag)::value;
        // kEdge: per-chunk load offsets made safe once per lane (chunks outside the row read offset 0: their pixels
        // only feed border outputs, which are zero), store predicates and border masks (x == 0, x == W-1, x >= W).
        int coff[NCH];
        bool cvalid[NCH];
        uint32_t wmask[NW];
        if (kEdge) {
#pragma unroll
            for (int k = 0; k < NCH; k++) {
                const int64_t o = boff + VBe * k;
                const bool inside = o >= 0 && o + VBe <= pitch;
                coff[k] = inside ? (int)o : 0;
                cvalid[k] = stores && inside;
            }
#pragma unroll
            for (int k = 0; k < NW; k++) {
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int64_t ob = boff + 4 * k + b;
                    if (ob >= 0) {
                        const int64_t px = ob / C;
                        if (px >= 1 && px <= W - 2) m |= 0xFFu << (8 * b);
                    }
                }
                wmask[k] = m;
            }
        }
        auto load = [&](const uint8_t* row) {
            RowWords<C> r;
            if constexpr (kMis) {
                const uint8_t* a = row + boff;                         // the lane's first byte (outside the row in edge strips)
                const unsigned mis = (unsigned)((uintptr_t)a & 3);
                uint32_t aw[NW + 1];                                   // aligned words [a - mis, a - mis + 4 NW + 4)
                if (!kEdge || lane_inside) {                           // (edge strips: all but the one or two lanes at the row's ends)
#pragma unroll
                    for (int j = 0; j <= NW; j++) aw[j] = __ldg(reinterpret_cast<const uint32_t*>(a - mis + 4 * j));
                } else {
#pragma unroll
                    for (int j = 0; j <= NW; j++) {
                        const int64_t o = boff - (int64_t)mis + 4 * j;   // row-relative position of aligned word j
                        uint32_t v = 0;
                        if (o >= 0 && o + 4 <= pitch) {
                            v = __ldg(reinterpret_cast<const uint32_t*>(row + o));
                        } else if (o + 4 > 0 && o < pitch) {
                            // Touches a row end.  The ALIGNED word that holds the row's first / last bytes is read whole: it
                            // shares a 4-byte word (hence a page) with a byte of the row, and its foreign bytes only reach
                            // border outputs, which are zero.  (Byte loads here -- up to 28 dependent ones per row in the one
                            // or two lanes at a row's end -- made the edge strips' warps three times slower than the others,
                            // and a launch is as slow as its slowest warp: 68 us on the c1 shape.)
                            v = __ldg(reinterpret_cast<const uint32_t*>(row + o));
                        }
                        aw[j] = v;
                    }
                }
#pragma unroll
                for (int k = 0; k < NW; k++) r.w[k] = __funnelshift_r(aw[k], aw[k + 1], 8 * mis);
            } else {
#pragma unroll
                for (int k = 0; k < NCH; k++) load_vec<VBe>(&r.w[WPC * k], kEdge ? row + coff[k] : row + boff + VBe * k);
            }
            return r;
        };
        auto emit = [&](const GrayRow<kInt>& T, const GrayRow<kInt>& M, const GrayRow<kInt>& Bt, int y) {
            uint32_t w[NW];
            if (y >= 1 && y <= y_last_interior) {
                uint32_t z[8];
                stencil(T, M, Bt, z);
                pack(z, w);
                if (kEdge) {
#pragma unroll
                    for (int k = 0; k < NW; k++) w[k] &= wmask[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < NW; k++) w[k] = 0;
            }
            if constexpr (kSlab && !kEdge) {
                // out points at the lane's own bytes: the warp's segment starts 8*C*lane bytes before it
                const uint32_t slab = slab0 + (uint32_t)((y & 1) * Slab<C, VB>::kBytes);
                if constexpr (C == 3) {
#pragma unroll
                    for (int k = 0; k < 3; k++) sts64(slab + 24 * lane + 8 * k, w[2 * k], w[2 * k + 1]);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int c = lane + 32 * j;                 // 8-byte chunk of the segment; its owner is lane c / 3
                        const uint2 v = lds64(slab + 8 * c);
                        if (c >= 3 && c < 93) stg64_stream(out - 24 * lane + 8 * c, v.x, v.y);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const int c = 2 * lane + k;                  // 16-byte chunk; slot = c ^ bit 3 of c: conflict-free both ways
                        sts128(slab + 16 * (c ^ ((c >> 3) & 1)), w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int c = lane + 32 * j;                 // owner is lane c / 2
                        const uint4 v = lds128(slab + 16 * (c ^ ((c >> 3) & 1)));
                        if (c >= 2 && c < 62) stg128_stream(out - 32 * lane + 16 * c, v);
                    }
                }
            } else if constexpr (kFlush) {
                const uint32_t slab = oslab0 + (uint32_t)((y & 1) * kOSlab);
#pragma unroll
                for (int k = 0; k < NW / 2; k++) sts64(slab + 8 + 8 * C * lane + 8 * k, w[2 * k], w[2 * k + 1]);
                __syncwarp();
                // out points at the lane's own bytes; lane 1's are the first of the 30 x 8C bytes this warp writes (the last
                // strips of a row: as many of them as the row has)
                int n = 30 * 8 * C;
                if (kEdge) { const int64_t left = pitch - (boff - (int64_t)(8 * C) * (lane - 1)); if (left < n) n = left > 0 ? (int)left : 0; }
                if (n > 0) flush_row_any<(30 * 8 * C + 511) / 512>(slab + 8 + 8 * C, out - 8 * C * (lane - 1), n, lane);
            } else if constexpr (kMis) {
                // aligned word k = bytes [4k - mo, 4k - mo + 4) of the lane: the left neighbour's last mo bytes lead word 0
                const unsigned mo = (unsigned)((uintptr_t)out & 3);
                const uint32_t left = __shfl_up_sync(0xffffffffu, w[NW - 1], 1);
                // lanes whose aligned words lie inside the row and whose left neighbour holds real pixels store words; in
                // edge strips the one or two lanes at the row's ends store their bytes one by one
                const bool word_lane = !kEdge || (boff >= 8 * C && boff + 8 * C <= pitch);
                if (word_lane) {
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        const uint32_t v = __funnelshift_l(k == 0 ? left : w[k - 1], w[k], 8 * mo);
                        // lane 31 (halo) only completes the word that holds lane 30's last bytes; its own leading bytes are
                        // pixel 0 of the next strip's lane 1, which writes the same values
                        if (stores || (lane == 31 && k == 0 && mo != 0)) stg32_stream(out - mo + 4 * k, v);
                    }
                } else if (stores) {
#pragma unroll
                    for (int k = 0; k < NW; k++)
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int64_t ob = boff + 4 * k + b;
                            if (ob >= 0 && ob < pitch) out[4 * k + b] = (uint8_t)(w[k] >> (8 * b));
                        }
                }
                if (kEdge && word_lane && stores) {
                    // the lane's last mo bytes belong to the right neighbour's word 0: if that lane does not store words
                    // (row end, or lane 31 of an edge strip), store them here
                    const bool right_words = boff + 16 * C <= pitch && lane != 30;
                    if (!right_words)
                        for (unsigned b = 4 - mo; b < 4; b++) {
                            const int64_t ob = boff + 4 * (NW - 1) + b;
                            if (mo != 0 && ob < pitch) out[4 * (NW - 1) + b] = (uint8_t)(w[NW - 1] >> (8 * b));
                        }
                }
            } else {
#pragma unroll
                for (int k = 0; k < NCH; k++)
                    if (kEdge ? cvalid[k] : stores) store_vec<VBe>(out + VBe * k, &w[WPC * k]);
            }
            out += pitch;
        };

        if constexpr (kMis && (!kEdge || kFlush)) {
            // ---- shared-memory ring pipeline (see ring_mem)
            const int64_t boff0 = boff - (int64_t)(8 * C) * lane;          // the segment's first byte in the row
            uint32_t apack = 0;                                            // address mod 16 of the segment in each slot, 4 bits each
            int s_put = 0, s_get = 0;
            auto issue = [&](const uint8_t* row) {
                const uint8_t* g0 = row + boff0;
                const uint32_t a = (uint32_t)((uintptr_t)g0 & 15);
                apack = (apack & ~(15u << (4 * s_put))) | (a << (4 * s_put));
                const uint8_t* src = g0 - a + 16 * lane;
                const uint32_t dst = ring0 + (uint32_t)(s_put * kSlotBytes + 16 * lane);
                // (an aligned segment ends with its last whole chunk: every chunk copied holds bytes of the segment)
                const int nch = kSegChunks - (a == 0 ? 1 : 0);
#pragma unroll
                for (int q = 0; q < (kSegChunks + 31) / 32; q++) {
                    // Edge strips: only the chunks that hold at least one byte of the row.  (Such an aligned chunk lies in a
                    // page of the row's buffer; its foreign bytes -- and the stale bytes of the chunks left out -- only reach
                    // border outputs, which are zero.)
                    const bool in_row = !kEdge || (src + 512 * q + 16 > row && src + 512 * q < row + pitch);
                    if (lane + 32 * q < nch && in_row) cp_async16(dst + 512 * q, src + 512 * q);
                }
                cp_async_commit();
                if (++s_put == kRingSlots) s_put = 0;
            };
            auto take = [&]() {
                cp_async_wait<kRingAhead - 1>();
                __syncwarp();
                const uint32_t a = (apack >> (4 * s_get)) & 15u;
                const uint32_t base = ring0 + (uint32_t)(s_get * kSlotBytes) + ((a + (uint32_t)(8 * C * lane)) & ~3u);
                uint32_t aw[NW + 1];
#pragma unroll
                for (int j = 0; j <= NW; j++) aw[j] = lds32(base + 4 * j);
                RowWords<C> r;
#pragma unroll
                for (int k = 0; k < NW; k++) r.w[k] = __funnelshift_r(aw[k], aw[k + 1], 8 * (a & 3u));
                if (++s_get == kRingSlots) s_get = 0;
                return r;
            };
            issue(p_first);
#pragma unroll
            for (int k = 1; k < kRingAhead; k++) issue(next_own_row());
            GrayRow<kInt> R0 = make_gray<C, kU8>(take()); issue(next_own_row());
            GrayRow<kInt> R1 = make_gray<C, kU8>(take()); issue(next_own_row());
            // One copy of the row body (the gray rows move through registers instead of rotating by unrolling): the
            // any-alignment variant is 2.8 x the aligned kernel's code and was starved by instruction fetch (1.6 stall
            // cycles per issue with the inner and the edge march, 26 + 37 KB, resident on the same SM).
#pragma unroll 1
            for (int i = 0; i < nrows; i++) {
                const GrayRow<kInt> R2 = make_gray<C, kU8>(take()); issue(next_own_row());
                emit(R0, R1, R2, y_first + i);
                R0 = R1; R1 = R2;
            }
            cp_async_wait<0>();
            return;
        }
        if constexpr (kMis) {
            // edge strips at any alignment: the per-lane loads, one copy of the row body as well
            GrayRow<kInt> R0 = make_gray<C, kU8>(load(p_first));
            GrayRow<kInt> R1 = make_gray<C, kU8>(load(next_own_row()));
            RowWords<C> W0 = load(next_own_row());
            RowWords<C> W1 = load(next_own_row());
            RowWords<C> W2 = load(next_own_row());
#pragma unroll 1
            for (int i = 0; i < nrows; i++) {
                const GrayRow<kInt> R2 = make_gray<C, kU8>(W0);
                W0 = W1; W1 = W2; W2 = load(next_own_row());
                emit(R0, R1, R2, y_first + i);
                R0 = R1; R1 = R2;
            }
            return;
        }
        // rows Y0-1 and Y0 prime the pipeline; the words of the next THREE rows are always in flight (three word
        // buffers rotate with the three gray rows, so nothing is copied between iterations)
        GrayRow<kInt> R0 = make_gray<C, kU8>(load(p_first));
        GrayRow<kInt> R1 = make_gray<C, kU8>(load(next_own_row()));
        GrayRow<kInt> R2;
        RowWords<C> W0 = load(next_own_row());
        RowWords<C> W1 = load(next_own_row());
        RowWords<C> W2 = load(next_own_row());
        for (int i = 0; i < nrows; i += 3) {
            const int y = y_first + i;
            // output row y needs rows y-1 (R0), y (R1), y+1 (W0 -> R2)
            R2 = make_gray<C, kU8>(W0);
            W0 = load(next_own_row());
            emit(R0, R1, R2, y);
            if (i + 1 >= nrows) break;
            R0 = make_gray<C, kU8>(W1);
            W1 = load(next_own_row());
            emit(R1, R2, R0, y + 1);
            if (i + 2 >= nrows) break;
            R1 = make_gray<C, kU8>(W2);
            W2 = load(next_own_row());
            emit(R2, R0, R1, y + 2);
        }
    };
    if (edge) march(std::true_type{});
    else march(std::false_type{});
}

template <int C, bool kU8, int VB>
cudaError_t launch(const Job& job, SobelTiling tl, int64_t per_band, int64_t rows, cudaStream_t stream, bool* handled) {
    // Resident warps per SM of this instantiation on the current device (the attribute is per device).
    static std::atomic<int> per_sm_cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int per_sm = per_sm_cache[dev];
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gip_sobel_fused<C, kU8, VB>, kThreads, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        per_sm_cache[dev] = per_sm;
    }
    // Row bands: the band count with the smallest (waves of resident warps) x (rows a warp marches, its 2 halo rows and
    // the 3-row load pipeline included).  One full wave of tall bands for a single image, several finer waves for batches.
    static const int warps_env = [] { const char* s = getenv("GIP_SOBEL_WARPS_PER_SM"); return s ? atoi(s) : 0; }();
    const int warps_per_sm = warps_env > 0 ? warps_env : per_sm * kWarpsPerBlock;
    const int64_t resident = (int64_t)num_sms() * warps_per_sm;
    // Shortest band: 12 rows (a band re-reads 2 halo rows and fills a 3-row pipeline).  It was 24 until the end of round 2,
    // which left a 21 MB image (the c1 shape: 14 strips) with 1200 warps for 2368 resident ones: 22 -> 16 us aligned,
    // 63 -> 42 us at odd pitch with 12; 8 is the same, 6 slower.
    static const int min_band_rows = [] { const char* s = getenv("GIP_SOBEL_MIN_BAND_ROWS"); return s && atoi(s) > 0 ? atoi(s) : 12; }();   // A/B runs
    int64_t max_bands = rows / min_band_rows; if (max_bands < 1) max_bands = 1;  // a band re-reads 2 halo rows
    if (max_bands > 1024) max_bands = 1024;
    int64_t bands = 1, best_cost = -1;
    for (int64_t nb = 1; nb <= max_bands; nb++) {
        const int64_t waves = (per_band * nb + resident - 1) / resident;
        const int64_t cost = waves * ((rows + nb - 1) / nb + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; bands = nb; }
    }
    static const int bands_env = [] { const char* s = getenv("GIP_SOBEL_BANDS"); return s ? atoi(s) : 0; }();   // A/B runs
    if (bands_env > 0 && bands_env <= max_bands) bands = bands_env;
    tl.bands = (int)bands;
    tl.band_rows = (int)((rows + bands - 1) / bands);
    tl.tiles = per_band * bands;
    const long long blocks = (tl.tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL) return cudaSuccess;      // general path
    gip_sobel_fused<C, kU8, VB><<<(unsigned)blocks, kThreads, 0, stream>>>(job, tl);
    count_launch();
    e = cudaGetLastError();
    *handled = (e == cudaSuccess);
    return e;
}

template <int C, int VB>
cudaError_t launch_c(const Job& job, const SobelTiling& tl, int64_t per_band, int64_t rows, cudaStream_t stream, bool* handled) {
    if (job.sobel_u8_gray && C != 1) return launch<C, true, VB>(job, tl, per_band, rows, stream, handled);
    return launch<C, false, VB>(job, tl, per_band, rows, stream, handled);
}

}  // namespace

cudaError_t launch_fast_sobel(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int C = job.channels;
    const int64_t pitch = job.src.pitch;
    auto aligned = [&](int a) {
        return (pitch % a == 0) && (job.src.image_stride % a == 0) && ((uintptr_t)job.src.band % a == 0) &&
               ((uintptr_t)job.out % a == 0) && (!job.src.above || (uintptr_t)job.src.above % a == 0) &&
               (!job.src.below || (uintptr_t)job.src.below % a == 0);
    };
    const bool mis = !aligned(4);                     // rows at any byte alignment: the VB = 1 variants
    if (num_sms() <= 0) return cudaErrorInvalidDevice;
    SobelTiling tl;
    tl.strips = (int)((job.width + kStripPixels - 1) / kStripPixels);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (rows > 0x3fffffff || job.height > 0x3fffffff || pitch > 0x7fffffff) return cudaSuccess;
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    if (per_band > (int64_t)1 << 40) return cudaSuccess;
    if (mis) {
        if (C == 4) return launch_c<4, 1>(job, tl, per_band, rows, stream, handled);
        if (C == 3) return launch_c<3, 1>(job, tl, per_band, rows, stream, handled);
        return launch_c<1, 1>(job, tl, per_band, rows, stream, handled);
    }
    if (C == 4) return aligned(16) ? launch_c<4, 16>(job, tl, per_band, rows, stream, handled) : launch_c<4, 4>(job, tl, per_band, rows, stream, handled);
    if (C == 3) return aligned(8) ? launch_c<3, 8>(job, tl, per_band, rows, stream, handled) : launch_c<3, 4>(job, tl, per_band, rows, stream, handled);
    return aligned(8) ? launch_c<1, 8>(job, tl, per_band, rows, stream, handled) : launch_c<1, 4>(job, tl, per_band, rows, stream, handled);
}

}  // namespace gip
