// fast_sobel.cu -- fused grayscale + 3x3 Sobel + magnitude for sm_100a, one pass, registers only.
//
// Replaces sobelEdgeDetectionNaive / sobelEdgeDetectionShared
// (/root/reference/cuda_lib/src/image_filters.cu:1152-1315, :1329-1597): the reference reads 27 bytes
// and converts 9 pixels to gray for every output pixel; here every input byte is read once and every
// gray value is computed once.
//
// A warp owns a strip of 120 output pixels and marches down a band of rows.  Lane l holds 4 consecutive
// pixels (12 / 16 / 4 bytes for RGB / RGBA / gray, loaded as 32-bit words: 384 contiguous bytes per warp
// row); lanes 0 and 31 are halo lanes whose pixels are only read by their neighbours.  The gray values of
// the previous two rows stay in registers; x+-1 neighbours come from __shfl_up/down.  No shared memory, no
// block barrier.  Arithmetic is packed float32x2 (FADD2/FMUL2/FFMA2: pixels 0,2 and 1,3 of a lane form
// the pairs) and reproduces the reference's rounding sequence exactly:
//   u8 -> float     PRMT into the mantissa of 2^23, minus 2^23 (exact)
//   gray            fma(B, .114f, fma(R, .299f, G * .587f))            (:1245, order read off the reference SASS)
//   level 2         gray := (float)(uchar)(gray + 0.5f)                  (:1443-1444)
//   gx, gy          single-rounded adds in row-major tap order          (:1246-1299)
//   magnitude       fma(gx, gx, gy*gy); sqrtf as the reference's inlined sequence
//                   (MUFU.RSQ, x*r, r/2, fma(-s,s,x), fma(d,h,s)); fminf(., 255); +0.5f; truncate (:1303-1305)
//   borders         pixels with x or y on the image edge are 0 in every channel (:1164-1176)
// The edge value is replicated into every channel, alpha included (:1311-1313).
#include <cstdlib>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kStripPixels = 120;       // 30 producing lanes x 4 pixels

struct SobelTiling {
    int strips, bands, band_rows;
    long long tiles;
};

struct GrayRow {          // gray of this lane's pixels (0,2), (1,3) and the shifted pairs (-1,1), (2,4)
    uint64_t A, B, PL, PR;
    // integer gray only (level 2, or gray input): right - left and left + 2*centre + right of this row, for the
    // pixel pairs (0,2) and (1,3)
    uint64_t DA, DB, SA, SB;
};

template <int C>
struct RowWords { uint32_t w[C]; };

// `off[k]` are in-row byte offsets made safe once per lane (words outside the row read offset 0:
// their pixels only feed border outputs, which are zero).
template <int C, bool kVec16>
__device__ __forceinline__ RowWords<C> load_words(const uint8_t* row, const int (&off)[C]) {
    RowWords<C> r;
    if (C == 4 && kVec16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + off[0]));
        r.w[0] = v.x; r.w[1 % C] = v.y; r.w[2 % C] = v.z; r.w[3 % C] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < C; k++) r.w[k] = __ldg(reinterpret_cast<const uint32_t*>(row + off[k]));
    }
    return r;
}

// byte `idx` (0..4C-1) of the lane's words as the float bit pattern 2^23 + byte
template <int C>
__device__ __forceinline__ uint32_t biased_byte(const RowWords<C>& r, int idx) {
    return __byte_perm(r.w[idx >> 2], 0x4B000000u, 0x7540 | (idx & 3));
}

template <int C, bool kU8>
__device__ __forceinline__ GrayRow make_gray(const RowWords<C>& r, int lane) {
    const uint64_t kNeg23 = splat_f2(-8388608.0f);
    uint64_t A, B;
    if (C == 1) {
        A = add_rn_x2(pack_f2(biased_byte<C>(r, 0), biased_byte<C>(r, 2)), kNeg23);
        B = add_rn_x2(pack_f2(biased_byte<C>(r, 1), biased_byte<C>(r, 3)), kNeg23);
    } else {
        // pixel j: R = byte C*j, G = C*j+1, B = C*j+2
        const uint64_t RA = add_rn_x2(pack_f2(biased_byte<C>(r, 0), biased_byte<C>(r, 2 * C)), kNeg23);
        const uint64_t GA = add_rn_x2(pack_f2(biased_byte<C>(r, 1), biased_byte<C>(r, 2 * C + 1)), kNeg23);
        const uint64_t BA = add_rn_x2(pack_f2(biased_byte<C>(r, 2), biased_byte<C>(r, 2 * C + 2)), kNeg23);
        const uint64_t RB = add_rn_x2(pack_f2(biased_byte<C>(r, C), biased_byte<C>(r, 3 * C)), kNeg23);
        const uint64_t GB = add_rn_x2(pack_f2(biased_byte<C>(r, C + 1), biased_byte<C>(r, 3 * C + 1)), kNeg23);
        const uint64_t BB = add_rn_x2(pack_f2(biased_byte<C>(r, C + 2), biased_byte<C>(r, 3 * C + 2)), kNeg23);
        const uint64_t kR = splat_f2(0.299f), kG = splat_f2(0.587f), kB = splat_f2(0.114f);
        A = fma_rn_x2(BA, kB, fma_rn_x2(RA, kR, mul_rn_x2(GA, kG)));
        B = fma_rn_x2(BB, kB, fma_rn_x2(RB, kR, mul_rn_x2(GB, kG)));
        if (kU8) {           // (float)(uchar)(gray + 0.5f): add, truncate on the 2^23 grid, remove the bias
            const uint64_t kHalf = splat_f2(0.5f), k23 = splat_f2(8388608.0f);
            A = add_rn_x2(add_rz_x2(add_rn_x2(A, kHalf), k23), kNeg23);
            B = add_rn_x2(add_rz_x2(add_rn_x2(B, kHalf), k23), kNeg23);
        }
    }
    GrayRow g;
    g.A = A; g.B = B;
    const uint32_t left = __shfl_up_sync(0xffffffffu, hi_f2(B), 1);     // pixel 3 of lane-1 == pixel -1
    const uint32_t right = __shfl_down_sync(0xffffffffu, lo_f2(A), 1);  // pixel 0 of lane+1 == pixel 4
    g.PL = pack_f2(left, lo_f2(B));      // pixels (-1, 1)
    g.PR = pack_f2(hi_f2(A), right);     // pixels ( 2, 4)
    if (kU8 || C == 1) {
        const uint64_t kP2 = splat_f2(2.0f);
        g.DA = sub_rn_x2(B, g.PL);       g.SA = add_rn_x2(fma_rn_x2(A, kP2, g.PL), B);
        g.DB = sub_rn_x2(g.PR, A);       g.SB = add_rn_x2(fma_rn_x2(B, kP2, A), g.PR);
    }
    (void)lane;
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

// Sobel magnitude of two pixels, as float bit patterns whose low byte is the rounded u8.
__device__ __forceinline__ uint64_t sobel_pair(uint64_t TL, uint64_t TC, uint64_t TR, uint64_t ML, uint64_t MR,
                                               uint64_t BL, uint64_t BC, uint64_t BR) {
    const uint64_t kM2 = splat_f2(-2.0f), kP2 = splat_f2(2.0f), kZero = splat_f2(0.0f);
    uint64_t gx = sub_rn_x2(kZero, TL);             // -TL
    uint64_t gy = fma_rn_x2(TC, kM2, gx);           // -TL - 2TC        (one rounding)
    gx = add_rn_x2(gx, TR);
    gy = sub_rn_x2(gy, TR);
    gx = fma_rn_x2(ML, kM2, gx);
    gx = fma_rn_x2(MR, kP2, gx);
    gx = sub_rn_x2(gx, BL);
    gy = add_rn_x2(gy, BL);
    gy = fma_rn_x2(BC, kP2, gy);
    gx = add_rn_x2(gx, BR);
    gy = add_rn_x2(gy, BR);
    return magnitude_pair(gx, gy);
}

// Integer gray values (level 2 rounds gray to u8, :1443-1444; gray input is u8 already): every partial sum of the
// reference's gx / gy is an integer below 2^24, so the float32 adds are exact in ANY order and the stencil can be
// taken apart: gx = D(top) + 2 D(mid) + D(bottom) with D = right - left, gy = S(bottom) - S(top) with
// S = left + 2 centre + right, D and S computed once per row.  Same bits as the reference, 6 instead of 11 adds.
__device__ __forceinline__ uint64_t sobel_pair_int(uint64_t Dt, uint64_t Dm, uint64_t Db, uint64_t St, uint64_t Sb) {
    const uint64_t gx = add_rn_x2(fma_rn_x2(Dm, splat_f2(2.0f), Dt), Db);
    const uint64_t gy = sub_rn_x2(Sb, St);
    return magnitude_pair(gx, gy);
}
__device__ __forceinline__ uint32_t clamp255(uint32_t z) { return min(z, 0x4B0000FFu); }

template <int C, bool kU8, bool kVec16>
__global__ void __launch_bounds__(kThreads, 3)
gip_sobel_fused(const __grid_constant__ Job job, const __grid_constant__ SobelTiling tl) {
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

    const int64_t x0 = (int64_t)strip * kStripPixels - 4 + 4 * lane;    // first pixel of this lane
    const int64_t boff = x0 * C;
    const bool stores = lane >= 1 && lane <= 30;
    // per-word load offsets, store predicates and border masks (x == 0, x == W-1 and x >= W give 0 / no store)
    int off[C];
    bool wvalid[C];
    uint32_t wmask[C];
#pragma unroll
    for (int k = 0; k < C; k++) {
        const int64_t o = boff + 4 * k;
        const bool inside = o >= 0 && o + 4 <= pitch;
        off[k] = inside ? (int)o : 0;
        wvalid[k] = stores && inside;
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t px = (o + b) / C;
            if (px >= 1 && px <= W - 2) m |= 0xFFu << (8 * b);
        }
        wmask[k] = m;
    }
    if (C == 4 && kVec16 && !(boff >= 0 && boff + 16 <= pitch)) off[0] = 0;

    // Row pointers advance by `pitch` inside the band's own rows and are recomputed at the seams
    // (image top / bottom clamp, rows owned by the neighbours above / below).
    const int nrows = (int)(Y1 - Y0);
    const int64_t src_lo = job.src.band_y0, src_hi = job.src.band_y1;
    const uint8_t* rp;                       // pointer of the row most recently loaded
    int64_t rp_y;
    // The last input row any output of this tile reads.  The load pipeline runs three rows ahead of the stencil: past
    // this row it reloads the same row instead (d_below holds only the rows the stencil needs, not the prefetch).
    const int64_t y_need_max = (Y1 < H - 1) ? Y1 : H - 1;
    auto next_row = [&](int64_t y) {         // rows are requested in increasing order
        if (y > y_need_max) y = y_need_max;
        if (y > src_lo && y < src_hi && y < H && rp_y == y - 1) rp += pitch;
        else rp = job.src.row(clamp64(y, 0, H - 1), img);
        rp_y = y;
        return rp;
    };
    uint8_t* out = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + boff;
    const int y_first = (int)Y0;             // image rows fit 31 bits here (checked on the host)
    const int y_last_interior = (int)(H - 2);

    auto emit = [&](const GrayRow& T, const GrayRow& M, const GrayRow& Bt, int y) {
        uint32_t w[C];
        if (y >= 1 && y <= y_last_interior) {
            uint64_t zA, zB;                                                                  // pixels 0, 2 and 1, 3
            if (kU8 || C == 1) {
                zA = sobel_pair_int(T.DA, M.DA, Bt.DA, T.SA, Bt.SA);
                zB = sobel_pair_int(T.DB, M.DB, Bt.DB, T.SB, Bt.SB);
            } else {
                zA = sobel_pair(T.PL, T.A, T.B, M.PL, M.B, Bt.PL, Bt.A, Bt.B);
                zB = sobel_pair(T.A, T.B, T.PR, M.A, M.PR, Bt.A, Bt.B, Bt.PR);
            }
            const uint32_t z0 = clamp255(lo_f2(zA)), z2 = clamp255(hi_f2(zA));
            const uint32_t z1 = clamp255(lo_f2(zB)), z3 = clamp255(hi_f2(zB));
            if (C == 1) {
                w[0] = __byte_perm(__byte_perm(z0, z1, 0x4040), __byte_perm(z2, z3, 0x4040), 0x5410);
            } else if (C == 3) {
                w[0] = __byte_perm(z0, z1, 0x4000); w[1 % C] = __byte_perm(z1, z2, 0x4400); w[2 % C] = __byte_perm(z2, z3, 0x4440);
            } else {
                w[0] = __byte_perm(z0, z0, 0x0000); w[1 % C] = __byte_perm(z1, z1, 0x0000);
                w[2 % C] = __byte_perm(z2, z2, 0x0000); w[3 % C] = __byte_perm(z3, z3, 0x0000);
            }
#pragma unroll
            for (int k = 0; k < C; k++) w[k] &= wmask[k];
        } else {
#pragma unroll
            for (int k = 0; k < C; k++) w[k] = 0;
        }
        if (C == 4 && kVec16) {
            if (wvalid[0]) stg128_stream(out, make_uint4(w[0], w[1 % C], w[2 % C], w[3 % C]));
        } else {
#pragma unroll
            for (int k = 0; k < C; k++)
                if (wvalid[k]) stg32_stream(out + 4 * k, w[k]);
        }
        out += pitch;
    };

    // rows Y0-1 and Y0 prime the pipeline; the words of the next THREE rows are always in flight (three word
    // buffers rotate with the three gray rows, so nothing is copied between iterations)
    rp_y = Y0 - 3;
    GrayRow R0 = make_gray<C, kU8>(load_words<C, kVec16>(next_row(Y0 - 1), off), lane);
    GrayRow R1 = make_gray<C, kU8>(load_words<C, kVec16>(next_row(Y0), off), lane);
    GrayRow R2;
    RowWords<C> W0 = load_words<C, kVec16>(next_row(Y0 + 1), off);
    RowWords<C> W1 = load_words<C, kVec16>(next_row(Y0 + 2), off);
    RowWords<C> W2 = load_words<C, kVec16>(next_row(Y0 + 3), off);
    for (int i = 0; i < nrows; i += 3) {
        const int y = y_first + i;
        // output row y needs rows y-1 (R0), y (R1), y+1 (W0 -> R2)
        R2 = make_gray<C, kU8>(W0, lane);
        W0 = load_words<C, kVec16>(next_row((int64_t)y + 4), off);
        emit(R0, R1, R2, y);
        if (i + 1 >= nrows) break;
        R0 = make_gray<C, kU8>(W1, lane);
        W1 = load_words<C, kVec16>(next_row((int64_t)y + 5), off);
        emit(R1, R2, R0, y + 1);
        if (i + 2 >= nrows) break;
        R1 = make_gray<C, kU8>(W2, lane);
        W2 = load_words<C, kVec16>(next_row((int64_t)y + 6), off);
        emit(R2, R0, R1, y + 2);
    }
}

int g_num_sms = 0;

template <int C, bool kU8, bool kVec16>
cudaError_t launch(const Job& job, const SobelTiling& tl, cudaStream_t stream) {
    const long long blocks = (tl.tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
    gip_sobel_fused<C, kU8, kVec16><<<(unsigned)blocks, kThreads, 0, stream>>>(job, tl);
    count_launch();
    return cudaGetLastError();
}

template <int C>
cudaError_t launch_c(const Job& job, const SobelTiling& tl, bool vec16, cudaStream_t stream) {
    if (job.sobel_u8_gray && C != 1)
        return vec16 ? launch<C, true, true>(job, tl, stream) : launch<C, true, false>(job, tl, stream);
    return vec16 ? launch<C, false, true>(job, tl, stream) : launch<C, false, false>(job, tl, stream);
}

}  // namespace

cudaError_t launch_fast_sobel(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int C = job.channels;
    const int64_t pitch = job.src.pitch;
    // 32-bit word loads/stores: every row must start on a 4-byte boundary
    const bool aligned4 = (pitch % 4 == 0) && (job.src.image_stride % 4 == 0) && ((uintptr_t)job.src.band % 4 == 0) &&
                          ((uintptr_t)job.out % 4 == 0) && (!job.src.above || (uintptr_t)job.src.above % 4 == 0) &&
                          (!job.src.below || (uintptr_t)job.src.below % 4 == 0);
    if (!aligned4) return cudaSuccess;               // general path
    if (g_num_sms == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    SobelTiling tl;
    tl.strips = (int)((job.width + kStripPixels - 1) / kStripPixels);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (rows > 0x3fffffff || job.height > 0x3fffffff || pitch > 0x7fffffff) return cudaSuccess;
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    // Row bands: the band count with the smallest (waves of resident warps) x (rows a warp marches, its 2 halo rows and
    // the 3-row load pipeline included).  24 warps are resident per SM (3 blocks of 8).  One full wave of tall bands
    // for a single image (measured on 8K RGB, warps per SM: 24 -> 76.6 us, 25 -> 95.6, 48 -> 78.3, 64 -> 79.9);
    // several finer waves for batches.
    static const int warps_per_sm = [] { const char* e = getenv("GIP_SOBEL_WARPS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 24; }();
    const int64_t resident = (int64_t)g_num_sms * warps_per_sm;
    int64_t max_bands = rows / 24; if (max_bands < 1) max_bands = 1;  // a band re-reads 2 halo rows
    if (max_bands > 1024) max_bands = 1024;
    int64_t bands = 1, best_cost = -1;
    for (int64_t nb = 1; nb <= max_bands; nb++) {
        const int64_t waves = (per_band * nb + resident - 1) / resident;
        const int64_t cost = waves * ((rows + nb - 1) / nb + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; bands = nb; }
    }
    tl.bands = (int)bands;
    tl.band_rows = (int)((rows + bands - 1) / bands);
    tl.tiles = per_band * bands;
    if ((tl.tiles + kWarpsPerBlock - 1) / kWarpsPerBlock > 0x7fffffffLL) return cudaSuccess;
    const bool vec16 = (C == 4) && (pitch % 16 == 0) && (job.src.image_stride % 16 == 0) &&
                       ((uintptr_t)job.src.band % 16 == 0) && ((uintptr_t)job.out % 16 == 0) &&
                       (!job.src.above || (uintptr_t)job.src.above % 16 == 0) &&
                       (!job.src.below || (uintptr_t)job.src.below % 16 == 0);
    cudaError_t err;
    if (C == 4)      err = launch_c<4>(job, tl, vec16, stream);
    else if (C == 3) err = launch_c<3>(job, tl, false, stream);
    else             err = launch_c<1>(job, tl, false, stream);
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
