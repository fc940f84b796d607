// gauss_shift_impl.cuh -- the shift formulation of the separable Gaussian for sm_100a, radius 16..31 (the reference's
// level 2 accepts up to 63 taps, /root/reference/cuda_lib/src/image_filters.cu:729).  Included after fast_gauss_impl.cuh by the
// translation units of those radii; same arithmetic, same tap order (:86-99), same u8 intermediate (:102).
//
// acc[t] holds the partial sum that has seen taps 0..t; a new input v moves every partial sum one slot up through the FMA's
// destination (acc[t] = fma(v, w[t], acc[t-1]), t = 2R..1; acc[0] = v * w[0]) and acc[2R] is a finished output.  One step
// of code instead of the 2R+1 unrolled steps of the rotating form, no dependency inside a step (2R+1 independent FMAs), and
// 2R+1 packed accumulators per byte pair.  The first 2R inputs of a segment / band only feed the partial sums of outputs
// inside it (step s: slots t <= s) and the last 2R only those that still complete (t > j): no FMA is spent on outputs nobody
// stores, so segments and bands can be short.  Weights are symmetric (w[t] == w[2R-t] bit for bit, :28-38): R+1 uniform
// registers hold them.
#pragma once
#include "fast_gauss_impl.cuh"

namespace gip {
namespace {

// ------------------------------------------------------------------------------------------------
// H pass march of gip_gauss_h<R, C, true>: this thread = (row, channel, segment pair); pa / pb = its two segments in the
// staged tile, qa / qb = the same in the output tile.
// ------------------------------------------------------------------------------------------------
template <int R, int C>
__device__ __forceinline__ void h_shift_march(const Job& job, const uint8_t* pa, const uint8_t* pb, uint8_t* qa, uint8_t* qb) {
    constexpr int R2 = 2 * R + 1;
    constexpr int kSegPixels = HCfg<C, R>::kSegPixels;
    uint64_t acc[R2];
#pragma unroll
    for (int i = 0; i < R2; i++) acc[i] = 0;
#define GIP_W(T) splat_f2(job.weights[(T) <= R ? (T) : 2 * R - (T)])
    // Two accumulator sets: a step reads one and writes the other, so no value has to be in a particular register
    // when a loop iterates (ptxas then shifts the partial sums through the FMA destinations instead of with register
    // moves).  Steps come in pairs; head and tail run in chunks of kWChunk steps, a chunk runs the slots its last /
    // first step needs, in a rolled loop.
#define GIP_H_SHIFT(DST, SRC, S_IN, COND, FIRST)                                                          \
    {                                                                                            \
        const uint64_t v = to_float_pair(pa[(S_IN) * C], pb[(S_IN) * C]);                        \
        _Pragma("unroll")                                                                        \
        for (int t = 2 * R; t >= 1; t--)                                                         \
            if (COND) DST[t] = fma_rn_x2(v, GIP_W(t), SRC[t - 1]);                               \
        if (FIRST) DST[0] = mul_rn_x2(v, GIP_W(0));                                              \
    }
#define GIP_H_EMIT(SRC, S_OUT)                                                                           \
    {                                                                                            \
        const uint64_t z = round_pair(SRC[2 * R]);                                               \
        qa[(S_OUT) * C] = (uint8_t)lo_f2(z);                                                     \
        qb[(S_OUT) * C] = (uint8_t)hi_f2(z);                                                     \
    }
    static_assert(kWChunk % 2 == 0 && kSegPixels % 2 == 0, "steps come in pairs");
    uint64_t accB[R2];
#pragma unroll
    for (int i = 0; i < R2; i++) accB[i] = 0;
    constexpr int NCH = (2 * R + kWChunk - 1) / kWChunk;
#pragma unroll
    for (int g = 0; g < NCH; g++) {
        const int s_end = kWChunk * g + kWChunk < 2 * R ? kWChunk * g + kWChunk : 2 * R;
#pragma unroll 1
        for (int s = kWChunk * g; s < s_end; s += 2) {
            GIP_H_SHIFT(accB, acc, s, t <= kWChunk * g + kWChunk - 1, true)
            GIP_H_SHIFT(acc, accB, s + 1, t <= kWChunk * g + kWChunk - 1, true)
        }
    }
#pragma unroll 1
    for (int s = 2 * R; s < kSegPixels; s += 2) {
        GIP_H_SHIFT(accB, acc, s, true, true)
        GIP_H_EMIT(accB, s - 2 * R)
        GIP_H_SHIFT(acc, accB, s + 1, true, true)
        GIP_H_EMIT(acc, s + 1 - 2 * R)
    }
#pragma unroll
    for (int g = 0; g < NCH; g++) {
        const int j_end = kWChunk * g + kWChunk < 2 * R ? kWChunk * g + kWChunk : 2 * R;
#pragma unroll 1
        for (int j = kWChunk * g; j < j_end; j += 2) {
            GIP_H_SHIFT(accB, acc, kSegPixels + j, t > kWChunk * g, false)
            GIP_H_EMIT(accB, kSegPixels + j - 2 * R)
            GIP_H_SHIFT(acc, accB, kSegPixels + j + 1, t > kWChunk * g, false)
            GIP_H_EMIT(acc, kSegPixels + j + 1 - 2 * R)
        }
    }
#undef GIP_H_SHIFT
#undef GIP_H_EMIT
#undef GIP_W
}

// ------------------------------------------------------------------------------------------------
// V pass, shift formulation (radius > kMaxRotateRadius).  A thread owns NP byte pairs (adjacent columns = the two lanes
// of FFMA2) and marches down a band: acc[t] = fma(v, w[t], acc[t-1]) for t = 2R..1, acc[0] = v * w[0]; acc[2R] is a
// finished output row.  NP * (2R+1) packed accumulators: NP = 2 (a 4-byte column group) up to R = 19, NP = 1 above
// (126 registers at R = 31).  The 2R+1 FMAs of a step are independent of one another, so one warp per scheduler already
// keeps the FP32 pipe busy.  Rows arrive through a per-thread ring of kWPrefetch words in shared memory filled by 4-byte
// cp.async (LDGSTS) kWPrefetch rows ahead: no register ever waits on a load in flight, and the ring slot is a runtime
// index, so every loop stays rolled.  Like the H pass, the first 2R rows only feed the partial sums of the band's own
// output rows and the last 2R rows only those that still complete: no FMA is spent on rows outside the band, so bands
// can be short and the launch fills whole waves.
// ------------------------------------------------------------------------------------------------
constexpr int kWPrefetch = 8;
template <int R> struct WVCfg { static constexpr int NP = (R <= 19) ? 2 : 1; };

template <int R, int NP>
__global__ void __launch_bounds__(128)
gip_gauss_wv(const __grid_constant__ Job job, const uint8_t* __restrict__ tmp, int64_t ty0, int64_t ty1,
             int64_t img0, int nbands, int band_rows, int groups, int64_t tpitch, int out_aligned) {
    constexpr int R2 = 2 * R + 1;
    constexpr int P = kWPrefetch;
    constexpr int GBYTES = 2 * NP;                                // bytes per thread
    __shared__ uint32_t ring[P][128];
    const int64_t pitch = job.src.pitch;
    const int ci = blockIdx.x * 128 + threadIdx.x;
    if (ci >= groups) return;                                     // (no block-wide barrier below)
    const int band = blockIdx.y % nbands;
    const int64_t li = blockIdx.y / nbands;                       // image index inside the chunk
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * band_rows;
    const int64_t Y1 = (Y0 + band_rows < job.src.band_y1) ? Y0 + band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;
    const int rows = (int)(Y1 - Y0);
    const int64_t H = job.height;
    const int64_t col = (int64_t)GBYTES * ci;
    const int nbytes = (pitch - col >= GBYTES) ? GBYTES : (int)(pitch - col);
    uint8_t* optr = job.out + (img0 + li) * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + col;
    const int nsteps = rows + 2 * R;                              // input rows Y0-R .. Y1-1+R, clamped to the image
    const int s_lo = (Y0 - R < 0) ? (int)(R - Y0) : 0;
    const int s_hi_img = (int)(H - 1 - (Y0 - R));
    const int s_hi = s_hi_img < nsteps - 1 ? s_hi_img : nsteps - 1;
    // the aligned scratch word that holds this thread's bytes (scratch rows are 16-byte aligned and padded); row of step 0
    const uint8_t* tbase0 = tmp + li * (ty1 - ty0) * tpitch + (col & ~int64_t(3)) + (Y0 - R - ty0) * tpitch;
    const unsigned half_shift = (NP == 1) ? 16u * (unsigned)(ci & 1) : 0u;
    const uint32_t ring_s = smem_addr(&ring[0][threadIdx.x]);
    auto issue = [&](int s) {                                     // row of step s -> ring slot s mod P
        const int sc = s < s_lo ? s_lo : (s > s_hi ? s_hi : s);
        cp_async4(ring_s + (uint32_t)((s & (P - 1)) * 128 * 4), tbase0 + (int64_t)sc * tpitch);
        cp_async_commit();
    };
    // Two accumulator sets: a step reads one and writes the other (accB[t] = fma(v, w[t], accA[t-1]), next step back), so
    // no value has to be in a particular register when a loop iterates -- ptxas then shifts the partial sums through the
    // FMA destinations for free instead of with 2R+1 register moves per step.  Steps come in pairs.
    uint64_t accA[NP][R2], accB[NP][R2];
#pragma unroll
    for (int q = 0; q < NP; q++)
#pragma unroll
        for (int i = 0; i < R2; i++) { accA[q][i] = 0; accB[q][i] = 0; }
    uint64_t v[NP];
    auto next = [&](int s) {                                      // v <- the thread's bytes of step s; refill the slot
        cp_async_wait<P - 1>();
        uint32_t w = lds32(ring_s + (uint32_t)((s & (P - 1)) * 128 * 4));
        if (NP == 1) {
            w >>= half_shift;
            v[0] = to_float_pair(w & 0xFFu, (w >> 8) & 0xFFu);
        } else {
            v[0] = to_float_pair(w & 0xFFu, (w >> 8) & 0xFFu);
            v[NP - 1] = to_float_pair((w >> 16) & 0xFFu, w >> 24);
        }
        issue(s + P);
    };
    auto emit = [&](const uint64_t (&acc)[NP][R2]) {
        uint32_t t[NP];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const uint64_t z = round_pair(acc[q][2 * R]);
            t[q] = __byte_perm(lo_f2(z), hi_f2(z), 0x4040);       // the pair's two bytes in the low half
        }
        const uint32_t o = (NP == 2) ? __byte_perm(t[0], t[NP - 1], 0x5410) : t[0];
        if (out_aligned) {
            if (NP == 2) stg32_stream(optr, o);
            else *reinterpret_cast<unsigned short*>(optr) = (unsigned short)o;
        } else {
#pragma unroll
            for (int b = 0; b < GBYTES; b++)
                if (b < nbytes) optr[b] = (uint8_t)(o >> (8 * b));
        }
        optr += pitch;
    };
#define GIP_W(T) splat_f2(job.weights[(T) <= R ? (T) : 2 * R - (T)])
#define GIP_WV_STEP(DST, SRC, COND, FIRST)                                                               \
    {                                                                                                    \
        _Pragma("unroll")                                                                                \
        for (int t = 2 * R; t >= 1; t--)                                                                 \
            if (COND) {                                                                                  \
                _Pragma("unroll")                                                                        \
                for (int q = 0; q < NP; q++) DST[q][t] = fma_rn_x2(v[q], GIP_W(t), SRC[q][t - 1]);        \
            }                                                                                            \
        if (FIRST) {                                                                                     \
            _Pragma("unroll")                                                                            \
            for (int q = 0; q < NP; q++) DST[q][0] = mul_rn_x2(v[q], GIP_W(0));                           \
        }                                                                                                \
    }
#pragma unroll
    for (int i = 0; i < P; i++) issue(i);
    if (rows >= 2 * R) {
        static_assert(kWChunk % 2 == 0, "steps come in pairs");
        constexpr int NCH = (2 * R + kWChunk - 1) / kWChunk;
#pragma unroll
        for (int g = 0; g < NCH; g++) {                           // head: rows that only fill partial sums
            const int s_end = kWChunk * g + kWChunk < 2 * R ? kWChunk * g + kWChunk : 2 * R;
#pragma unroll 1
            for (int s = kWChunk * g; s < s_end; s += 2) {
                next(s);
                GIP_WV_STEP(accB, accA, t <= kWChunk * g + kWChunk - 1, true)
                next(s + 1);
                GIP_WV_STEP(accA, accB, t <= kWChunk * g + kWChunk - 1, true)
            }
        }
        int s = 2 * R;
#pragma unroll 1
        for (; s + 2 <= rows; s += 2) {                           // middle: full steps, one finished output row each
            next(s);
            GIP_WV_STEP(accB, accA, true, true)
            emit(accB);
            next(s + 1);
            GIP_WV_STEP(accA, accB, true, true)
            emit(accA);
        }
        if (s < rows) {                                           // odd count: one more step, then back into set A
            next(s);
            GIP_WV_STEP(accB, accA, true, true)
            emit(accB);
#pragma unroll
            for (int q = 0; q < NP; q++)
#pragma unroll
                for (int i = 0; i < R2; i++) accA[q][i] = accB[q][i];
        }
#pragma unroll
        for (int g = 0; g < NCH; g++) {                           // tail: only the partial sums that still complete
            const int j_end = kWChunk * g + kWChunk < 2 * R ? kWChunk * g + kWChunk : 2 * R;
#pragma unroll 1
            for (int j = kWChunk * g; j < j_end; j += 2) {
                next(rows + j);
                GIP_WV_STEP(accB, accA, t > kWChunk * g, false)
                emit(accB);
                next(rows + j + 1);
                GIP_WV_STEP(accA, accB, t > kWChunk * g, false)
                emit(accA);
            }
        }
    } else {
        // a band shorter than 2R rows (small images): every step in full, in place (register moves and all)
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            next(s);
            GIP_WV_STEP(accA, accA, true, true)
            if (s >= 2 * R) emit(accA);
        }
    }
    cp_async_wait<0>();
#undef GIP_WV_STEP
#undef GIP_W
}

template <int R>
cudaError_t launch_wv(const Job& job, const uint8_t* tmp, int64_t tpitch, int64_t ty0, int64_t ty1, int64_t img0,
                      int64_t nimg, cudaStream_t stream) {
    constexpr int NP = WVCfg<R>::NP;
    constexpr int GBYTES = 2 * NP;
    const int groups = (int)((job.src.pitch + GBYTES - 1) / GBYTES);
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    const int out_aligned = (job.src.pitch % GBYTES == 0) && (job.src.image_stride % GBYTES == 0) && ((uintptr_t)job.out % GBYTES == 0);
    const int64_t col_blocks = (groups + 127) / 128;
    static std::atomic<int> per_sm_cache[64];
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int per_sm = per_sm_cache[dev];
    if (per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gip_gauss_wv<R, NP>, 128, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        per_sm_cache[dev] = per_sm;
    }
    // Bands: a block's time is its rows (+ half of the 4R head / tail rows, which run partial steps, + a fixed start-up);
    // take the band count with the smallest (waves of resident blocks) x (block time).
    const int64_t resident = (int64_t)num_sms() * per_sm;
    int64_t band_rows = rows, best_cost = -1;
    for (int64_t nb = 1; nb <= 1024 && nb <= rows; nb++) {
        int64_t br = (rows + nb - 1) / nb;
        if (br < 2 * R && nb > 1) break;                          // shorter bands take the slow path
        const int64_t n_actual = (rows + br - 1) / br;
        if (n_actual * nimg > 65535) continue;
        const int64_t blocks = col_blocks * n_actual * nimg;
        const int64_t waves = (blocks + resident - 1) / resident;
        const int64_t cost = waves * (br + R + 12);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; band_rows = br; }
    }
    const int64_t nbands = (rows + band_rows - 1) / band_rows;
    if (nbands * nimg > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)col_blocks, (unsigned)(nbands * nimg));
    gip_gauss_wv<R, NP><<<grid, 128, 0, stream>>>(job, tmp, ty0, ty1, img0, (int)nbands, (int)band_rows, groups, tpitch, out_aligned);
    count_launch();
    return cudaGetLastError();
}

}  // namespace
}  // namespace gip
