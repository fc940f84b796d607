// fast_box.cu -- fused, radius-independent box blur for sm_100a.
//
// Replaces boxBlur{Horizontal,Vertical}{Naive,Shared} + the d_temp round trip
// (/root/reference/cuda_lib/src/image_filters.cu:362-431, :448-673, :972-974) with ONE kernel
// that reads every input byte once and writes every output byte once.
//
// Decomposition.  A row of the image is a byte stream (pixel x, channel c at byte x*C+c); a
// horizontal box tap at pixel offset i is byte offset C*i, so the H pass is the same code for
// C = 1, 3, 4.  A CTA owns a column strip of `useful` output bytes and a band of rows and
// marches down the band K = 15 rows at a time:
//   stage   warp w copies input row w of the step (strip + halo) into its private shared-memory
//           row with 16-byte cp.async (LDGSTS); the copy of step s+1 flies during the V pass of s.
//   H pass  one warp per row.  Lane l owns 60 consecutive bytes (15 words: an odd word stride,
//           so per-lane LDS.32/STS.32 are bank-conflict free).  It forms, with IDP.4A, the
//           running difference D = sum(entering byte - leaving byte) along its run; a warp
//           inclusive scan of the lane totals (SHFL) turns D into the true sliding-window sum.
//           The first `nw` lanes are a warm-up zone whose leaving bytes read as zero, so the
//           window fills without a separate O(radius) initial sum: cost is independent of radius.
//           The rounded average (the reference's u8 intermediate, :394) goes to a ring of
//           2r+1+2K u8 rows in shared memory.
//   V pass  one thread per 4-byte column group keeps its four window sums in registers across
//           the whole band: add the entering ring row, subtract the leaving one (IDP.4A), round,
//           store.  The intermediate never leaves the SM.  One __syncthreads per step.
// Rounding.  The reference computes (uchar)(S*(1.0f/k)+0.5f) (:394, :429), which equals
// floor((S+r)/k) for every S in [0,255k], k odd <= 63 (tests/test_oracle.py proves it
// exhaustively).  Sums are kept as float bit patterns (2^23+S), and one FFMA.RZ with per-radius
// constants (tools/box_magic.py, verified exhaustively in exact arithmetic) leaves
// floor((S+r)/k) in the low mantissa byte: no integer divide, no I2F/F2I.
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kLaneWords = 15;
constexpr int kLaneBytes = 4 * kLaneWords;      // 60
constexpr int kWarpRun = 32 * kLaneBytes;       // 1920 bytes of recurrence per staged row
constexpr int kWarps = 15;
constexpr int kThreads = 32 * kWarps;           // 480 = kWarpRun / 4 V-pass column groups
constexpr int K = kWarps;                       // rows per step
constexpr uint32_t kBias = 0x4B000000u;         // float 2^23
constexpr uint32_t kBiasMid = 0x4B400000u;      // float 1.5 * 2^23: integer steps on both sides

struct BoxMagic { uint32_t a_bits, c_bits; };
__constant__ BoxMagic c_box_magic[32] = {
#include "box_magic.inc"
};

struct BoxTiling {
    int nw;              // warm-up lanes = ceil((2r+1)*C / 60)
    int useful;          // output bytes per strip = (32 - nw) * 60
    int strips;          // strips per row
    int bands;           // row bands per image
    int band_rows;       // rows per band
    int ring_rows;       // 2r+1+2K rounded up to a multiple of K
    int stage_row;       // bytes per staged row
};

__device__ __forceinline__ uint32_t pack_low_bytes(float z0, float z1, float z2, float z3) {
    const uint32_t t0 = __byte_perm(__float_as_uint(z0), __float_as_uint(z1), 0x4040);
    const uint32_t t1 = __byte_perm(__float_as_uint(z2), __float_as_uint(z3), 0x4040);
    return __byte_perm(t0, t1, 0x5410);
}

template <int C, bool kVec>
__global__ void __launch_bounds__(kThreads, 1)
gip_box_fused(const __grid_constant__ Job job, const __grid_constant__ BoxTiling tl) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int NACC = (C == 3) ? 3 : 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = job.radius;
    const int sh = (2 * r + 1) * C;
    const int64_t pitch = job.src.pitch;

    // ---- tile -> (image, band, strip)
    unsigned tile = blockIdx.x;
    const int strip = (int)(tile % (unsigned)tl.strips); tile /= (unsigned)tl.strips;
    const int band = (int)(tile % (unsigned)tl.bands);
    const int64_t img = tile / (unsigned)tl.bands;
    const int64_t Y0 = job.src.band_y0 + (int64_t)band * tl.band_rows;
    const int64_t Y1 = (Y0 + tl.band_rows < job.src.band_y1) ? Y0 + tl.band_rows : job.src.band_y1;
    if (Y0 >= Y1) return;
    const int64_t Ystart = Y0 - r;                   // first input row fed to the recurrence
    const int nrows_in = (int)(Y1 - Y0) + 2 * r;
    const int64_t bxs = (int64_t)strip * tl.useful;  // first output byte of the strip
    const int64_t p0 = bxs - (int64_t)kLaneBytes * tl.nw;
    const int64_t e0 = p0 + (int64_t)r * C;          // image-row byte position of buffer index sh
    const int64_t B0 = e0 - sh;                      // image-row byte position of buffer index 0
    const int skew = kVec ? (int)(((B0 % 16) + 16) % 16) : 0;

    uint8_t* my_row = smem + (size_t)warp * tl.stage_row + skew;   // this warp's staged row, buffer index 0
    const uint32_t my_row_s = smem_addr(my_row);
    const uint32_t ring_s = smem_addr(smem + (size_t)K * tl.stage_row);
    const int ring_pitch = tl.useful;

    const float mag_a = __uint_as_float(c_box_magic[r].a_bits);
    const float mag_c = __uint_as_float(c_box_magic[r].c_bits);

    // Copy plan of one staged row: image-row bytes [cs, ce) land at buffer index (pos - B0).
    // Only [e0, e0+1920) is ever read; positions outside the image are replicated edge pixels.
    int64_t lo = e0 < 0 ? 0 : e0;
    int64_t hi = e0 + kWarpRun; if (hi > pitch) hi = pitch;
    int64_t cs = lo, ce = hi;                          // byte-exact range for the scalar path
    int nhead = 0;
    if (kVec) {                                        // whole 16-byte chunks; the ragged head goes by register
        cs = (lo + 15) & ~int64_t(15);
        ce = (hi + 15) & ~int64_t(15); if (ce > pitch) ce = pitch;
        if (cs > ce) cs = ce;
        nhead = (int)((cs < hi ? cs : hi) - lo); if (nhead < 0) nhead = 0;
    }
    const int ncopy = ce > cs ? (int)(ce - cs) : 0;
    const int dst0 = (int)(cs - B0);
    // clamp-to-edge: positions [bxs - rC, 0) and [pitch, pitch + rC) that fall inside the run
    const int nleft = (e0 < 0) ? (int)(-e0 < (int64_t)r * C + kLaneBytes * tl.nw ? -e0 : (int64_t)r * C + kLaneBytes * tl.nw) : 0;
    const int nright = (e0 + kWarpRun > pitch) ? (int)((e0 + kWarpRun - pitch) < (int64_t)r * C ? (e0 + kWarpRun - pitch) : (int64_t)r * C) : 0;

    // zero prefix: the leaving bytes of the warm-up zone.  Written once; the copies never touch it.
    for (int i = lane; i < sh; i += 32) my_row[i] = 0;

    uint32_t head_byte = 0, edge_l = 0, edge_r = 0;
    auto stage_row = [&](int rel) {                    // rel = row index relative to Ystart, this warp's row
        if (rel < nrows_in) {
            const int64_t y = clamp64(Ystart + rel, 0, job.height - 1);
            const uint8_t* grow = job.src.row(y, img);
            if (kVec) {
                const uint8_t* src = grow + cs;
                const uint32_t dst = my_row_s + (uint32_t)dst0;
                for (int o = lane * 16; o < ncopy; o += 512) cp_async16(dst + o, src + o);
                if (lane < nhead) head_byte = grow[lo + lane];
            } else {
                const uint8_t* src = grow + cs;
                uint8_t* dst = my_row + dst0;
                for (int o = lane; o < ncopy; o += 32) dst[o] = src[o];
            }
            if (nleft > 0) {       // every lane keeps the C bytes of pixel 0 / the last pixel
                uint32_t e = 0;
#pragma unroll
                for (int c = 0; c < C; c++) e |= (uint32_t)grow[c] << (8 * c);
                edge_l = e;
            }
            if (nright > 0) {
                uint32_t e = 0;
#pragma unroll
                for (int c = 0; c < C; c++) e |= (uint32_t)grow[pitch - C + c] << (8 * c);
                edge_r = e;
            }
        }
        cp_async_commit();
    };

    // ---- V-pass state: this thread's 4-byte column group
    const int nvw = tl.useful >> 2;
    const int64_t col = bxs + 4 * tid;
    int vbytes = 0;
    if (tid < nvw && col < pitch) vbytes = (pitch - col >= 4) ? 4 : (int)(pitch - col);
    int S[4] = {(int)kBias, (int)kBias, (int)kBias, (int)kBias};
    uint8_t* optr = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + col;  // next output row
    const uint32_t ring_tid = ring_s + 4u * (uint32_t)tid;

    const int nsteps = (nrows_in + K - 1) / K;
    stage_row(warp);
    int slot_in = 0;                                            // ring slot of the step's first row (multiple of K)
    int slot_out = tl.ring_rows - (2 * r + 1);                  // ring slot of (first row - (2r+1))

    for (int step = 0; step < nsteps; step++) {
        const int rel0 = step * K;

        // ================= H pass: warp `warp` filters its staged row =================
        if (rel0 + warp < nrows_in) {
            cp_async_wait<0>();
            if (kVec && lane < nhead) my_row[(int)(lo - B0) + lane] = (uint8_t)head_byte;
            if (nleft > 0) {       // left image edge: replicate pixel 0
                for (int i = lane; i < nleft; i += 32) {
                    const int64_t pos = -(int64_t)nleft + i;                    // negative image position
                    const int ch = (int)(((pos % C) + C) % C);
                    my_row[(int)(pos - B0)] = (uint8_t)(edge_l >> (8 * ch));
                }
            }
            if (nright > 0) {      // right image edge: replicate the last pixel
                for (int i = lane; i < nright; i += 32) {
                    const int64_t pos = pitch + i;
                    const int ch = (int)(pos % C);
                    my_row[(int)(pos - B0)] = (uint8_t)(edge_r >> (8 * ch));
                }
            }
            __syncwarp();

            // leaving and entering words of this lane's run
            uint32_t Lw[kLaneWords], Ew[kLaneWords];
            const uint32_t aL = my_row_s + (uint32_t)(kLaneBytes * lane);
            const uint32_t aE = aL + (uint32_t)sh;
            if (C == 4) {
#pragma unroll
                for (int j = 0; j < kLaneWords; j++) { Lw[j] = lds32(aL + 4 * j); Ew[j] = lds32(aE + 4 * j); }
            } else {
                const uint32_t bL = aL & ~3u, sL = (aL & 3u) * 8u, bE = aE & ~3u, sE = (aE & 3u) * 8u;
                uint32_t prevL = lds32(bL), prevE = lds32(bE);
#pragma unroll
                for (int j = 0; j < kLaneWords; j++) {
                    const uint32_t nl = lds32(bL + 4 * (j + 1)), ne = lds32(bE + 4 * (j + 1));
                    Lw[j] = funnel_bytes(prevL, nl, sL); Ew[j] = funnel_bytes(prevE, ne, sE);
                    prevL = nl; prevE = ne;
                }
            }
            __syncwarp();
            // the staged row is consumed: start copying this warp's row of the next step
            stage_row(rel0 + K + warp);

            int acc[NACC];
#pragma unroll
            for (int c = 0; c < NACC; c++) acc[c] = (int)kBiasMid;
            int val[kLaneBytes];
#pragma unroll
            for (int j = 0; j < kLaneWords; j++) {
                const uint32_t pa = __byte_perm(Ew[j], Lw[j], 0x5140);   // in.b0 out.b0 in.b1 out.b1
                const uint32_t pb = __byte_perm(Ew[j], Lw[j], 0x7362);   // in.b2 out.b2 in.b3 out.b3
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int q = 4 * j + k;
                    const int ch = (C == 1) ? 0 : (q % C);
                    acc[ch] = dp4a_us(k < 2 ? pa : pb, (k & 1) ? (int)0xFF010000 : 0x0000FF01, acc[ch]);
                    val[q] = acc[ch];
                }
            }
            // inclusive scan of the lane totals -> window sum at the byte before this lane's run
            float basef[NACC];
#pragma unroll
            for (int c = 0; c < ((C == 1) ? 1 : NACC); c++) {
                const int t = acc[c] - (int)kBiasMid;
                int x = t;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += v;
                }
                basef[c] = (float)(x - t - 4194304);       // (1.5*2^23 + D) + basef = 2^23 + D + base, exact
            }
            if (lane >= tl.nw) {
                const uint32_t dst = ring_s + (uint32_t)((slot_in + warp) * ring_pitch + (lane - tl.nw) * kLaneBytes);
#pragma unroll
                for (int j = 0; j < kLaneWords; j++) {
                    float z[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int q = 4 * j + k;
                        const int ch = (C == 1) ? 0 : (q % C);
                        z[k] = fma_rz(__fadd_rn(__int_as_float(val[q]), basef[ch]), mag_a, mag_c);
                    }
                    sts32(dst + 4 * j, pack_low_bytes(z[0], z[1], z[2], z[3]));
                }
            }
        }
        __syncthreads();      // ring rows of this step complete (and every warp finished the previous V pass)

        // ================= V pass: every thread slides its column group down the step's rows =================
        if (tid < nvw) {
            const int nk = (nrows_in - rel0 < K) ? nrows_in - rel0 : K;
            const int k_leave = 2 * r + 1 - rel0;      // rows k >= k_leave have a leaving row
            const int k_store = 2 * r - rel0;          // rows k >= k_store produce an output row
            uint32_t a_in = ring_tid + (uint32_t)(slot_in * ring_pitch);
            int so = slot_out;
#pragma unroll 5
            for (int k = 0; k < K; k++) {
                if (k < nk) {
                    const uint32_t in_w = lds32(a_in);
                    uint32_t out_w = 0;
                    if (k >= k_leave) out_w = lds32(ring_tid + (uint32_t)(so * ring_pitch));
                    const uint32_t pa = __byte_perm(in_w, out_w, 0x5140);
                    const uint32_t pb = __byte_perm(in_w, out_w, 0x7362);
                    S[0] = dp4a_us(pa, 0x0000FF01, S[0]);
                    S[1] = dp4a_us(pa, (int)0xFF010000, S[1]);
                    S[2] = dp4a_us(pb, 0x0000FF01, S[2]);
                    S[3] = dp4a_us(pb, (int)0xFF010000, S[3]);
                    if (k >= k_store) {
                        if (vbytes > 0) {
                            const uint32_t w = pack_low_bytes(fma_rz(__int_as_float(S[0]), mag_a, mag_c),
                                                              fma_rz(__int_as_float(S[1]), mag_a, mag_c),
                                                              fma_rz(__int_as_float(S[2]), mag_a, mag_c),
                                                              fma_rz(__int_as_float(S[3]), mag_a, mag_c));
                            if (kVec) {
                                stg32_stream(optr, w);
                            } else {
                                for (int b = 0; b < vbytes; b++) optr[b] = (uint8_t)(w >> (8 * b));
                            }
                        }
                        optr += pitch;
                    }
                    a_in += (uint32_t)ring_pitch;
                    if (++so == tl.ring_rows) so = 0;
                }
            }
        }
        slot_in += K; if (slot_in == tl.ring_rows) slot_in = 0;
        slot_out += K; if (slot_out >= tl.ring_rows) slot_out -= tl.ring_rows;
    }
}

int g_num_sms = 0;

template <int C, bool kVec>
cudaError_t launch(const Job& job, const BoxTiling& tl, size_t smem, int64_t tiles, cudaStream_t stream) {
    static bool attr_set = false;   // per instantiation; the opt-in is idempotent
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gip_box_fused<C, kVec>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             225 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    gip_box_fused<C, kVec><<<(unsigned)tiles, kThreads, smem, stream>>>(job, tl);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fast_box(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int r = job.radius, C = job.channels;
    if (r < 0 || r > kMaxFusedRadius) return cudaSuccess;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    BoxTiling tl;
    const int sh = (2 * r + 1) * C;
    tl.nw = (sh + kLaneBytes - 1) / kLaneBytes;
    tl.useful = (32 - tl.nw) * kLaneBytes;
    const int64_t pitch = job.src.pitch;
    tl.strips = (int)((pitch + tl.useful - 1) / tl.useful);
    tl.ring_rows = ((2 * r + 1 + 2 * K + K - 1) / K) * K;
    tl.stage_row = (sh + kWarpRun + 31 + 15) & ~15;
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (rows > 0x3fffffff) return cudaSuccess;
    // row bands: enough tiles to fill the SMs, but each band re-filters 2r halo rows
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    int64_t min_rows = 8 * (2 * r + 1); if (min_rows < 64) min_rows = 64;
    int64_t max_bands = rows / min_rows; if (max_bands < 1) max_bands = 1;
    int64_t want = (2 * (int64_t)g_num_sms + per_band - 1) / per_band;
    if (want < 1) want = 1;
    if (want > max_bands) want = max_bands;
    tl.bands = (int)want;
    tl.band_rows = (int)((rows + tl.bands - 1) / tl.bands);
    const int64_t tiles = per_band * tl.bands;
    if (tiles > 0x7fffffff) return cudaSuccess;      // general path
    const size_t smem = (size_t)K * tl.stage_row + (size_t)tl.ring_rows * tl.useful;
    if (smem > 225 * 1024) return cudaSuccess;
    const bool vec = (pitch % 16 == 0) && (job.src.image_stride % 16 == 0) &&
                     ((uintptr_t)job.src.band % 16 == 0) && ((uintptr_t)job.out % 4 == 0) &&
                     (!job.src.above || (uintptr_t)job.src.above % 16 == 0) &&
                     (!job.src.below || (uintptr_t)job.src.below % 16 == 0);
    cudaError_t err;
    if (C == 4)      err = vec ? launch<4, true>(job, tl, smem, tiles, stream) : launch<4, false>(job, tl, smem, tiles, stream);
    else if (C == 3) err = vec ? launch<3, true>(job, tl, smem, tiles, stream) : launch<3, false>(job, tl, smem, tiles, stream);
    else             err = vec ? launch<1, true>(job, tl, smem, tiles, stream) : launch<1, false>(job, tl, smem, tiles, stream);
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
