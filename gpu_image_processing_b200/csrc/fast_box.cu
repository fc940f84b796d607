// fast_box.cu -- fused, radius-independent box blur for sm_100a.
//
// Replaces boxBlur{Horizontal,Vertical}{Naive,Shared} + the d_temp round trip
// (/root/reference/cuda_lib/src/image_filters.cu:362-431, :448-673, :972-974) with ONE kernel
// that reads every input byte once and writes every output byte once.
//
// Decomposition.  A row of the image is a byte stream (pixel x, channel c at byte x*C+c); a
// horizontal box tap at pixel offset i is byte offset C*i, so the H pass is the same code for
// C = 1, 3, 4.  A CTA owns a column strip of `useful` output bytes and a band of rows and
// marches down the band K = 16 rows at a time.  The CTA is warp-specialised: 16 producer warps
// (stage + H pass, one row each) run one step ahead of 4 consumer warps (V pass + store); one
// __syncthreads per step hands a step's rows over.  640 threads, one CTA per SM (the first shape,
// 10 + 4 warps and two CTAs per SM, is still instantiated for A/B runs: GIP_BOX_HW=10).
//   stage   producer warp w copies input row w of the step (strip + halo) into its private
//           shared-memory row with 16-byte cp.async (LDGSTS), issued as soon as the previous row
//           has been read into registers, so the copy flies under the arithmetic.
//   H pass  one warp per row.  Lane l owns 60 consecutive bytes (15 words: an odd word stride,
//           so per-lane LDS.32/STS.32 are bank-conflict free).  Phase A forms, with IDP.4A, the
//           lane total of (entering byte - leaving byte) per channel; a warp inclusive scan of
//           the totals (SHFL) gives the true sliding-window sum at the start of each lane's run;
//           phase B replays the IDP.4A recurrence from that start and rounds every byte.
//           The first `nw` lanes are a warm-up zone whose leaving bytes read as zero, so the
//           window fills without a separate O(radius) initial sum: cost is independent of radius.
//           The rounded average (the reference's u8 intermediate, :394) goes to a ring of
//           2r+1+2K u8 rows in shared memory.
//   V pass  one thread per 16-byte column group keeps its sixteen window sums in registers across
//           the whole band: add the entering ring row, subtract the leaving one (IDP.4A), round,
//           store.  The intermediate never leaves the SM.
// Rounding.  The reference computes (uchar)(S*(1.0f/k)+0.5f) (:394, :429), which equals
// floor((S+r)/k) for every S in [0,255k], k odd <= 63 (tests/test_oracle.py proves it
// exhaustively).  Sums are kept as float bit patterns (2^23+S), and one FFMA2.RZ with per-radius
// constants (tools/box_magic.py, verified exhaustively in exact arithmetic) leaves
// floor((S+r)/k) in the low mantissa byte of two sums at once: no integer divide, no I2F/F2I.
#include <cstdlib>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kLaneWords = 15;
constexpr int kLaneBytes = 4 * kLaneWords;      // 60
constexpr int kWarpRun = 32 * kLaneBytes;       // 1920 bytes of recurrence per staged row
// Producer warps (HW, a template parameter) stage and filter one row each per step: K = HW rows per step.
//   HW = 16  640 threads, one CTA of 20 warps per SM: the default.  Measured on 4096x4096 RGBA against two
//            CTAs of 14 warps: -5 % at r <= 7, -9 % at r = 16, -27 % at r >= 17 (where only one 14-warp CTA fits);
//            20 or 24 producer warps, or 8 consumer warps on 8-byte column groups, were all slower.
//   HW = 10  448 threads, two CTAs per SM while the ring fits twice (radius <= 16 for RGBA): kept for A/B runs
constexpr int kHWarpsSmall = 10, kHWarpsBig = 16;
constexpr int kVWarps = 4;                      // consumer warps: 128 threads x 16-byte column groups >= 1856 bytes
constexpr int kGroupBytes = 16;                 // V-pass column group: one LDS.128 / STG.128
constexpr uint32_t kBias = 0x4B000000u;         // float 2^23
constexpr int kSmemLimit = 225 * 1024;
constexpr int kSmemTwoPerSM = 113 * 1024;

struct BoxMagic { uint32_t a_bits, c_bits; };
__constant__ BoxMagic c_box_magic[32] = {
#include "box_magic.inc"
};

struct BoxTiling {
    int nw;              // warm-up lanes = ceil((2r+1)*C / 60)
    int useful;          // output bytes per strip = (32 - nw) * 60 rounded down to a multiple of 16
    int ring_pitch;      // bytes per ring row = (32 - nw) * 60 rounded up to a multiple of 16
    int strips;          // strips per row
    int bands;           // row bands per image
    int band_rows;       // rows per band
    int ring_rows;       // 2r+1+2K
    int stage_row;       // bytes per staged row
};

// 4 window sums (float bit patterns) -> 4 rounded bytes packed in a word
__device__ __forceinline__ uint32_t round_pack(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                                               uint64_t a2, uint64_t c2) {
    uint32_t z0, z1, z2, z3;
    unpack_f2(fma_rz_x2(pack_f2(s0, s1), a2, c2), z0, z1);
    unpack_f2(fma_rz_x2(pack_f2(s2, s3), a2, c2), z2, z3);
    const uint32_t t0 = __byte_perm(z0, z1, 0x4040), t1 = __byte_perm(z2, z3, 0x4040);
    return __byte_perm(t0, t1, 0x5410);
}

template <int C, bool kVec, int HW>
__global__ void __launch_bounds__(32 * (HW + kVWarps), HW == kHWarpsSmall ? 2 : 1)
gip_box_fused(const __grid_constant__ Job job, const __grid_constant__ BoxTiling tl) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int NACC = (C == 3) ? 3 : 4;
    constexpr int kHWarps = HW, K = HW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < kHWarps;
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
    const int nsteps = (nrows_in + K - 1) / K;
    const int64_t bxs = (int64_t)strip * tl.useful;  // first output byte of the strip
    const int64_t e0 = bxs - (int64_t)kLaneBytes * tl.nw + (int64_t)r * C;   // image-row position of buffer index sh
    const int64_t B0 = e0 - sh;                      // image-row byte position of buffer index 0
    const uint32_t ring_s = smem_addr(smem + (size_t)K * tl.stage_row);
    const int ring_pitch = tl.ring_pitch;
    const int ring_bytes = tl.ring_rows * ring_pitch;
    const uint32_t mag_a = c_box_magic[r].a_bits, mag_c = c_box_magic[r].c_bits;
    const uint64_t mag_a2 = pack_f2(mag_a, mag_a), mag_c2 = pack_f2(mag_c, mag_c);

    if (producer) {
        // ==================================== producer warp ====================================
        const int skew = kVec ? (int)(((B0 % 16) + 16) % 16) : 0;
        uint8_t* my_row = smem + (size_t)warp * tl.stage_row + skew;     // buffer index 0 of this warp's staged row
        const uint32_t my_row_s = smem_addr(my_row);
        // Copy plan: image-row bytes [cs, ce) land at buffer index (pos - B0).  Only [e0, e0+1920) is
        // ever read; positions outside the image are replicated edge pixels.
        const int64_t lo = e0 < 0 ? 0 : e0;
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
        const uint32_t copy_dst = my_row_s + (uint32_t)(int)(cs - B0) + 16u * lane;
        uint8_t* copy_dst_g = my_row + (int)(cs - B0);
        const int head_idx = (int)(lo - B0);
        // clamp-to-edge: positions [e0, 0) (strip 0 only) and [pitch, pitch + rC) that fall inside the run;
        // e0 and pitch are multiples of C, so the channel of a replicated byte is its offset mod C.
        const int nleft = (e0 < 0) ? (int)(-e0) : 0;                       // buffer indices [sh, sh + nleft)
        int nright = 0;                                                    // buffer indices [right_idx, +nright)
        if (e0 + kWarpRun > pitch) {
            const int64_t over = e0 + kWarpRun - pitch;
            nright = (int)(over < (int64_t)r * C ? over : (int64_t)r * C);
        }
        const int right_idx = (int)(pitch - B0);
        const int64_t lane_off = cs + 16 * lane;

        // zero prefix: the leaving bytes of the warm-up zone.  Written once; the copies never touch it.
        for (int i = lane; i < sh; i += 32) my_row[i] = 0;

        uint32_t head_byte = 0, edge_l = 0, edge_r = 0;
        // Row pointers advance by K rows inside the band's own memory and are recomputed at the seams (clamped
        // rows at the image top / bottom, halo rows that live in a neighbour's buffer).
        const int64_t fast_lo = (job.src.band_y0 > 0 ? job.src.band_y0 : 0) + K;
        const int64_t fast_hi = job.src.band_y1 < job.height ? job.src.band_y1 : job.height;
        const int64_t step_bytes = (int64_t)K * pitch;
        const uint8_t* grow = nullptr;                     // row pointer of the row staged last
        const bool c0 = 16 * lane < ncopy, c1 = 16 * lane + 512 < ncopy, c2 = 16 * lane + 1024 < ncopy,
                   c3 = 16 * lane + 1536 < ncopy;
        const bool has_head = lane < nhead;
        auto stage_row = [&](int rel) {                    // rel = row index relative to Ystart (this warp: rel = warp mod K)
            if (rel < nrows_in) {
                const int64_t yy = Ystart + rel;
                if (yy >= fast_lo && yy < fast_hi && grow != nullptr) grow += step_bytes;
                else grow = job.src.row(clamp64(yy, 0, job.height - 1), img);
                if (kVec) {
                    const uint8_t* src = grow + lane_off;
                    if (c0) cp_async16(copy_dst, src);
                    if (c1) cp_async16(copy_dst + 512, src + 512);
                    if (c2) cp_async16(copy_dst + 1024, src + 1024);
                    if (c3) cp_async16(copy_dst + 1536, src + 1536);
                    if (has_head) head_byte = grow[lo + lane];
                } else {
                    const uint8_t* src = grow + cs;
                    for (int o = lane; o < ncopy; o += 32) copy_dst_g[o] = src[o];
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

        const uint32_t aL = my_row_s + (uint32_t)(kLaneBytes * lane);
        const uint32_t aE = aL + (uint32_t)sh;
        const uint32_t ring_lane = ring_s + (uint32_t)((lane - tl.nw) * kLaneBytes);

        stage_row(warp);
        int slot = 0;                                      // ring slot of the step's first row
        for (int step = 0; step < nsteps; step++) {
            const int rel0 = step * K;
            if (rel0 + warp < nrows_in) {
                cp_async_wait<0>();
                if (kVec && has_head) my_row[head_idx + lane] = (uint8_t)head_byte;
                if (nleft > 0)         // left image edge: replicate pixel 0
                    for (int i = lane; i < nleft; i += 32) my_row[sh + i] = (uint8_t)(edge_l >> (8 * (i % C)));
                if (nright > 0)        // right image edge: replicate the last pixel
                    for (int i = lane; i < nright; i += 32) my_row[right_idx + i] = (uint8_t)(edge_r >> (8 * (i % C)));
                __syncwarp();

                // leaving and entering words of this lane's run
                uint32_t Lw[kLaneWords], Ew[kLaneWords];
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
                stage_row(rel0 + K + warp);    // the staged row is in registers: refill it for the next step

                // phase A: lane totals of (entering - leaving) per channel
                int acc[NACC];
#pragma unroll
                for (int c = 0; c < NACC; c++) acc[c] = 0;
#pragma unroll
                for (int j = 0; j < kLaneWords; j++) {
                    const uint32_t pa = __byte_perm(Ew[j], Lw[j], 0x5140);   // in.b0 out.b0 in.b1 out.b1
                    const uint32_t pb = __byte_perm(Ew[j], Lw[j], 0x7362);   // in.b2 out.b2 in.b3 out.b3
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int ch = (C == 1) ? 0 : ((4 * j + k) % C);
                        acc[ch] = dp4a_us(k < 2 ? pa : pb, (k & 1) ? (int)0xFF010000 : 0x0000FF01, acc[ch]);
                    }
                }
                // exclusive scan over lanes -> window sum at the byte before this lane's run (as 2^23 + sum).
                // Window sums and lane totals are below 2^15 in magnitude, so two channels share one register
                // (lo + 65536 * hi in two's complement) and the SHFL scan runs on half as many registers.
                auto scan_incl = [&](int x) {
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= d) x += v;
                    }
                    return x;
                };
                if (C == 1) {
                    acc[0] = scan_incl(acc[0]) - acc[0] + (int)kBias;
                } else {
                    const int p01 = acc[0] + acc[1] * 65536;
                    const int e01 = scan_incl(p01) - p01;                    // exclusive, still packed
                    const int b0 = (int)(short)(e01 & 0xFFFF);
                    acc[0] = b0 + (int)kBias;
                    acc[1] = ((e01 - b0) >> 16) + (int)kBias;
                    if (C == 3) {
                        acc[2] = scan_incl(acc[2]) - acc[2] + (int)kBias;
                    } else {
                        const int p23 = acc[2] + acc[3 % NACC] * 65536;
                        const int e23 = scan_incl(p23) - p23;
                        const int b2 = (int)(short)(e23 & 0xFFFF);
                        acc[2] = b2 + (int)kBias;
                        acc[3 % NACC] = ((e23 - b2) >> 16) + (int)kBias;
                    }
                }
                // phase B: replay the recurrence from the true start value, round, store to the ring
                if (lane >= tl.nw) {
                    int rs = slot + warp; if (rs >= tl.ring_rows) rs -= tl.ring_rows;
                    const uint32_t dst = ring_lane + (uint32_t)(rs * ring_pitch);
#pragma unroll
                    for (int j = 0; j < kLaneWords; j++) {
                        const uint32_t pa = __byte_perm(Ew[j], Lw[j], 0x5140);
                        const uint32_t pb = __byte_perm(Ew[j], Lw[j], 0x7362);
                        uint32_t v[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int ch = (C == 1) ? 0 : ((4 * j + k) % C);
                            acc[ch] = dp4a_us(k < 2 ? pa : pb, (k & 1) ? (int)0xFF010000 : 0x0000FF01, acc[ch]);
                            v[k] = (uint32_t)acc[ch];
                        }
                        sts32(dst + 4 * j, round_pack(v[0], v[1], v[2], v[3], mag_a2, mag_c2));
                    }
                }
            }
            __syncthreads();      // this step's rows are in the ring
            slot += K; if (slot >= tl.ring_rows) slot -= tl.ring_rows;
        }
        __syncthreads();          // matches the consumers' last barrier
    } else {
        // ==================================== consumer warp ====================================
        // Thread vt owns the 16-byte column group vt of the strip: one LDS.128 per ring row, one STG.128 per output row,
        // a warp covers 512 contiguous bytes.
        const int vt = tid - 32 * kHWarps;
        const int nvt = tl.useful / kGroupBytes;
        const int64_t col = bxs + (int64_t)kGroupBytes * vt;
        int vbytes = 0;
        if (vt < nvt && col < pitch) vbytes = (pitch - col >= kGroupBytes) ? kGroupBytes : (int)(pitch - col);
        const bool any = vbytes > 0;
        uint32_t S[kGroupBytes];
#pragma unroll
        for (int i = 0; i < kGroupBytes; i++) S[i] = kBias;
        uint8_t* optr = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + col;  // next output row
        const uint32_t ring_tid = ring_s + (uint32_t)(kGroupBytes * (vt < nvt ? vt : 0));

        auto v_row = [&](uint32_t a_in, uint32_t a_out, bool leave, bool store) {
            const uint4 in4 = lds128(a_in);
            uint4 out4 = make_uint4(0u, 0u, 0u, 0u);
            if (leave) out4 = lds128(a_out);
            const uint32_t iw[4] = {in4.x, in4.y, in4.z, in4.w}, ow[4] = {out4.x, out4.y, out4.z, out4.w};
            uint32_t res[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t pa = __byte_perm(iw[w], ow[w], 0x5140);
                const uint32_t pb = __byte_perm(iw[w], ow[w], 0x7362);
                S[4 * w + 0] = (uint32_t)dp4a_us(pa, 0x0000FF01, (int)S[4 * w + 0]);
                S[4 * w + 1] = (uint32_t)dp4a_us(pa, (int)0xFF010000, (int)S[4 * w + 1]);
                S[4 * w + 2] = (uint32_t)dp4a_us(pb, 0x0000FF01, (int)S[4 * w + 2]);
                S[4 * w + 3] = (uint32_t)dp4a_us(pb, (int)0xFF010000, (int)S[4 * w + 3]);
                if (store) res[w] = round_pack(S[4 * w], S[4 * w + 1], S[4 * w + 2], S[4 * w + 3], mag_a2, mag_c2);
            }
            if (store) {
                if (kVec) {
                    stg128_stream(optr, make_uint4(res[0], res[1], res[2], res[3]));
                } else {
                    for (int b = 0; b < vbytes; b++) optr[b] = (uint8_t)(res[b >> 2] >> (8 * (b & 3)));
                }
                optr += pitch;
            }
        };

        __syncthreads();          // step 0 rows are in the ring
        int slot_in = 0;                                            // ring slot of the step's first row
        int slot_out = tl.ring_rows - (2 * r + 1);                  // ring slot of (first row - (2r+1))
        for (int step = 0; step < nsteps; step++) {
            const int rel0 = step * K;
            if (any) {
                uint32_t a_in = ring_tid + (uint32_t)(slot_in * ring_pitch);
                uint32_t a_out = ring_tid + (uint32_t)(slot_out * ring_pitch);
                const int wrap_in = tl.ring_rows - slot_in;            // first k whose entering slot wraps
                const int wrap_out = tl.ring_rows - slot_out;          // first k whose leaving slot wraps
                if (rel0 >= 2 * r + 1 && rel0 + K <= nrows_in) {
                    // steady state: every row has a leaving row and produces an output row
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        if (k == wrap_in) a_in -= (uint32_t)ring_bytes;
                        if (k == wrap_out) a_out -= (uint32_t)ring_bytes;
                        v_row(a_in, a_out, true, true);
                        a_in += (uint32_t)ring_pitch; a_out += (uint32_t)ring_pitch;
                    }
                } else {
                    const int nk = (nrows_in - rel0 < K) ? nrows_in - rel0 : K;
                    const int k_leave = 2 * r + 1 - rel0;      // rows k >= k_leave have a leaving row
                    const int k_store = 2 * r - rel0;          // rows k >= k_store produce an output row
#pragma unroll 1
                    for (int k = 0; k < nk; k++) {
                        if (k == wrap_in) a_in -= (uint32_t)ring_bytes;
                        if (k == wrap_out) a_out -= (uint32_t)ring_bytes;
                        v_row(a_in, a_out, k >= k_leave, k >= k_store);
                        a_in += (uint32_t)ring_pitch; a_out += (uint32_t)ring_pitch;
                    }
                }
            }
            __syncthreads();      // the producers may overwrite this step's leaving rows; step+1 rows are ready
            slot_in += K; if (slot_in >= tl.ring_rows) slot_in -= tl.ring_rows;
            slot_out += K; if (slot_out >= tl.ring_rows) slot_out -= tl.ring_rows;
        }
    }
}

int g_num_sms = 0;
const int g_box_hw = [] { const char* e = getenv("GIP_BOX_HW"); return e ? atoi(e) : 0; }();   // GIP_BOX_HW=10: the two-CTA form, for A/B runs

template <int C, bool kVec, int HW>
cudaError_t launch(const Job& job, const BoxTiling& tl, size_t smem, int64_t tiles, cudaStream_t stream) {
    static bool attr_set = false;   // per instantiation; the opt-in is idempotent
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gip_box_fused<C, kVec, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    gip_box_fused<C, kVec, HW><<<(unsigned)tiles, 32 * (HW + kVWarps), smem, stream>>>(job, tl);
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
    tl.useful = ((32 - tl.nw) * kLaneBytes) & ~15;
    tl.ring_pitch = ((32 - tl.nw) * kLaneBytes + 15) & ~15;
    const int64_t pitch = job.src.pitch;
    tl.strips = (int)((pitch + tl.useful - 1) / tl.useful);
    tl.stage_row = (sh + kWarpRun + 31 + 15) & ~15;
    auto smem_for = [&](int hw) { return (size_t)hw * tl.stage_row + (size_t)(2 * r + 1 + 2 * hw) * tl.ring_pitch; };
    // two CTAs of 14 warps per SM while they fit; otherwise one CTA of 20 warps (or of 14 if even that is too big)
    int hw = kHWarpsBig;
    if (g_box_hw == kHWarpsSmall || smem_for(hw) > (size_t)kSmemLimit) hw = kHWarpsSmall;
    const int ctas_per_sm = (hw == kHWarpsSmall && smem_for(hw) <= (size_t)kSmemTwoPerSM) ? 2 : 1;
    tl.ring_rows = 2 * r + 1 + 2 * hw;
    const size_t smem = smem_for(hw);
    if (smem > (size_t)kSmemLimit) return cudaSuccess;
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (rows > 0x3fffffff) return cudaSuccess;
    // Row bands.  A tile's time is proportional to its row steps (band rows + 2r halo rows + the pipeline fill), the
    // launch's to the number of waves of resident CTAs: take the band count with the smallest waves x steps.
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    const int64_t resident = (int64_t)g_num_sms * ctas_per_sm;
    const int64_t min_rows = 16;   // small images: short bands re-filter more halo rows but the march is latency-bound
    int64_t max_bands = rows / min_rows; if (max_bands < 1) max_bands = 1;
    if (max_bands > 1024) max_bands = 1024;
    int64_t want = 1, best_cost = -1;
    for (int64_t nb = 1; nb <= max_bands; nb++) {
        if (per_band * nb > 0x7fffffff) break;
        const int64_t waves = (per_band * nb + resident - 1) / resident;
        const int64_t steps = (rows + nb - 1) / nb + 2 * r + 2 * hw;
        const int64_t cost = waves * steps;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; want = nb; }
    }
    tl.bands = (int)want;
    tl.band_rows = (int)((rows + tl.bands - 1) / tl.bands);
    const int64_t tiles = per_band * tl.bands;
    if (tiles > 0x7fffffff) return cudaSuccess;      // general path
    const bool vec = (pitch % 16 == 0) && (job.src.image_stride % 16 == 0) &&
                     ((uintptr_t)job.src.band % 16 == 0) && ((uintptr_t)job.out % 16 == 0) &&
                     (!job.src.above || (uintptr_t)job.src.above % 16 == 0) &&
                     (!job.src.below || (uintptr_t)job.src.below % 16 == 0);
    cudaError_t err;
#define GIP_BOX_LAUNCH(C_, HW_) (vec ? launch<C_, true, HW_>(job, tl, smem, tiles, stream) : launch<C_, false, HW_>(job, tl, smem, tiles, stream))
    if (hw == kHWarpsBig) err = C == 4 ? GIP_BOX_LAUNCH(4, kHWarpsBig) : C == 3 ? GIP_BOX_LAUNCH(3, kHWarpsBig) : GIP_BOX_LAUNCH(1, kHWarpsBig);
    else                  err = C == 4 ? GIP_BOX_LAUNCH(4, kHWarpsSmall) : C == 3 ? GIP_BOX_LAUNCH(3, kHWarpsSmall) : GIP_BOX_LAUNCH(1, kHWarpsSmall);
#undef GIP_BOX_LAUNCH
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
