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
// __syncthreads per step hands a step's rows over.  640 threads, one CTA per SM.
//   stage   producer warp w copies input row w of the step (strip + halo) into its private
//           shared-memory row with 16-byte cp.async (LDGSTS), issued as soon as the previous row
//           has been read into registers, so the copy flies under the arithmetic.  The copy starts
//           on the 16-byte boundary at or below the first byte it needs: no ragged head.
//   H pass  one warp per row.  Lane l owns 60 consecutive bytes (15 words: an odd word stride,
//           so per-lane LDS.32/STS.32 are bank-conflict free) and runs the sliding-window
//           recurrence S += entering byte - leaving byte with IDP.4A, rounding every byte.
//           The window sum at the start of a lane's run comes from one of two places:
//             direct  (2r+1)*C <= 60: the window is the first (2r+1)*C leaving bytes of the lane's
//                     own run -- summed per channel with the same PRMT + IDP.4A pairs, word by word
//                     under warp-uniform predicates.  All 32 lanes produce output; no scan, no warm-up.
//             scan    wider windows: phase A forms the lane totals of (entering - leaving) per
//                     channel, a warp inclusive scan (SHFL) turns them into the window sum at the
//                     start of each lane's run, phase B replays the recurrence.  The first `nw`
//                     lanes are a warm-up zone whose leaving bytes read as zero, so the window fills
//                     without an O(radius) initial sum: cost is independent of radius.
//           The rounded average (the reference's u8 intermediate, :394) goes to a ring of
//           u8 rows in shared memory.
//   V pass  one thread per 16-byte column group keeps its sixteen window sums in registers across
//           the whole band: add the entering ring row, subtract the leaving one (IDP.4A), round,
//           store.  The intermediate never leaves the SM.  The ring has a multiple of K rows and a
//           compile-time pitch, so a step's entering rows never wrap and its 16 rows are 16 LDS.128
//           at immediate offsets; the leaving rows wrap at most once per step, at a row that is
//           the same in every step (a step that wraps runs two short loops instead).
// Rows at any byte alignment (odd pitches, unaligned base pointers) run the same kernel: see kMode at the kernel and
// stage_row_any / flush_row_any -- chunks are copied from aligned global addresses into a staged row whose origin
// moves with the row, and the output rows return through shared memory so that the producer warps can store them with
// whole 16-byte stores at the aligned addresses.
// Rounding.  The reference computes (uchar)(S*(1.0f/k)+0.5f) (:394, :429), which equals
// floor((S+r)/k) for every S in [0,255k], k odd <= 63 (tests/test_oracle.py proves it
// exhaustively).  Sums are kept as float bit patterns (2^23+S), and one FFMA2.RZ with per-radius
// constants (tools/box_magic.py, verified exhaustively in exact arithmetic) leaves
// floor((S+r)/k) in the low mantissa byte of two sums at once: no integer divide, no I2F/F2I.
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <type_traits>
#include "common.cuh"
#include "device_utils.cuh"

namespace gip {
namespace {

constexpr int kLaneWords = 15;
constexpr int kLaneBytes = 4 * kLaneWords;      // 60
constexpr int kWarpRun = 32 * kLaneBytes;       // 1920 bytes of recurrence per staged row
constexpr int kHWarps = 16;                     // producer warps = rows per step
constexpr int K = kHWarps;
constexpr int kRingPitch = kWarpRun;            // bytes per ring row (compile time: row k of a step is an immediate offset)
// Consumer threads own GB-byte column groups (GB = 16: 4 consumer warps, LDS.128 / STG.128; GB = 8: 8 consumer warps,
// LDS.64 / STG.64 -- half the work per warp and step, so the V pass is less of a critical path between two barriers).
constexpr int v_warps(int gb) { return (kWarpRun / gb + 31) / 32; }
constexpr int box_threads(int gb) { return 32 * (kHWarps + v_warps(gb)); }
constexpr uint32_t kBias = 0x4B000000u;         // float 2^23
constexpr int kSmemLimit = 225 * 1024;

struct BoxMagic { uint32_t a_bits, c_bits; };
__constant__ BoxMagic c_box_magic[32] = {
#include "box_magic.inc"
};

struct BoxTiling {
    int nw;              // warm-up lanes: 0 (direct) or ceil((2r+1)*C / 60) (scan)
    int useful;          // output bytes per strip = (32 - nw) * 60 rounded down to a multiple of 16
    int strips;          // strips per row
    int bands;           // row bands per image
    int band_rows;       // rows per band
    int ring_rows;       // 2r+1+2K rounded up to a multiple of K
    int stage_row;       // bytes per staged row
    int decoupled;       // 1: producers and consumers meet on named barriers (full / empty per step parity), not __syncthreads
    int ret;             // kMode 2: 1 = output rows return through shared memory (the producers store them); 0 = output rows are
                         // 8-byte aligned (only the staging needs the per-row skew) and the consumers store them directly
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

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void stg64_stream(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// Named barriers (ids 1..4; 0 is __syncthreads): producers arrive on FULL[s & 1] when the rows of step s are in the ring,
// consumers wait there; consumers arrive on EMPTY[s & 1] when they are done with step s, producers wait there before
// they overwrite ring rows in step s + 2.  Every barrier counts all threads of the CTA (arrivals + waiters).
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// kMode 1: every row and buffer 16-byte aligned.  Rows at any byte alignment (odd pitches, unaligned base pointers):
//   kMode 2  output rows return through shared memory: the consumers write them (8-byte STS, exactly the aligned kernel's
//            work) into a double-buffered block of K rows after the ring, and one step later producer warp w stores row w
//            with whole 16-byte stores at the aligned global addresses (flush_row_any).  The producers have the slack:
//            with the stores on the consumer side (kMode 0) they sat on the EMPTY barrier for a third of their time.
//   kMode 0  the consumers store themselves (store_segment_dup); only when the extra 60 KB do not fit (radius > 15).
template <int C, int kMode, bool kDirect, int GB>
__global__ void __launch_bounds__(box_threads(GB), 1)
gip_box_fused(const __grid_constant__ Job job, const __grid_constant__ BoxTiling tl) {
    constexpr bool kVec = kMode == 1, kRet = kMode == 2;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int NACC = (C == 3) ? 3 : 4;
    constexpr int kGroupBytes = GB, GW = GB / 4;
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
    const int nw = kDirect ? 0 : tl.nw;
    const int64_t bxs = (int64_t)strip * tl.useful;  // first output byte of the strip
    const int64_t e0 = bxs - (int64_t)kLaneBytes * nw + (int64_t)r * C;   // image-row position of buffer index sh
    const int64_t B0 = e0 - sh;                      // image-row byte position of buffer index 0
    const uint32_t ring_s = smem_addr(smem + (size_t)K * tl.stage_row);
    const int ring_bytes = tl.ring_rows * kRingPitch;
    const uint32_t obuf_s = ring_s + (uint32_t)ring_bytes;        // kRet: 2 x K output rows (+ 16 bytes of slack)
    const int nthreads = box_threads(GB);
    const uint32_t mag_a = c_box_magic[r].a_bits, mag_c = c_box_magic[r].c_bits;
    const uint64_t mag_a2 = pack_f2(mag_a, mag_a), mag_c2 = pack_f2(mag_c, mag_c);

    if (producer) {
        // ==================================== producer warp ====================================
        // Buffer index i of the staged row holds image-row byte B0 + i.  direct: indices [0, sh + 1920) are real data
        // (clamp-to-edge outside the image).  scan: indices [0, sh) are zero (the leaving bytes of the warm-up zone),
        // [sh, sh + 1920) are real data.
        // kVec: every row and buffer is 16-byte aligned, so the staged row's origin is fixed.  !kVec (rows at any byte
        // alignment: odd pitches, unaligned base pointers): the origin moves with every row so that 16-byte chunks of
        // global memory land on 16-byte chunks of shared memory (skew = (row address + B0) mod 16, see stage_row_any).
        const int skew = kVec ? (int)(((B0 % 16) + 16) % 16) : 0;
        uint8_t* const stage_base = smem + (size_t)warp * tl.stage_row;
        const uint32_t stage_base_s = smem_addr(stage_base);
        uint8_t* my_row = stage_base + 16 + skew;                             // buffer index 0 of this warp's staged row
        uint32_t my_row_s = smem_addr(my_row);
        const int64_t first = kDirect ? B0 : e0;           // first image-row position the recurrence reads
        const int first_idx = kDirect ? 0 : sh;
        const int64_t lo = first < 0 ? 0 : first;
        int64_t hi = e0 + kWarpRun; if (hi > pitch) hi = pitch;
        // Copy plan: image-row bytes [cs, ce) land at buffer index (pos - B0); whole 16-byte chunks when kVec.
        int64_t cs = lo, ce = hi;
        if (kVec) {
            cs = lo & ~int64_t(15);
            ce = (hi + 15) & ~int64_t(15); if (ce > pitch) ce = pitch;
            if (cs > ce) cs = ce;
        }
        const int ncopy = ce > cs ? (int)(ce - cs) : 0;
        const int nzero = (!kDirect && kVec && lo > cs) ? (int)(lo - cs) : 0;   // scan: bytes the aligned copy drops on the zero prefix
        const uint32_t copy_dst = my_row_s + (uint32_t)(int)(cs - B0) + 16u * lane;
        // clamp-to-edge: positions [first, 0) (strip 0 only) and [pitch, pitch + rC) that fall inside the run;
        // the buffer origin and pitch are multiples of C there, so the channel of a replicated byte is its offset mod C.
        const int nleft = (first < 0) ? (int)(-first) : 0;                 // buffer indices [first_idx, first_idx + nleft)
        int nright = 0;                                                    // buffer indices [right_idx, +nright)
        if (e0 + kWarpRun > pitch) {
            const int64_t over = e0 + kWarpRun - pitch;
            nright = (int)(over < (int64_t)r * C ? over : (int64_t)r * C);
        }
        const int left_idx = (int)(0 - B0);                // buffer index of image-row byte 0
        const int right_idx = (int)(pitch - B0);
        const bool edge_strip = nleft > 0 || nright > 0;
        const int64_t lane_off = cs + 16 * lane;

        if (kVec && !kDirect)      // zero prefix: written once; the copies never touch it (but see nzero)
            for (int i = lane; i < sh; i += 32) my_row[i] = 0;

        // Row pointers advance by K rows inside the band's own memory and are recomputed at the seams (clamped
        // rows at the image top / bottom, halo rows that live in a neighbour's buffer).  rel = row index relative to Ystart.
        const int64_t own_lo = job.src.band_y0 > 0 ? job.src.band_y0 : 0;
        const int64_t own_hi = job.src.band_y1 < job.height ? job.src.band_y1 : job.height;
        const int rel_fast_lo = (int)(own_lo + K - Ystart);       // row rel and row rel - K both lie in the band's own memory
        const int rel_fast_hi = (int)(own_hi - Ystart);
        const int64_t step_bytes = (int64_t)K * pitch;
        const uint8_t* gsrc = nullptr;                     // this lane's first chunk of the row staged last
        const bool c0 = 16 * lane < ncopy, c1 = 16 * lane + 512 < ncopy, c2 = 16 * lane + 1024 < ncopy,
                   c3 = 16 * lane + 1536 < ncopy;
        auto stage_row = [&](int rel) {                    // this warp: rel = warp mod K
            if (rel < nrows_in) {
                if (rel >= rel_fast_lo && rel < rel_fast_hi && gsrc != nullptr) gsrc += step_bytes;
                else gsrc = job.src.row(clamp64(Ystart + rel, 0, job.height - 1), img) + lane_off;
                {
                    if (c0) cp_async16(copy_dst, gsrc);
                    if (c1) cp_async16(copy_dst + 512, gsrc + 512);
                    if (c2) cp_async16(copy_dst + 1024, gsrc + 1024);
                    if (c3) cp_async16(copy_dst + 1536, gsrc + 1536);
                }
            }
            cp_async_commit();
        };
        // Rows at any alignment.  The whole 16-byte chunks of global memory that lie inside the row and touch [lo, hi) are
        // copied with cp.async; the (at most 15 + 15, without a whole chunk at most 30) bytes of [lo, hi) before the
        // first and after the last such chunk are loaded here, one per lane, and stored when the row is consumed.
        // Strips that stay 15 bytes away from both row ends (`inner`) have neither and skip the clipping.
        int st_skew = 0, st_nhead = 0, st_head_idx = 0, st_ntail = 0, st_tail_idx = 0;
        uint32_t st_hb = 0, st_tb = 0;
        const bool inner = lo >= 15 && hi + 15 <= pitch;
        const int lo_i = (int)lo, hi_i = (int)hi, B0_i = (int)B0;      // row positions fit 31 bits (checked on the host)
        const uint8_t* rowp = nullptr;
        auto stage_row_any = [&](int rel) {
            if (rel < nrows_in) {
                if (rel >= rel_fast_lo && rel < rel_fast_hi && rowp != nullptr) rowp += step_bytes;
                else rowp = job.src.row(clamp64(Ystart + rel, 0, job.height - 1), img);
                const int a = (int)((uintptr_t)rowp & 15);
                st_skew = (a + B0_i) & 15;
                int p_lo = ((a + lo_i) & ~15) - a, p_hi = ((a + hi_i + 15) & ~15) - a;   // row positions the chunks cover
                if (!inner) {
                    if (p_lo < 0) p_lo += 16;
                    if (p_hi > (int)pitch) p_hi -= 16;
                    if (p_hi < p_lo) p_hi = p_lo;
                    const int h_end = hi_i < p_lo ? hi_i : p_lo, t_beg = lo_i > p_hi ? lo_i : p_hi;
                    st_nhead = h_end > lo_i ? h_end - lo_i : 0;
                    st_ntail = hi_i > t_beg ? hi_i - t_beg : 0;
                    st_head_idx = lo_i - B0_i;
                    st_tail_idx = t_beg - B0_i;
                    if (lane < st_nhead) st_hb = rowp[lo_i + lane];
                    if (lane < st_ntail) st_tb = rowp[t_beg + lane];
                }
                const int nch = (p_hi - p_lo) >> 4;
                const uint8_t* src = rowp + p_lo + 16 * lane;
                const uint32_t dst = stage_base_s + 16u + (uint32_t)(st_skew + p_lo - B0_i) + 16u * lane;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (lane + 32 * q < nch) cp_async16(dst + 512 * q, src + 512 * q);
            }
            cp_async_commit();
        };
        auto stage_next = [&](int rel) {
            if (kVec) stage_row(rel); else stage_row_any(rel);
        };

        const uint32_t ring_lane = ring_s + (uint32_t)((lane - nw) * kLaneBytes);

        // direct mode: the window before a lane's run = its first sh leaving bytes = init_full whole 16-byte groups and
        // a last group of which the bytes under init_mask count
        const int init_full = sh >> 4;
        const bool init_part = (sh & 15) != 0;
        uint32_t init_mask[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            int nb = (sh & 15) - 4 * jj;
            nb = nb < 0 ? 0 : (nb > 4 ? 4 : nb);
            init_mask[jj] = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
        }

        // kRet: store output row `warp` of consumer step s (named barriers 5, 6 = OUT_FULL[s & 1], 7, 8 = OUT_EMPTY[s & 1])
        auto flush = [&](int s) {
            bar_sync(5 + (s & 1), nthreads);
            const int rel = s * K + warp;                      // the input row whose arrival completed this output row
            if (rel >= 2 * r && rel < nrows_in) {
                uint8_t* dst = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0 + (rel - 2 * r)) * pitch + bxs;
                int64_t n = pitch - bxs; if (n > tl.useful) n = tl.useful;
                flush_row_any(obuf_s + (uint32_t)(((s & 1) * K + warp) * kRingPitch), dst, (int)n, lane);
            }
            if (s + 2 < nsteps) bar_arrive(7 + (s & 1), nthreads);
        };
        stage_next(warp);
        int slot = 0;                                      // ring slot of the step's first row
        for (int step = 0; step < nsteps; step++) {
            const int rel0 = step * K;
            if (rel0 + warp < nrows_in) {
                cp_async_wait<0>();
                if (!kVec || !kDirect || edge_strip) __syncwarp();      // the fix-ups below touch bytes other lanes copied
                if (kVec) {
                    if (nzero > 0 && lane < nzero) my_row[sh - nzero + lane] = 0;
                } else {
                    my_row = stage_base + 16 + st_skew;
                    my_row_s = stage_base_s + 16u + (uint32_t)st_skew;
                    if (!inner) {
                        if (lane < st_nhead) my_row[st_head_idx + lane] = (uint8_t)st_hb;
                        if (lane < st_ntail) my_row[st_tail_idx + lane] = (uint8_t)st_tb;
                    }
                    if (!kDirect)          // zero prefix (the chunk copies may have spilled into its last bytes)
                        for (int i = lane; i < sh; i += 32) my_row[i] = 0;
                    if (edge_strip) __syncwarp();
                }
                if (edge_strip) {
                    if (nleft > 0) {       // left image edge: replicate pixel 0
                        uint32_t e = 0;
#pragma unroll
                        for (int c = 0; c < C; c++) e |= (uint32_t)my_row[left_idx + c] << (8 * c);
                        for (int i = lane; i < nleft; i += 32) my_row[first_idx + i] = (uint8_t)(e >> (8 * (i % C)));
                    }
                    if (nright > 0) {      // right image edge: replicate the last pixel
                        uint32_t e = 0;
#pragma unroll
                        for (int c = 0; c < C; c++) e |= (uint32_t)my_row[right_idx - C + c] << (8 * c);
                        for (int i = lane; i < nright; i += 32) my_row[right_idx + i] = (uint8_t)(e >> (8 * (i % C)));
                    }
                }
                __syncwarp();

                // leaving and entering words of this lane's run
                uint32_t Lw[kLaneWords], Ew[kLaneWords];
                const uint32_t aL = my_row_s + (uint32_t)(kLaneBytes * lane);
                const uint32_t aE = aL + (uint32_t)sh;
                if (C == 4 && kVec) {
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
                stage_next(rel0 + K + warp);   // the staged row is in registers: refill it for the next step

                int acc[NACC];
                if (kDirect) {
                    // window sum before the run = the first sh leaving bytes, per channel (as 2^23 + sum)
#pragma unroll
                    for (int c = 0; c < NACC; c++) acc[c] = (int)kBias;
                    auto add_word = [&](int j, uint32_t w) {
                        const uint32_t pa = __byte_perm(w, 0u, 0x4140);          // b0 0 b1 0
                        const uint32_t pb = __byte_perm(w, 0u, 0x4342);          // b2 0 b3 0
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int ch = (C == 1) ? 0 : ((4 * j + k) % C);
                            acc[ch] = dp4a_us(k < 2 ? pa : pb, (k & 1) ? (int)0xFF010000 : 0x0000FF01, acc[ch]);
                        }
                    };
                    // whole 16-byte groups of the window, then the group it ends in (words masked); warp-uniform branches
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (g < init_full) {
#pragma unroll
                            for (int j = 4 * g; j < 4 * g + 4 && j < kLaneWords; j++) add_word(j, Lw[j]);
                        } else if (g == init_full && init_part) {
#pragma unroll
                            for (int j = 4 * g; j < 4 * g + 4 && j < kLaneWords; j++) add_word(j, Lw[j] & init_mask[j - 4 * g]);
                        }
                    }
                } else {
                    // phase A: lane totals of (entering - leaving) per channel
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
                }
                // the recurrence from the true start value: round, store to the ring
                if (tl.decoupled && step >= 2) bar_sync(3 + (step & 1), box_threads(GB));   // consumers are done with step - 2
                if (kDirect || lane >= nw) {
                    int rs = slot + warp; if (rs >= tl.ring_rows) rs -= tl.ring_rows;
                    const uint32_t dst = ring_lane + (uint32_t)(rs * kRingPitch);
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
            } else if (tl.decoupled && step >= 2) {
                bar_sync(3 + (step & 1), box_threads(GB));
            }
            if (tl.decoupled) bar_arrive(1 + (step & 1), box_threads(GB));
            else __syncthreads();      // this step's rows are in the ring
            slot += K; if (slot >= tl.ring_rows) slot -= tl.ring_rows;
            if (kRet && tl.ret && step >= 1) flush(step - 1);
        }
        if (kRet && tl.ret) flush(nsteps - 1);
        if (!tl.decoupled) __syncthreads();          // matches the consumers' last barrier
    } else {
        // ==================================== consumer warp ====================================
        // Thread vt owns the GB-byte column group vt of the strip: one vector LDS per ring row, one vector STG per output
        // row, a warp covers 32 * GB contiguous bytes.
        // kVec: thread vt owns group vt.  Rows at any alignment (!kVec): a warp owns 31 groups and lane 0 repeats the last
        // group of the warp before it, so that every lane >= 1 can store the ALIGNED 8-byte word that ends inside its own
        // bytes (its left neighbour's last bytes come by shuffle) and no word is split between two warps; see
        // store_segment_dup.  8 consumer warps x 31 groups cover the 240 groups of a strip.
        const int nvt = tl.useful / kGroupBytes;
        const int vt = (kVec || kRet) ? tid - 32 * kHWarps : 31 * (warp - kHWarps) + lane - 1;
        const int64_t col = bxs + (int64_t)kGroupBytes * vt;
        int vbytes = 0;
        if (vt >= 0 && vt < nvt && col < pitch) vbytes = (pitch - col >= kGroupBytes) ? kGroupBytes : (int)(pitch - col);
        const bool any_lane = vbytes > 0;
        int seg_lo = 0, seg_hi = 0;                  // valid bytes of this warp's 32-group segment (position 0 = lane 0's first byte)
        if (kMode == 0) {
            const int64_t col0 = col - (int64_t)kGroupBytes * lane;
            int64_t lim = bxs + tl.useful; if (lim > pitch) lim = pitch;
            const int64_t v = lim - col0;
            seg_hi = v < 0 ? 0 : (v > 32 * kGroupBytes ? 32 * kGroupBytes : (int)v);
            seg_lo = (warp == kHWarps) ? kGroupBytes : 0;
        }
        const bool seg_full = seg_lo == 0 && seg_hi == 32 * kGroupBytes;
        const bool any = (kVec || kRet) ? any_lane : seg_hi > seg_lo;
        uint32_t S[kGroupBytes];
#pragma unroll
        for (int i = 0; i < kGroupBytes; i++) S[i] = kBias;
        uint8_t* optr = job.out + img * job.src.image_stride + (Y0 - job.src.band_y0) * pitch + col;  // next output row
        const uint32_t ring_tid = ring_s + (uint32_t)(kGroupBytes * ((vt >= 0 && vt < nvt) ? vt : 0));

        auto v_row = [&](uint32_t a_in, uint32_t a_out, uint32_t a_o, bool leave, bool store) {
            uint32_t iw[GW], ow[GW];
            if (GB == 16) {
                const uint4 in4 = lds128(a_in);
                uint4 out4 = make_uint4(0u, 0u, 0u, 0u);
                if (leave) out4 = lds128(a_out);
                iw[0] = in4.x; iw[1] = in4.y; iw[2 % GW] = in4.z; iw[3 % GW] = in4.w;
                ow[0] = out4.x; ow[1] = out4.y; ow[2 % GW] = out4.z; ow[3 % GW] = out4.w;
            } else {
                const uint2 in2 = lds64(a_in);
                uint2 out2 = make_uint2(0u, 0u);
                if (leave) out2 = lds64(a_out);
                iw[0] = in2.x; iw[1] = in2.y; ow[0] = out2.x; ow[1] = out2.y;
            }
            uint32_t res[GW];
#pragma unroll
            for (int w = 0; w < GW; w++) {
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
                    if (GB == 16) stg128_stream(optr, make_uint4(res[0], res[1], res[2 % GW], res[3 % GW]));
                    else stg64_stream(optr, res[0], res[1]);
                } else if (kRet) {
                    if (tl.ret) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a_o), "r"(res[0]), "r"(res[1 % GW]) : "memory");
                    else if (vbytes == kGroupBytes) stg64_stream(optr, res[0], res[1 % GW]);
                    else for (int b = 0; b < vbytes; b++) optr[b] = (uint8_t)(res[b >> 2] >> (8 * (b & 3)));
                } else {
                    store_segment_dup(optr, res[0], res[1 % GW], lane, seg_lo, seg_hi, seg_full);
                }
                optr += pitch;
            }
        };

        if (!tl.decoupled) __syncthreads();          // step 0 rows are in the ring
        int slot_in = 0;                                            // ring slot of the step's first row (a multiple of K: never wraps inside a step)
        int slot_out = tl.ring_rows - (2 * r + 1);                  // ring slot of (first row - (2r+1))
        for (int step = 0; step < nsteps; step++) {
            const int rel0 = step * K;
            if (tl.decoupled) bar_sync(1 + (step & 1), box_threads(GB));
            if (kRet && tl.ret && step >= 2) bar_sync(7 + (step & 1), nthreads);      // the producers have stored the rows of step - 2
            if (any) {
                uint32_t a_in = ring_tid + (uint32_t)(slot_in * kRingPitch);
                uint32_t a_out = ring_tid + (uint32_t)(slot_out * kRingPitch);
                uint32_t a_o = obuf_s + (uint32_t)((step & 1) * K * kRingPitch + kGroupBytes * vt);   // kRet: output row k of this step
                const int wrap_out = tl.ring_rows - slot_out;          // first k whose leaving slot wraps
                if (rel0 >= 2 * r + 1 && rel0 + K <= nrows_in) {
                    // steady state: every row has a leaving row and produces an output row
                    if (wrap_out >= K) {
#pragma unroll
                        for (int k = 0; k < K; k++) v_row(a_in + k * kRingPitch, a_out + k * kRingPitch, a_o + k * kRingPitch, true, true);
                    } else {
#pragma unroll 1
                        for (int k = 0; k < wrap_out; k++) {
                            v_row(a_in, a_out, a_o, true, true);
                            a_in += kRingPitch; a_out += kRingPitch; a_o += kRingPitch;
                        }
                        a_out -= (uint32_t)ring_bytes;
#pragma unroll 1
                        for (int k = wrap_out; k < K; k++) {
                            v_row(a_in, a_out, a_o, true, true);
                            a_in += kRingPitch; a_out += kRingPitch; a_o += kRingPitch;
                        }
                    }
                } else {
                    const int nk = (nrows_in - rel0 < K) ? nrows_in - rel0 : K;
                    const int k_leave = 2 * r + 1 - rel0;      // rows k >= k_leave have a leaving row
                    const int k_store = 2 * r - rel0;          // rows k >= k_store produce an output row
#pragma unroll 1
                    for (int k = 0; k < nk; k++) {
                        if (k == wrap_out) a_out -= (uint32_t)ring_bytes;
                        v_row(a_in, a_out, a_o, k >= k_leave, k >= k_store);
                        a_in += kRingPitch; a_out += kRingPitch; a_o += kRingPitch;
                    }
                }
            }
            if (kRet && tl.ret) bar_arrive(5 + (step & 1), nthreads);          // this step's output rows are in shared memory
            if (!tl.decoupled) __syncthreads();      // the producers may overwrite this step's leaving rows; step+1 rows are ready
            else if (step + 2 < nsteps) bar_arrive(3 + (step & 1), box_threads(GB));
            slot_in += K; if (slot_in >= tl.ring_rows) slot_in -= tl.ring_rows;
            slot_out += K; if (slot_out >= tl.ring_rows) slot_out -= tl.ring_rows;
        }
    }
}

template <int C, int kMode, bool kDirect, int GB>
cudaError_t launch(const Job& job, const BoxTiling& tl, size_t smem, int64_t tiles, cudaStream_t stream) {
    static std::atomic<bool> attr_set[64];   // per instantiation and per device: the opt-in is a per-device attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_set[dev]) {
        e = cudaFuncSetAttribute(gip_box_fused<C, kMode, kDirect, GB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    gip_box_fused<C, kMode, kDirect, GB><<<(unsigned)tiles, box_threads(GB), smem, stream>>>(job, tl);
    count_launch();
    return cudaGetLastError();
}

// mode: 1 aligned, 2 any alignment with the producers storing, 0 any alignment with the consumers storing
template <int C>
cudaError_t launch_c(const Job& job, const BoxTiling& tl, size_t smem, int64_t tiles, int mode, int gb, bool direct, cudaStream_t stream) {
    if (mode == 1 && gb == 16)
        return direct ? launch<C, 1, true, 16>(job, tl, smem, tiles, stream) : launch<C, 1, false, 16>(job, tl, smem, tiles, stream);
    if (mode == 1) return direct ? launch<C, 1, true, 8>(job, tl, smem, tiles, stream) : launch<C, 1, false, 8>(job, tl, smem, tiles, stream);
    if (mode == 2) return direct ? launch<C, 2, true, 8>(job, tl, smem, tiles, stream) : launch<C, 2, false, 8>(job, tl, smem, tiles, stream);
    return direct ? launch<C, 0, true, 8>(job, tl, smem, tiles, stream) : launch<C, 0, false, 8>(job, tl, smem, tiles, stream);
}

}  // namespace

cudaError_t launch_fast_box(const Job& job, cudaStream_t stream, bool* handled) {
    *handled = false;
    const int r = job.radius, C = job.channels;
    if (r < 0 || r > kMaxFusedRadius) return cudaSuccess;
    const int sms = num_sms();
    if (sms <= 0) return cudaErrorInvalidDevice;
    BoxTiling tl;
    memset(&tl, 0, sizeof(tl));
    const int sh = (2 * r + 1) * C;
    static const int no_direct = [] { const char* e = getenv("GIP_BOX_NO_DIRECT"); return e ? atoi(e) : 0; }();   // A/B runs
    const bool direct = sh <= kLaneBytes && !no_direct;
    tl.nw = direct ? 0 : (sh + kLaneBytes - 1) / kLaneBytes;
    tl.useful = ((32 - tl.nw) * kLaneBytes) & ~15;
    const int64_t pitch = job.src.pitch;
    tl.strips = (int)((pitch + tl.useful - 1) / tl.useful);
    tl.stage_row = (sh + kWarpRun + 63) & ~15;
    tl.ring_rows = (2 * r + 1 + 2 * K + K - 1) / K * K;
    const size_t smem = (size_t)K * tl.stage_row + (size_t)tl.ring_rows * kRingPitch;
    if (smem > (size_t)kSmemLimit) return cudaSuccess;
    const int64_t rows = job.src.band_y1 - job.src.band_y0;
    if (rows > 0x3fffffff || job.height > 0x3fffffff) return cudaSuccess;
    // Row bands.  A tile's time is proportional to its row steps (band rows + 2r halo rows + the pipeline fill), the
    // launch's to the number of waves of resident CTAs: take the band count with the smallest waves x steps.
    const int64_t per_band = (int64_t)tl.strips * job.batch;
    const int64_t resident = (int64_t)sms;
    const int64_t min_rows = 16;   // small images: short bands re-filter more halo rows but the march is latency-bound
    int64_t max_bands = rows / min_rows; if (max_bands < 1) max_bands = 1;
    if (max_bands > 1024) max_bands = 1024;
    int64_t want = 1, best_cost = -1;
    for (int64_t nb = 1; nb <= max_bands; nb++) {
        if (per_band * nb > 0x7fffffff) break;
        const int64_t waves = (per_band * nb + resident - 1) / resident;
        const int64_t steps = (rows + nb - 1) / nb + 2 * r + 2 * K;
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
    if (!vec && pitch > 0x7fff0000) return cudaSuccess;  // the any-alignment path keeps row positions in 32 bits
    static const int coupled_env = [] { const char* e = getenv("GIP_BOX_COUPLED"); return e ? atoi(e) : 0; }();   // A/B runs
    tl.decoupled = coupled_env ? 0 : 1;
    static const int gb_env = [] { const char* e = getenv("GIP_BOX_GB"); return e ? atoi(e) : 0; }();   // A/B runs
    const int gb = (gb_env == 16 && vec) ? 16 : 8;       // rows at any alignment: 8-byte groups only
    // rows at any alignment: the producers store (mode 2) when the double-buffered block of output rows fits next to the ring
    static const int no_ret = [] { const char* e = getenv("GIP_BOX_NO_RETURN"); return e ? atoi(e) : 0; }();   // A/B runs
    const size_t smem_ret = smem + (size_t)2 * K * kRingPitch + 16;
    const bool ret = !vec && tl.decoupled && !no_ret && smem_ret <= (size_t)kSmemLimit;
    const int mode = vec ? 1 : (ret ? 2 : 0);
    tl.ret = (pitch % 8 != 0 || job.src.image_stride % 8 != 0 || (uintptr_t)job.out % 8 != 0) ? 1 : 0;
    const size_t smem_used = ret ? smem_ret : smem;
    cudaError_t err = C == 4 ? launch_c<4>(job, tl, smem_used, tiles, mode, gb, direct, stream)
                    : C == 3 ? launch_c<3>(job, tl, smem_used, tiles, mode, gb, direct, stream)
                             : launch_c<1>(job, tl, smem_used, tiles, mode, gb, direct, stream);
    *handled = (err == cudaSuccess);
    return err;
}

}  // namespace gip
