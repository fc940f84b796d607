// device_utils.cuh -- small PTX wrappers shared by the fused sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gip {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte asynchronous global -> shared copy (LDGSTS), L2-only caching: the data is streamed once.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// 4-byte variant (L1-allocating form is the only one for sizes below 16)
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// streaming global stores / loads (the images are touched once: keep them out of L1)
__device__ __forceinline__ void stg32_stream(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg128_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// d = c + sum_i a.u8[i] * b.s8[i]   (IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}


// packed float pairs (FFMA2 / FADD2): two lanes per issue slot
__device__ __forceinline__ uint64_t pack_f2(uint32_t lo, uint32_t hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_rz_x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_rn_x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub_rn_x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul_rn_x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma_rn_x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_rz_x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t splat_f2(float x) {
    return pack_f2(__float_as_uint(x), __float_as_uint(x));
}
// u32 -> f32 (I2FP.F32.U32: 64 lanes/clk/SM on the integer side; the FP32 pipe stays free for the taps)
__device__ __forceinline__ uint32_t u2f_bits(uint32_t x) {
    uint32_t r;
    asm("{.reg .f32 t; cvt.rn.f32.u32 t, %1; mov.b32 %0, t;}" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t lo_f2(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi_f2(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


// bytes [k, k+4) of the 8-byte little-endian pair (lo, hi), k = shift_bits/8 in 0..3
__device__ __forceinline__ uint32_t funnel_bytes(uint32_t lo, uint32_t hi, uint32_t shift_bits) {
    return __funnelshift_r(lo, hi, shift_bits);
}

// bytes [a, a + 16) of the 32-byte sequence A B (a = 0..15, the same in every lane: the branches are warp-uniform)
__device__ __forceinline__ uint4 shift_chunk(const uint4& A, const uint4& B, int a) {
    const int ws = a >> 2;
    const uint32_t bs = 8u * (uint32_t)(a & 3);
    uint4 o;
    if (ws == 0) {
        o.x = __funnelshift_r(A.x, A.y, bs); o.y = __funnelshift_r(A.y, A.z, bs); o.z = __funnelshift_r(A.z, A.w, bs); o.w = __funnelshift_r(A.w, B.x, bs);
    } else if (ws == 1) {
        o.x = __funnelshift_r(A.y, A.z, bs); o.y = __funnelshift_r(A.z, A.w, bs); o.z = __funnelshift_r(A.w, B.x, bs); o.w = __funnelshift_r(B.x, B.y, bs);
    } else if (ws == 2) {
        o.x = __funnelshift_r(A.z, A.w, bs); o.y = __funnelshift_r(A.w, B.x, bs); o.z = __funnelshift_r(B.x, B.y, bs); o.w = __funnelshift_r(B.y, B.z, bs);
    } else {
        o.x = __funnelshift_r(A.w, B.x, bs); o.y = __funnelshift_r(B.x, B.y, bs); o.z = __funnelshift_r(B.y, B.z, bs); o.w = __funnelshift_r(B.z, B.w, bs);
    }
    return o;
}

// A warp copies n bytes of a 16-byte aligned shared-memory row to global memory at ANY address: whole 16-byte stores at
// the aligned addresses inside [dst, dst + n) (each is bytes [head + 16 j, +16) of the row: two LDS.128 and four funnel
// shifts), the up to 15 bytes before the first and after the last of them one per lane.  The row must be readable up to
// 16 bytes past n.  n <= 512 kQ (2048 by default).
template <int kQ = 4>
__device__ __forceinline__ void flush_row_any(uint32_t row_s, uint8_t* dst, int n, int lane) {
    int head = (int)((16u - (unsigned)((uintptr_t)dst & 15)) & 15u);
    if (head > n) head = n;
    const int nch = (n - head) >> 4;
    const int tail0 = head + 16 * nch, ntail = n - tail0;
    if (lane < head) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row_s + (uint32_t)lane));
        dst[lane] = (uint8_t)v;
    }
    if (lane < ntail) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row_s + (uint32_t)(tail0 + lane)));
        dst[tail0 + lane] = (uint8_t)v;
    }
    uint8_t* body = dst + head + 16 * lane;
    const uint32_t src = row_s + 16u * (uint32_t)lane;
#pragma unroll
    for (int q = 0; q < kQ; q++) {
        if (lane + 32 * q < nch) {
            const uint4 A = lds128(src + 512u * q), B = lds128(src + 512u * q + 16u);
            stg128_stream(body + 512 * q, shift_chunk(A, B, head));
        }
    }
}

// Store one row segment at any byte alignment.  Lane l of a converged warp holds bytes [8l, 8l + 8) of a segment of
// contiguous output bytes (w0 = the first four, little endian); p_lane = address of the lane's first byte; bytes
// [seg_lo, seg_hi) of the segment exist and are this warp's to write, except that lane 0's own bytes are ALSO held by the
// last lane of the warp to the left (the caller overlaps the warps by one lane).  Every lane >= 1 stores the ALIGNED
// 8-byte word that ends inside its own bytes: its left neighbour's last mo bytes (by shuffle) followed by its own
// first 8 - mo.  Lane 31's last mo bytes are left to the next warp (whose lane 0 repeats them), so no word is ever split
// between warps; only words that straddle seg_lo / seg_hi (the ends of a strip or row) go out byte by byte.
// seg_full = (seg_lo == 0 && seg_hi == 256): no range checks.  Nothing outside [seg_lo, seg_hi) is written.
__device__ __forceinline__ void store_segment_dup(uint8_t* p_lane, uint32_t w0, uint32_t w1, int lane, int seg_lo, int seg_hi,
                                                  bool seg_full) {
    const unsigned mo = (unsigned)((uintptr_t)p_lane & 7);          // the same in every lane
    uint32_t v0 = w0, v1 = w1;
    if (mo != 0) {                                                   // warp-uniform branches
        const uint32_t sh = 8u * ((8u - mo) & 3u);
        const uint32_t b = __shfl_up_sync(0xffffffffu, w1, 1);
        if (mo <= 4) {            // V = bytes [8 - mo, 16 - mo) of the 16-byte sequence a b w0 w1
            v0 = __funnelshift_r(b, w0, sh); v1 = __funnelshift_r(w0, w1, sh);
        } else {
            const uint32_t a = __shfl_up_sync(0xffffffffu, w0, 1);
            v0 = __funnelshift_r(a, b, sh); v1 = __funnelshift_r(b, w0, sh);
        }
    }
    if (lane == 0) return;
    uint8_t* q = p_lane - mo;
    const int s = 8 * lane - (int)mo;                                // segment position of V's first byte
    if (seg_full || (s >= seg_lo && s + 8 <= seg_hi)) {
        asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(q), "r"(v0), "r"(v1) : "memory");
    } else if (s + 8 > seg_lo && s < seg_hi) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (s + k >= seg_lo && s + k < seg_hi) q[k] = (uint8_t)((k < 4 ? v0 : v1) >> (8 * (k & 3)));
    }
}

}  // namespace gip
