// ubench.cu -- B200 micro-benchmarks that size the filter kernels' instruction and memory budgets.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu && tools/ubench
// (1) issue throughput of the instructions the filters are made of, in lane-ops / clk / SM,
// (2) shared-memory LDS.128 bandwidth, (3) HBM copy bandwidth by mechanism (LDG/STG 128-bit,
// 1-D bulk TMA in + out, read-only, write-only).  Results feed DESIGN.md's instruction roofline.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITER = 2048;
constexpr int THREADS = 512;

#define REP8(M) M(a0) M(a1) M(a2) M(a3) M(a4) M(a5) M(a6) M(a7)

#define DEF_U32_KERNEL(NAME, BODY)                                                           \
__global__ void __launch_bounds__(THREADS) NAME(unsigned* out, long long* cyc, unsigned b, unsigned c) { \
    unsigned a0 = b + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3,          \
             a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;           \
    __syncthreads();                                                                          \
    long long t0 = clock64();                                                                 \
    _Pragma("unroll 4")                                                                       \
    for (int i = 0; i < ITER; i++) { REP8(BODY) }                                             \
    long long t1 = clock64();                                                                 \
    out[blockIdx.x * THREADS + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;           \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                           \
}

#define DEF_U64_KERNEL(NAME, BODY)                                                           \
__global__ void __launch_bounds__(THREADS) NAME(unsigned* out, long long* cyc, unsigned b_, unsigned c_) { \
    unsigned long long b = ((unsigned long long)0x3f800001u << 32) | 0x3f800003u;             \
    unsigned long long c = ((unsigned long long)(0x3a000000u + c_) << 32) | (0x3a000000u + b_); \
    unsigned long long a0 = 0x3f8000003f800000ull + threadIdx.x, a1 = a0 + 11, a2 = a0 + 22, a3 = a0 + 33, \
                       a4 = a0 + 44, a5 = a0 + 55, a6 = a0 + 66, a7 = a0 + 77;                 \
    __syncthreads();                                                                          \
    long long t0 = clock64();                                                                 \
    _Pragma("unroll 4")                                                                       \
    for (int i = 0; i < ITER; i++) { REP8(BODY) }                                             \
    long long t1 = clock64();                                                                 \
    unsigned long long r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;                              \
    out[blockIdx.x * THREADS + threadIdx.x] = (unsigned)(r ^ (r >> 32));                       \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                           \
}

#define B_IADD(x)  asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
#define B_IADD3(x) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
#define B_LOP3(x)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
#define B_PRMT(x)  asm volatile("prmt.b32 %0, %0, %1, 0x3715;" : "+r"(x) : "r"(b));
#define B_SHF(x)   asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(x) : "r"(b));
#define B_IMAD(x)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_IMADHI(x) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_DP4A(x)  asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(b), "r"(c));
#define B_DP4A_CHAIN(x) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_FFMA(x)  asm volatile("{.reg .f32 t; mov.b32 t, %0; fma.rn.f32 t, t, 0f3F800001, 0f3A000000; mov.b32 %0, t;}" : "+r"(x));
#define B_FFMA_REG(x) asm volatile("{.reg .f32 t,u,v; mov.b32 t, %0; mov.b32 u, %1; mov.b32 v, %2; fma.rn.f32 t, t, u, v; mov.b32 %0, t;}" : "+r"(x) : "r"(b), "r"(c));
#define B_FADD_RZ(x) asm volatile("{.reg .f32 t; mov.b32 t, %0; add.rz.f32 t, t, 0f4B000000; mov.b32 %0, t;}" : "+r"(x));
#define B_I2F_U8(x) asm volatile("{.reg .f32 t; cvt.rn.f32.u8 t, %0; mov.b32 %0, t;}" : "+r"(x));
#define B_I2FP(x)  asm volatile("{.reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t;}" : "+r"(x));
#define B_F2I(x)   asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rzi.u32.f32 %0, t;}" : "+r"(x));
#define B_RSQ(x)   asm volatile("{.reg .f32 t; mov.b32 t, %0; rsqrt.approx.f32 t, t; mov.b32 %0, t;}" : "+r"(x));
#define B_SQRT_APPROX(x) asm volatile("{.reg .f32 t; mov.b32 t, %0; sqrt.approx.f32 t, t; mov.b32 %0, t;}" : "+r"(x));
#define B_SQRT_RN(x) asm volatile("{.reg .f32 t; mov.b32 t, %0; sqrt.rn.f32 t, t; mov.b32 %0, t;}" : "+r"(x));
#define B_FMNMX(x) asm volatile("{.reg .f32 t,u; mov.b32 t, %0; mov.b32 u, %1; min.f32 t, t, u; mov.b32 %0, t;}" : "+r"(x) : "r"(b));
#define B_PRMT_IMAD(x) asm volatile("prmt.b32 %0, %0, %1, 0x3715; mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_LOP3_IMAD(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_PRMT_DP4A(x) asm volatile("prmt.b32 %0, %0, %1, 0x3715; dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_IADD3_DP4A(x) asm volatile("add.u32 %0, %0, %1; dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_IMAD_DP4A(x) asm volatile("mad.lo.u32 %0, %0, %1, %2; dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define B_FFMA2(x)    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(b), "l"(c));
#define B_FFMA2_ACC(x) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(b), "l"(c));
#define B_FFMA2_RZ(x) asm volatile("fma.rz.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(b), "l"(c));
#define B_FADD2(x)    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c));
#define B_FFMA2_PRMT(x) asm volatile("{.reg .b32 lo, hi; fma.rn.f32x2 %0, %0, %1, %2; mov.b64 {lo,hi}, %0; prmt.b32 lo, lo, hi, 0x3210; mov.b64 %0, {lo,hi};}" : "+l"(x) : "l"(b), "l"(c));

DEF_U32_KERNEL(k_iadd, B_IADD)
DEF_U32_KERNEL(k_iadd3, B_IADD3)
DEF_U32_KERNEL(k_lop3, B_LOP3)
DEF_U32_KERNEL(k_prmt, B_PRMT)
DEF_U32_KERNEL(k_shf, B_SHF)
DEF_U32_KERNEL(k_imad, B_IMAD)
DEF_U32_KERNEL(k_imadhi, B_IMADHI)
DEF_U32_KERNEL(k_dp4a, B_DP4A)
DEF_U32_KERNEL(k_dp4a_chain, B_DP4A_CHAIN)
DEF_U32_KERNEL(k_ffma_imm, B_FFMA)
DEF_U32_KERNEL(k_ffma_reg, B_FFMA_REG)
DEF_U32_KERNEL(k_fadd_rz, B_FADD_RZ)
DEF_U32_KERNEL(k_i2f_u8, B_I2F_U8)
DEF_U32_KERNEL(k_i2fp, B_I2FP)
DEF_U32_KERNEL(k_f2i, B_F2I)
DEF_U32_KERNEL(k_rsq, B_RSQ)
DEF_U32_KERNEL(k_sqrt_approx, B_SQRT_APPROX)
DEF_U32_KERNEL(k_sqrt_rn, B_SQRT_RN)
DEF_U32_KERNEL(k_fmnmx, B_FMNMX)
DEF_U32_KERNEL(k_prmt_imad, B_PRMT_IMAD)
DEF_U32_KERNEL(k_lop3_imad, B_LOP3_IMAD)
DEF_U32_KERNEL(k_prmt_dp4a, B_PRMT_DP4A)
DEF_U32_KERNEL(k_iadd_dp4a, B_IADD3_DP4A)
DEF_U32_KERNEL(k_imad_dp4a, B_IMAD_DP4A)
DEF_U64_KERNEL(k_ffma2, B_FFMA2)
DEF_U64_KERNEL(k_ffma2_acc, B_FFMA2_ACC)
DEF_U64_KERNEL(k_ffma2_rz, B_FFMA2_RZ)
DEF_U64_KERNEL(k_fadd2, B_FADD2)
DEF_U64_KERNEL(k_ffma2_prmt, B_FFMA2_PRMT)

// shared-memory LDS.128 bandwidth: every lane reads 16 B, conflict-free.
__global__ void __launch_bounds__(THREADS) k_lds128(unsigned* out, long long* cyc, unsigned b, unsigned c) {
    __shared__ uint4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += THREADS) buf[i] = make_uint4(i, b, c, i * 3);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITER; i++) {
        uint4 v = buf[(idx + i * 32) & 2047];
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(unsigned*, long long*, unsigned, unsigned);

static void run_issue(const char* name, kern_t k, int ops_per_body, int nsm, double clk_mhz_hint) {
    const int blocks = nsm * 2;   // 2 x 512 threads per SM: 32 warps, 8 per scheduler
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, (size_t)blocks * THREADS * 4));
    CK(cudaMalloc(&cyc, blocks * 8));
    k<<<blocks, THREADS>>>(out, cyc, 3, 5);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<blocks, THREADS>>>(out, cyc, 3, 5);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks);
    CK(cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const double med = (double)h[blocks / 2];
    const double lane_ops = 2.0 * THREADS * (double)ITER * 8 * ops_per_body;   // per SM
    printf("%-16s %8.1f lane-ops/clk/SM   (median %.0f cyc, event %.3f ms -> %.0f MHz eff)\n", name,
           lane_ops / med, med, ms, med / (ms * 1e3));
    (void)clk_mhz_hint;
    cudaFree(out); cudaFree(cyc);
}

// ---------------- memory ----------------
__global__ void __launch_bounds__(256) k_copy128(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_read128(const uint4* __restrict__ in, unsigned* sink, size_t n) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    unsigned acc = 0;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        acc ^= a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n; i += stride) acc ^= in[i].x;
    if (acc == 0x12345678u) sink[0] = acc;
}
__global__ void __launch_bounds__(256) k_write128(uint4* __restrict__ out, size_t n, unsigned v) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    const uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
    for (; i < n; i += stride) out[i] = x;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// One elected thread per CTA moves CHUNK-byte pieces global -> smem -> global with a STAGES-deep ring.
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(128) k_tma_copy(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t nchunks) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)CHUNK * STAGES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        size_t first = blockIdx.x, step = gridDim.x;
        size_t issued = first; int pi = 0;
        // prologue
        for (int s = 0; s < STAGES && issued < nchunks; s++, issued += step) {
            mbar_expect_tx(&bars[s], CHUNK);
            tma_load_1d(smem + (size_t)s * CHUNK, in + issued * CHUNK, CHUNK, &bars[s]);
        }
        unsigned phase = 0; int s = 0;
        for (size_t c = first; c < nchunks; c += step) {
            mbar_wait(&bars[s], phase);
            tma_store_1d(out + c * CHUNK, smem + (size_t)s * CHUNK, CHUNK);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (issued < nchunks) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem slot readable again
                mbar_expect_tx(&bars[s], CHUNK);
                tma_load_1d(smem + (size_t)s * CHUNK, in + issued * CHUNK, CHUNK, &bars[s]);
                issued += step;
            }
            if (++s == STAGES) { s = 0; phase ^= 1; }
            (void)pi;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static double time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clockRate %d kHz, L2 %d MB, smem/block optin %zu\n", p.name, nsm, p.clockRate,
           p.l2CacheSize >> 20, p.sharedMemPerBlockOptin);
    printf("--- issue throughput (2 x 512 threads per SM, 8 independent chains per thread) ---\n");
    run_issue("IADD", k_iadd, 1, nsm, 0);
    run_issue("IADD+ISUB", k_iadd3, 2, nsm, 0);
    run_issue("LOP3", k_lop3, 1, nsm, 0);
    run_issue("PRMT", k_prmt, 1, nsm, 0);
    run_issue("SHF", k_shf, 1, nsm, 0);
    run_issue("IMAD", k_imad, 1, nsm, 0);
    run_issue("IMAD.HI", k_imadhi, 1, nsm, 0);
    run_issue("IDP4A(acc)", k_dp4a, 1, nsm, 0);
    run_issue("IDP4A(chain)", k_dp4a_chain, 1, nsm, 0);
    run_issue("FFMA imm", k_ffma_imm, 1, nsm, 0);
    run_issue("FFMA reg", k_ffma_reg, 1, nsm, 0);
    run_issue("FADD.RZ imm", k_fadd_rz, 1, nsm, 0);
    run_issue("I2F.U8", k_i2f_u8, 1, nsm, 0);
    run_issue("I2FP.F32.U32", k_i2fp, 1, nsm, 0);
    run_issue("F2I", k_f2i, 1, nsm, 0);
    run_issue("MUFU.RSQ", k_rsq, 1, nsm, 0);
    run_issue("MUFU.SQRT", k_sqrt_approx, 1, nsm, 0);
    run_issue("sqrt.rn", k_sqrt_rn, 1, nsm, 0);
    run_issue("FMNMX", k_fmnmx, 1, nsm, 0);
    run_issue("PRMT+IMAD", k_prmt_imad, 2, nsm, 0);
    run_issue("LOP3+IMAD", k_lop3_imad, 2, nsm, 0);
    run_issue("PRMT+IDP4A", k_prmt_dp4a, 2, nsm, 0);
    run_issue("IADD+IDP4A", k_iadd_dp4a, 2, nsm, 0);
    run_issue("IMAD+IDP4A", k_imad_dp4a, 2, nsm, 0);
    run_issue("FFMA2 (x2 lanes)", k_ffma2, 2, nsm, 0);
    run_issue("FFMA2 acc (x2)", k_ffma2_acc, 2, nsm, 0);
    run_issue("FFMA2.RZ (x2)", k_ffma2_rz, 2, nsm, 0);
    run_issue("FADD2 (x2)", k_fadd2, 2, nsm, 0);
    run_issue("FFMA2+PRMT (3)", k_ffma2_prmt, 3, nsm, 0);
    {
        const int blocks = nsm * 2;
        unsigned* out; long long* cyc;
        CK(cudaMalloc(&out, (size_t)blocks * THREADS * 4)); CK(cudaMalloc(&cyc, blocks * 8));
        k_lds128<<<blocks, THREADS>>>(out, cyc, 1, 2); CK(cudaDeviceSynchronize());
        k_lds128<<<blocks, THREADS>>>(out, cyc, 1, 2); CK(cudaDeviceSynchronize());
        std::vector<long long> h(blocks);
        CK(cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        printf("%-16s %8.1f B/clk/SM\n", "LDS.128", 2.0 * THREADS * ITER * 16.0 / (double)h[blocks / 2]);
        cudaFree(out); cudaFree(cyc);
    }

    printf("--- HBM bandwidth, 1 GiB in + 1 GiB out (GB/s = bytes moved / time) ---\n");
    const size_t bytes = (size_t)1 << 30;
    uint8_t *a, *b; unsigned* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    const size_t n16 = bytes / 16;
    for (int mult : {2, 4, 8, 16}) {
        const int grid = nsm * mult;
        double ms = time_ms([&] { k_copy128<<<grid, 256>>>((const uint4*)a, (uint4*)b, n16); }, 9);
        printf("copy LDG.128/STG.128 grid=%4d: %7.1f GB/s (%.3f ms)\n", grid, 2.0 * bytes / ms / 1e6, ms);
    }
    {
        double ms = time_ms([&] { k_read128<<<nsm * 8, 256>>>((const uint4*)a, sink, n16); }, 9);
        printf("read-only  LDG.128          : %7.1f GB/s (%.3f ms)\n", 1.0 * bytes / ms / 1e6, ms);
        ms = time_ms([&] { k_write128<<<nsm * 8, 256>>>((uint4*)b, n16, 7); }, 9);
        printf("write-only STG.128          : %7.1f GB/s (%.3f ms)\n", 1.0 * bytes / ms / 1e6, ms);
        ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 9);
        printf("cudaMemcpy D2D              : %7.1f GB/s (%.3f ms)\n", 2.0 * bytes / ms / 1e6, ms);
    }
    {
        constexpr int CH = 16384, ST = 4;
        const size_t smem = (size_t)CH * ST + 64;
        CK(cudaFuncSetAttribute(k_tma_copy<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int mult : {1, 2, 3}) {
            double ms = time_ms([&] { k_tma_copy<CH, ST><<<nsm * mult, 128, smem>>>(a, b, bytes / CH); }, 9);
            printf("copy bulk-TMA 16K x4 grid=%4d: %7.1f GB/s (%.3f ms)\n", nsm * mult, 2.0 * bytes / ms / 1e6, ms);
        }
        constexpr int CH2 = 8192, ST2 = 8;
        const size_t smem2 = (size_t)CH2 * ST2 + 64;
        CK(cudaFuncSetAttribute(k_tma_copy<CH2, ST2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        for (int mult : {1, 2, 3}) {
            double ms = time_ms([&] { k_tma_copy<CH2, ST2><<<nsm * mult, 128, smem2>>>(a, b, bytes / CH2); }, 9);
            printf("copy bulk-TMA 8K x8  grid=%4d: %7.1f GB/s (%.3f ms)\n", nsm * mult, 2.0 * bytes / ms / 1e6, ms);
        }
        CK(cudaDeviceSynchronize());
        // verify the TMA copy
        CK(cudaMemset(b, 0, bytes));
        k_tma_copy<CH, ST><<<nsm * 2, 128, smem>>>(a, b, bytes / CH);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> h(1 << 20);
        CK(cudaMemcpy(h.data(), b + bytes - h.size(), h.size(), cudaMemcpyDeviceToHost));
        bool ok = true; for (auto v : h) ok &= (v == 1);
        printf("bulk-TMA copy verified: %s\n", ok ? "yes" : "NO");
    }
    return 0;
}
