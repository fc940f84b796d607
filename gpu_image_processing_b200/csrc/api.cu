// api.cu -- the C ABI (include/gip_b200.h) and the reference's C++ entry points
// (include/image_filters.h) on top of the kernels.  Validation, job construction, dispatch
// between the fused fast path and the general path, CUDA-event timing, host-buffer staging.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <set>
#include <thread>
#include <vector>

#include "../../include/gip_b200.h"
#include "../../include/image_filters.h"
#include "common.cuh"

namespace gip {

static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_path{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    int n = cache[dev].load();
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cache[dev].store(n);
    }
    return n;
}

static bool verbose() {
    static const bool v = [] { const char* e = getenv("GIP_VERBOSE"); return e && e[0] == '1'; }();
    return v;
}

// image_filters.cu:25-39 -- float32 expf, running float32 sum, divide.  Host libm, like the reference.
static void gaussian_weights_host(float* k, int radius, float sigma) {
    float sum = 0.0f;
    for (int i = -radius; i <= radius; i++) {
        const float x = static_cast<float>(i);
        const float v = expf(-(x * x) / (2.0f * sigma * sigma));
        k[radius + i] = v;
        sum += v;
    }
    for (int i = 0; i < 2 * radius + 1; i++) k[i] /= sum;
}

static int level_ok(FilterKind kind, int level) {
    if (kind == kGaussian) return level == GIP_LEVEL_NAIVE || level == GIP_LEVEL_TEXTURE_MEMORY;
    return level == GIP_LEVEL_NAIVE || level == GIP_LEVEL_SHARED_MEMORY;
}

struct BandArgs {
    const uint8_t* above = nullptr;
    const uint8_t* below = nullptr;
    int64_t y0 = 0, rows = -1, rows_above = 0, rows_below = 0;
};

// Stream-ordered scratch.  The general and two-kernel Gaussian paths allocate up to 1 GiB per call; giving that back to
// the driver at every synchronisation (the default release threshold is 0) costs milliseconds per call, so the scratch
// comes from a pool owned by this library, one per device, whose release threshold is unlimited: it retains its
// high-water mark until gip_release_cache() trims it (documented in gip_b200.h).
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};

static cudaError_t scratch_pool(cudaMemPool_t* out) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pools[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if ((err = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return err;
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        g_pools[dev] = pool;
    }
    *out = g_pools[dev];
    return cudaSuccess;
}

cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t stream) {
    cudaMemPool_t pool = nullptr;
    cudaError_t err = scratch_pool(&pool);
    if (err != cudaSuccess) return err;
    return cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
}

// Validate and enqueue one filter.  All entry points funnel through here.
static cudaError_t enqueue(FilterKind kind, const uint8_t* d_in, uint8_t* d_out, int64_t width,
                           int64_t height, int channels, int64_t batch, float sigma, int radius,
                           int level, const BandArgs* band, cudaStream_t stream) {
    if (!level_ok(kind, level)) {
        if (verbose()) fprintf(stderr, "gip: level %d is not implemented for this filter\n", level);
        return cudaErrorNotSupported;   // image_filters.cu:693-696, :958-961, :1615-1618
    }
    if (!d_in || !d_out || width <= 0 || height <= 0 || batch <= 0) return cudaErrorInvalidValue;
    if (channels != 1 && channels != 3 && channels != 4) return cudaErrorInvalidValue;
    if (kind != kSobel && radius < 0) return cudaErrorInvalidValue;
    if (kind == kGaussian && !(sigma > 0.0f)) return cudaErrorInvalidValue;
    if (kind == kSobel) radius = 1;

    Job job;
    memset(&job, 0, sizeof(job));
    job.width = width; job.height = height; job.channels = channels; job.batch = batch;
    job.radius = radius;
    job.sobel_u8_gray = (kind == kSobel && level == GIP_LEVEL_SHARED_MEMORY) ? 1 : 0;
    job.out = d_out;
    job.src.pitch = width * channels;
    job.src.image_stride = job.src.pitch * height;
    job.src.band = d_in;
    job.src.band_y0 = 0; job.src.band_y1 = height; job.src.above_y0 = 0;
    if (band) {
        if (batch != 1 || band->y0 < 0 || band->rows <= 0 || band->y0 + band->rows > height)
            return cudaErrorInvalidValue;
        const int64_t y1 = band->y0 + band->rows;
        const int64_t need_above = band->y0 < radius ? band->y0 : radius;
        const int64_t need_below = (height - y1) < radius ? (height - y1) : radius;
        if (band->rows_above < need_above || (need_above > 0 && !band->above)) return cudaErrorInvalidValue;
        if (band->rows_below < need_below || (need_below > 0 && !band->below)) return cudaErrorInvalidValue;
        job.src.band_y0 = band->y0; job.src.band_y1 = y1;
        job.src.above = band->above; job.src.above_y0 = band->y0 - band->rows_above;
        job.src.below = band->below;
    }

    float* d_wide = nullptr;
    cudaError_t err = cudaSuccess;
    // In-place calls (d_output overlapping d_input).  The reference's blurs tolerate them because their two passes
    // go through a temp image (image_filters.cu:760-880); the fused kernels here read halo rows that another CTA
    // may already have overwritten, so an overlapping input is first copied to stream-ordered scratch.
    uint8_t* d_copy = nullptr;
    {
        const size_t out_bytes = (size_t)job.src.pitch * (size_t)(job.src.band_y1 - job.src.band_y0) * (size_t)batch;
        const uintptr_t i0 = (uintptr_t)d_in, o0 = (uintptr_t)d_out;
        if (i0 < o0 + out_bytes && o0 < i0 + out_bytes) {
            if ((err = scratch_alloc((void**)&d_copy, out_bytes, stream)) != cudaSuccess) return err;
            if ((err = cudaMemcpyAsync(d_copy, d_in, out_bytes, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) {
                cudaFreeAsync(d_copy, stream);
                return err;
            }
            job.src.band = d_copy;
        }
        if (band) {     // halo rows that overlap the output cannot be saved the same way: refuse
            const size_t ha = (size_t)job.src.pitch * (size_t)band->rows_above, hb = (size_t)job.src.pitch * (size_t)band->rows_below;
            const uintptr_t a0 = (uintptr_t)band->above, b0 = (uintptr_t)band->below;
            if ((band->above && a0 < o0 + out_bytes && o0 < a0 + ha) || (band->below && b0 < o0 + out_bytes && o0 < b0 + hb)) {
                if (d_copy) cudaFreeAsync(d_copy, stream);
                return cudaErrorInvalidValue;
            }
        }
    }
    if (kind == kGaussian) {
        if (radius <= kMaxFusedRadius) {
            gaussian_weights_host(job.weights, radius, sigma);
        } else {
            const size_t n = 2 * (size_t)radius + 1;
            float* h = (float*)malloc(n * sizeof(float));
            if (!h) { if (d_copy) cudaFreeAsync(d_copy, stream); return cudaErrorMemoryAllocation; }
            gaussian_weights_host(h, radius, sigma);
            err = scratch_alloc((void**)&d_wide, n * sizeof(float), stream);
            if (err == cudaSuccess)   // pageable source: staged before the call returns
                err = cudaMemcpyAsync(d_wide, h, n * sizeof(float), cudaMemcpyHostToDevice, stream);
            free(h);
            if (err != cudaSuccess) {
                if (d_wide) cudaFreeAsync(d_wide, stream);
                if (d_copy) cudaFreeAsync(d_copy, stream);
                return err;
            }
        }
    }

    bool handled = false;
    if (g_path.load() == 0 && radius <= kMaxFusedRadius) err = launch_fast(kind, job, stream, &handled);
    if (!handled && err == cudaSuccess) err = launch_general(kind, job, d_wide, stream);
    if (d_wide) cudaFreeAsync(d_wide, stream);
    if (d_copy) cudaFreeAsync(d_copy, stream);
    return err;
}

static void fill_metrics(gip_metrics* m, float ms, FilterKind kind, int64_t bytes) {
    if (!m) return;
    // reference convention: blurs count 4x the image (two passes, :905-906, :1094-1095), Sobel 2x (:1711-1712)
    const double moved = (double)bytes * (kind == kSobel ? 2.0 : 4.0);
    m->time_ms = ms;
    m->bandwidth_gbps = ms > 0.0f ? (float)(moved / (ms / 1000.0) / (1024.0 * 1024.0 * 1024.0)) : 0.0f;
    m->fps = ms > 0.0f ? 1000.0f / ms : 0.0f;
}

// The reference's time_ms is kernel time only (image_filters.cu:804, :893-901).  CUDA loads a kernel's module lazily
// at its first launch (tens of ms of host time between the two timing events), so the first timed call that reaches a
// given kernel variant on a device runs it once untimed.  Kernel choice depends on exactly the fields of the key.
static bool first_use(FilterKind kind, int64_t pitch, int channels, int radius, int level, const void* d_in, const void* d_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t key = (uint64_t)(dev & 63) | (uint64_t)kind << 6 | (uint64_t)(channels & 7) << 8 |
                         (uint64_t)(radius < 0 ? 0 : (radius > 63 ? 63 : radius)) << 11 | (uint64_t)(level & 7) << 17 |
                         (uint64_t)(pitch & 15) << 20 | (uint64_t)((uintptr_t)d_in & 15) << 24 |
                         (uint64_t)((uintptr_t)d_out & 15) << 28 | (uint64_t)(g_path.load() & 1) << 32;
    static std::mutex mu;
    static std::set<uint64_t> seen;
    std::lock_guard<std::mutex> lock(mu);
    return seen.insert(key).second;
}

// Synchronous device-pointer call on the legacy default stream, timed with events like the reference.
static cudaError_t run_sync(FilterKind kind, const uint8_t* d_in, uint8_t* d_out, int width, int height,
                            int channels, float sigma, int radius, int level, gip_metrics* metrics) {
    if (!level_ok(kind, level)) {           // before anything touches the device, like the reference
        if (verbose()) fprintf(stderr, "gip: level %d is not implemented for this filter\n", level);
        return cudaErrorNotSupported;
    }
    if (d_in && d_out && width > 0 && height > 0 &&
        first_use(kind, (int64_t)width * channels, channels, radius, level, d_in, d_out)) {
        const size_t bytes = (size_t)width * channels * (size_t)height;
        const uintptr_t i0 = (uintptr_t)d_in, o0 = (uintptr_t)d_out;
        const bool overlap = i0 < o0 + bytes && o0 < i0 + bytes;          // an in-place call must not run twice
        if (!overlap) {
            cudaError_t werr = enqueue(kind, d_in, d_out, width, height, channels, 1, sigma, radius, level, nullptr, 0);
            if (werr != cudaSuccess) return werr;
        }
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaEventCreate(&e0);
    if (err != cudaSuccess) return err;
    err = cudaEventCreate(&e1);
    if (err != cudaSuccess) { cudaEventDestroy(e0); return err; }
    cudaEventRecord(e0, 0);
    err = enqueue(kind, d_in, d_out, width, height, channels, 1, sigma, radius, level, nullptr, 0);
    cudaEventRecord(e1, 0);
    cudaError_t serr = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (err == cudaSuccess && serr == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) return err;
    if (serr != cudaSuccess) return serr;
    fill_metrics(metrics, ms, kind, (int64_t)width * height * channels);
    if (verbose()) printf("gip: %s %dx%dx%d r=%d level=%d: %.3f ms\n",
                          kind == kGaussian ? "gaussian" : kind == kBox ? "box" : "sobel",
                          width, height, channels, radius, level, ms);
    return cudaSuccess;
}

// ---- host-buffer path: cached pinned + device staging ---------------------------------------
#ifndef GIP_MAX_CHUNKS
#define GIP_MAX_CHUNKS 32
#endif
constexpr int kMaxChunks = GIP_MAX_CHUNKS;

static long env_long(const char* name, long fallback) {
    const char* e = getenv(name);
    const long v = e ? atol(e) : 0;
    return v > 0 ? v : fallback;
}
struct HostCache {
    std::mutex mu;
    int device = -1;
    uint8_t *d_in = nullptr, *d_out = nullptr, *p_in = nullptr, *p_out = nullptr;
    size_t d_cap = 0, p_cap = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;      // upload, kernels, download
    cudaEvent_t up[kMaxChunks] = {}, k0[kMaxChunks] = {}, k1[kMaxChunks] = {}, down[kMaxChunks] = {};
    cudaEvent_t t0 = nullptr;                                          // start of the call on s_in (GIP_VERBOSE trace)

    void release() {
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        if (p_in) cudaFreeHost(p_in);
        if (p_out) cudaFreeHost(p_out);
        d_in = d_out = p_in = p_out = nullptr; d_cap = p_cap = 0;
        for (cudaStream_t* s : {&s_in, &s_k, &s_out})
            if (*s) { cudaStreamDestroy(*s); *s = nullptr; }
        for (int i = 0; i < kMaxChunks; i++) {
            for (cudaEvent_t* e : {&up[i], &k0[i], &k1[i], &down[i]})
                if (*e) { cudaEventDestroy(*e); *e = nullptr; }
        }
        if (t0) { cudaEventDestroy(t0); t0 = nullptr; }
        device = -1;
    }
    cudaError_t ensure(size_t bytes, bool pinned_in, bool pinned_out) {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        if (dev != device) { release(); device = dev; }
        if (!s_out) {
            for (cudaStream_t* s : {&s_in, &s_k, &s_out})
                if (!*s && (err = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)) != cudaSuccess) return err;
            const unsigned flags = verbose() ? cudaEventDefault : cudaEventDisableTiming;
            for (int i = 0; i < kMaxChunks; i++) {
                if (!up[i] && (err = cudaEventCreateWithFlags(&up[i], flags)) != cudaSuccess) return err;
                if (!k0[i] && (err = cudaEventCreate(&k0[i])) != cudaSuccess) return err;
                if (!k1[i] && (err = cudaEventCreate(&k1[i])) != cudaSuccess) return err;
                if (!down[i] && (err = cudaEventCreateWithFlags(&down[i], flags)) != cudaSuccess) return err;
            }
            if (!t0 && (err = cudaEventCreate(&t0)) != cudaSuccess) return err;
        }
        if (bytes > d_cap) {
            if (d_in) cudaFree(d_in);
            if (d_out) cudaFree(d_out);
            d_in = d_out = nullptr; d_cap = 0;
            if ((err = cudaMalloc((void**)&d_in, bytes)) != cudaSuccess) return err;
            if ((err = cudaMalloc((void**)&d_out, bytes)) != cudaSuccess) return err;
            d_cap = bytes;
        }
        if ((pinned_in || pinned_out) && bytes > p_cap) {
            if (p_in) cudaFreeHost(p_in);
            if (p_out) cudaFreeHost(p_out);
            p_in = p_out = nullptr; p_cap = 0;
            if ((err = cudaMallocHost((void**)&p_in, bytes)) != cudaSuccess) return err;
            if ((err = cudaMallocHost((void**)&p_out, bytes)) != cudaSuccess) return err;
            p_cap = bytes;
        }
        return cudaSuccess;
    }
};
static HostCache g_cache;

// Target chunk size of the host pipeline (GIP_HOST_CHUNK_KB, default 4 MB) and the number of helper threads that
// copy between pageable caller memory and the pinned staging buffers (GIP_HOST_THREADS, default 4: half stage the
// input, half drain the output).  Tuning knobs; the defaults are what bench.py / the tests run.
static size_t host_chunk_bytes() {
    static const size_t v = (size_t)env_long("GIP_HOST_CHUNK_KB", 4096) << 10;
    return v;
}
static int host_threads() {
    static const int v = [] {
        const unsigned hw = std::thread::hardware_concurrency();
        long dflt = hw >= 16 ? 8 : (hw >= 8 ? 4 : 2);           // a memcpy thread moves ~8 GB/s; PCIe Gen5 x16 wants ~50
        long t = env_long("GIP_HOST_THREADS", dflt);
        return (int)(t < 2 ? 2 : (t > 16 ? 16 : t));
    }();
    return v;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// One host call, cut into chunks (row bands of one image, or image ranges of a batch).
struct HostPlan {
    FilterKind kind; int64_t width, height; int channels; int64_t batch; float sigma; int radius, level;
    bool by_rows; int halo; int64_t pitch, units, unit_bytes, n;
    const uint8_t* src; uint8_t* dst;          // pinned: the caller's own memory or the staging buffers
    int64_t lo(int64_t k) const { return units * k / n; }
    size_t off(int64_t k) const { return (size_t)(lo(k) * unit_bytes); }
    size_t len(int64_t k) const { return (size_t)((lo(k + 1) - lo(k)) * unit_bytes); }
};

static cudaError_t issue_upload(const HostPlan& p, HostCache& c, int64_t k) {
    cudaError_t e = cudaMemcpyAsync(c.d_in + p.off(k), p.src + p.off(k), p.len(k), cudaMemcpyHostToDevice, c.s_in);
    cudaEventRecord(c.up[k], c.s_in);
    return e;
}

// chunk k: kernel on s_k once the uploads it reads are done, download on s_out once the kernel is done.
// Three streams, so the download of chunk k-1 runs under the kernel of chunk k and under the upload of k+2.
// (Measured alternatives, tools/zero_copy.py: kernels storing straight into the pinned output reach the DMA's
// 51 GB/s only as one whole-image launch; as per-chunk band launches they are slower than this staged form.)
static cudaError_t issue_compute(const HostPlan& p, HostCache& c, int64_t k) {
    const int64_t u0 = p.lo(k), u1 = p.lo(k + 1);
    uint8_t* const out_base = c.d_out;
    cudaStreamWaitEvent(c.s_k, c.up[k + 1 < p.n ? k + 1 : k], 0);   // the halo rows below live in chunk k+1
    cudaEventRecord(c.k0[k], c.s_k);
    cudaError_t e;
    if (p.by_rows) {
        BandArgs b;
        b.y0 = u0; b.rows = u1 - u0;
        b.rows_above = u0 < p.halo ? u0 : p.halo;
        b.rows_below = (p.height - u1) < p.halo ? (p.height - u1) : p.halo;
        b.above = b.rows_above ? c.d_in + (u0 - b.rows_above) * p.pitch : nullptr;
        b.below = b.rows_below ? c.d_in + u1 * p.pitch : nullptr;
        e = enqueue(p.kind, c.d_in + u0 * p.pitch, out_base + u0 * p.pitch, p.width, p.height, p.channels, 1, p.sigma,
                    p.radius, p.level, p.n > 1 ? &b : nullptr, c.s_k);
    } else {
        e = enqueue(p.kind, c.d_in + u0 * p.unit_bytes, out_base + u0 * p.unit_bytes, p.width, p.height, p.channels,
                    u1 - u0, p.sigma, p.radius, p.level, nullptr, c.s_k);
    }
    cudaEventRecord(c.k1[k], c.s_k);
    if (e != cudaSuccess) return e;
    cudaStreamWaitEvent(c.s_out, c.k1[k], 0);
    e = cudaMemcpyAsync(p.dst + p.off(k), c.d_out + p.off(k), p.len(k), cudaMemcpyDeviceToHost, c.s_out);
    cudaEventRecord(c.down[k], c.s_out);
    return e;
}

// Persistent helper threads of the host path (created on first use, parked on a condition variable between calls:
// spawning and tearing down eight threads per call cost ~0.3 ms, as much as the transfer of a 1080p frame).
class HelperPool {
  public:
    // run fn(j) for j in [0, n) on pool threads; returns at once, wait() blocks until all have finished
    void start(int n, std::function<void(int)> fn) {
        std::unique_lock<std::mutex> lock(mu_);
        while ((int)threads_.size() < n) {
            const int id = (int)threads_.size();
            threads_.emplace_back([this, id] { loop(id); });
        }
        fn_ = std::move(fn);
        n_ = n; pending_ = n; generation_++;
        cv_work_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lock(mu_);
        cv_done_.wait(lock, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~HelperPool() {
        { std::unique_lock<std::mutex> lock(mu_); stop_ = true; cv_work_.notify_all(); }
        for (std::thread& t : threads_) t.join();
    }
  private:
    void loop(int id) {
        long seen = 0;
        for (;;) {
            std::function<void(int)> fn;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [&] { return stop_ || (generation_ != seen && id < n_); });
                if (stop_) return;
                seen = generation_;
                fn = fn_;
            }
            fn(id);
            std::unique_lock<std::mutex> lock(mu_);
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> threads_;
    std::function<void(int)> fn_;
    int n_ = 0, pending_ = 0;
    long generation_ = 0;
    bool stop_ = false;
};
static HelperPool* helper_pool() {
    static HelperPool* pool = new HelperPool();      // leaked on purpose: no thread joins during process teardown
    return pool;
}

// Pageable caller memory: helper threads copy slices of every chunk into / out of the pinned staging buffers
// while the calling thread issues the kernels.  Progress is published through per-chunk atomics.
struct HostProgress {
    std::atomic<int> staged[kMaxChunks];     // slices of chunk k copied into p_in
    std::atomic<int> uploaded[kMaxChunks];   // up[k] has been recorded
    std::atomic<int> issued[kMaxChunks];     // down[k] has been recorded (or the call failed before it)
    std::atomic<int> err{0};
    HostProgress() {
        for (int i = 0; i < kMaxChunks; i++) { staged[i].store(0); uploaded[i].store(0); issued[i].store(0); }
    }
    void fail(cudaError_t e) { int zero = 0; if (e != cudaSuccess) err.compare_exchange_strong(zero, (int)e); }
};
static void slice_of(size_t len, int j, int parts, size_t* a, size_t* b) {
    const size_t step = ((len + parts - 1) / parts + 63) & ~(size_t)63;
    *a = step * j < len ? step * j : len;
    *b = step * (j + 1) < len ? step * (j + 1) : len;
}
static void wait_flag(const std::atomic<int>& f) {
    for (int spin = 0; f.load(std::memory_order_acquire) == 0; spin++)
        if (spin > 64) std::this_thread::yield();
}

// bindings.cpp:37-42, :57-63, :77-81 -- H2D, filter, D2H.  Here: cached device buffers, pinned staging (skipped
// when the caller's memory is already pinned), and the transfer is cut into chunks so that the upload of chunk
// k+2, the kernel of chunk k and the download of chunk k-1 overlap: PCIe runs in both directions at once
// instead of H2D, kernel, D2H back to back.
static cudaError_t run_host(FilterKind kind, const uint8_t* h_in, uint8_t* h_out, int64_t width,
                            int64_t height, int channels, int64_t batch, float sigma, int radius,
                            int level, gip_metrics* metrics) {
    if (!level_ok(kind, level)) return cudaErrorNotSupported;
    if (!h_in || !h_out || width <= 0 || height <= 0 || batch <= 0) return cudaErrorInvalidValue;
    if (channels != 1 && channels != 3 && channels != 4) return cudaErrorInvalidValue;
    if (kind != kSobel && radius < 0) return cudaErrorInvalidValue;
    HostPlan p;
    p.kind = kind; p.width = width; p.height = height; p.channels = channels; p.batch = batch;
    p.sigma = sigma; p.radius = radius; p.level = level;
    p.halo = kind == kSobel ? 1 : radius;
    p.pitch = width * channels;
    const size_t bytes = (size_t)p.pitch * height * batch;
    std::lock_guard<std::mutex> lock(g_cache.mu);
    const bool pin_in = is_pinned(h_in), pin_out = is_pinned(h_out);
    cudaError_t err = g_cache.ensure(bytes, !pin_in, !pin_out);
    if (err != cudaSuccess) return err;
    HostCache& c = g_cache;

    // chunks: units are rows (one image) or images (a batch)
    p.by_rows = batch == 1;
    p.units = p.by_rows ? height : batch;
    p.unit_bytes = p.by_rows ? p.pitch : p.pitch * height;
    int64_t n = (int64_t)(bytes / host_chunk_bytes());
    if (n > kMaxChunks) n = kMaxChunks;
    if (n > p.units) n = p.units;
    if (p.by_rows && p.halo > 0 && n > 1 && p.units / n < 4 * (int64_t)p.halo) n = p.units / (4 * (int64_t)p.halo);
    if (n < 1) n = 1;
    p.n = n;
    p.src = pin_in ? h_in : c.p_in;
    p.dst = pin_out ? h_out : c.p_out;

    if (first_use(kind, p.pitch, channels, radius, level, c.d_in, c.d_out)) {
        // untimed first launch of this kernel variant (module load): a few rows of the staging buffer, whatever they hold
        const int64_t wrows = height < 2 * (int64_t)p.halo + 8 ? height : 2 * (int64_t)p.halo + 8;
        err = enqueue(kind, c.d_in, c.d_out, width, wrows, channels, 1, sigma, radius, level, nullptr, c.s_k);
        if (err == cudaSuccess) err = cudaStreamSynchronize(c.s_k);
        if (err != cudaSuccess) return err;
    }
    const bool threaded = (!pin_in || !pin_out) && bytes >= ((size_t)2 << 20);
    if (verbose()) cudaEventRecord(c.t0, c.s_in);
    double host_us[kMaxChunks + 1] = {};
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_now = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count(); };
    if (!threaded) {
        for (int64_t k = 0; k < n && err == cudaSuccess; k++) {
            if (!pin_in) memcpy(c.p_in + p.off(k), h_in + p.off(k), p.len(k));
            host_us[k] = host_now();
            err = issue_upload(p, c, k);
            if (err == cudaSuccess && k >= 1) err = issue_compute(p, c, k - 1);
        }
        if (err == cudaSuccess) err = issue_compute(p, c, n - 1);
        host_us[n] = host_now();
        if (err != cudaSuccess) {
            cudaStreamSynchronize(c.s_in); cudaStreamSynchronize(c.s_k); cudaStreamSynchronize(c.s_out);
            return err;
        }
        for (int64_t k = 0; k < n; k++) {                          // drain in order: chunk k while k+1 is in flight
            if ((err = cudaEventSynchronize(c.down[k])) != cudaSuccess) return err;
            if (!pin_out) memcpy(h_out + p.off(k), c.p_out + p.off(k), p.len(k));
        }
    } else {
        HostProgress prog;
        const int dev = c.device;
        // staging threads: split between the two directions, or all on the one direction that needs them
        const int t_in = pin_in ? 0 : (pin_out ? host_threads() : host_threads() / 2);
        const int t_out = pin_out ? 0 : (pin_in ? host_threads() : host_threads() - host_threads() / 2);
        // helper j < t_in stages slice j of every chunk (the last one in uploads); helper j >= t_in drains slice j - t_in
        helper_pool()->start(t_in + t_out, [&, t_in, t_out, dev, n](int j) {
            cudaSetDevice(dev);
            if (j < t_in) {
                for (int64_t k = 0; k < n; k++) {
                    size_t a, b;
                    slice_of(p.len(k), j, t_in, &a, &b);
                    if (b > a) memcpy(c.p_in + p.off(k) + a, h_in + p.off(k) + a, b - a);
                    if (prog.staged[k].fetch_add(1, std::memory_order_acq_rel) == t_in - 1) {
                        if (k > 0) wait_flag(prog.uploaded[k - 1]);          // keep s_in in chunk order
                        host_us[k] = host_now();                             // (GIP_VERBOSE timeline: staged, upload issued)
                        prog.fail(issue_upload(p, c, k));
                        prog.uploaded[k].store(1, std::memory_order_release);
                    }
                }
                // Staging done: keep polling until the call's last download has landed.  Measured on the B200 box (64 MiB
                // image, 8 helpers): when the helpers go to sleep as soon as the last chunk is staged, the remaining 3-4
                // chunks crawl (4 MB copies take 0.6-1.0 ms instead of 0.1, band kernels 250 us instead of 40) and the call
                // takes 3.6 ms; with the helpers polling it takes 2.4 ms (the pinned-buffer call: 2.05 ms).  4 pollers
                // recover half of it.  GIP_HOST_SPIN=0 turns the polling off.
                static const int spin_env = [] { const char* e = getenv("GIP_HOST_SPIN"); return e ? atoi(e) : 1; }();
                if (spin_env) {
                    wait_flag(prog.issued[n - 1]);
                    while (prog.err.load() == 0 && cudaEventQuery(c.down[n - 1]) == cudaErrorNotReady) {}
                }
            } else {
                const int jo = j - t_in;
                for (int64_t k = 0; k < n; k++) {
                    wait_flag(prog.issued[k]);
                    if (prog.err.load() != 0) continue;
                    cudaError_t e = cudaEventSynchronize(c.down[k]);
                    if (e != cudaSuccess) { prog.fail(e); continue; }
                    size_t a, b;
                    slice_of(p.len(k), jo, t_out, &a, &b);
                    if (b > a) memcpy(h_out + p.off(k) + a, c.p_out + p.off(k) + a, b - a);
                }
            }
        });
        for (int64_t k = 0; k < n; k++) {
            if (pin_in) {
                prog.fail(issue_upload(p, c, k));
                if (k >= 1 && prog.err.load() == 0) prog.fail(issue_compute(p, c, k - 1));
                if (k >= 1) prog.issued[k - 1].store(1, std::memory_order_release);
            } else {
                wait_flag(prog.uploaded[k + 1 < n ? k + 1 : k]);
                if (prog.err.load() == 0) prog.fail(issue_compute(p, c, k));
                prog.issued[k].store(1, std::memory_order_release);
            }
        }
        if (pin_in) {
            if (prog.err.load() == 0) prog.fail(issue_compute(p, c, n - 1));
            prog.issued[n - 1].store(1, std::memory_order_release);
        }
        helper_pool()->wait();
        cudaError_t e1 = cudaStreamSynchronize(c.s_in), e2 = cudaStreamSynchronize(c.s_k), e3 = cudaStreamSynchronize(c.s_out);
        if (prog.err.load() != 0) return (cudaError_t)prog.err.load();
        if (e1 != cudaSuccess) return e1;
        if (e2 != cudaSuccess) return e2;
        if (e3 != cudaSuccess) return e3;
    }
    float ms_total = 0.0f;
    for (int64_t k = 0; k < n; k++) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, c.k0[k], c.k1[k]);
        ms_total += ms;
        if (verbose()) {                                           // device timeline of the pipeline, ms since the call began
            float a = 0, b = 0, d = 0, e = 0;
            cudaEventElapsedTime(&a, c.t0, c.up[k]); cudaEventElapsedTime(&b, c.t0, c.k0[k]);
            cudaEventElapsedTime(&d, c.t0, c.k1[k]); cudaEventElapsedTime(&e, c.t0, c.down[k]);
            fprintf(stderr, "gip: chunk %2d/%d  host issue %.3f  uploaded %.3f  kernel %.3f..%.3f  downloaded %.3f\n", (int)k, (int)n,
                    host_us[k] * 1e-3, a, b, d, e);
            if (k == n - 1) fprintf(stderr, "gip: all issued at host %.3f ms\n", host_us[n] * 1e-3);
        }
    }
    fill_metrics(metrics, ms_total, kind, (int64_t)bytes);
    return cudaSuccess;
}

}  // namespace gip

using namespace gip;

// ============================== extern "C" ABI ===============================================
extern "C" {

int gip_gaussian_blur(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
                      float sigma, int radius, int level, gip_metrics* metrics) {
    return (int)run_sync(kGaussian, d_input, d_output, width, height, channels, sigma, radius, level, metrics);
}
int gip_box_blur(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
                 int radius, int level, gip_metrics* metrics) {
    return (int)run_sync(kBox, d_input, d_output, width, height, channels, 0.0f, radius, level, metrics);
}
int gip_sobel(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
              int level, gip_metrics* metrics) {
    return (int)run_sync(kSobel, d_input, d_output, width, height, channels, 0.0f, 1, level, metrics);
}

int gip_gaussian_blur_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                            int channels, int64_t batch, float sigma, int radius, int level, void* stream) {
    return (int)enqueue(kGaussian, d_input, d_output, width, height, channels, batch, sigma, radius, level,
                        nullptr, (cudaStream_t)stream);
}
int gip_box_blur_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                       int channels, int64_t batch, int radius, int level, void* stream) {
    return (int)enqueue(kBox, d_input, d_output, width, height, channels, batch, 0.0f, radius, level,
                        nullptr, (cudaStream_t)stream);
}
int gip_sobel_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                    int channels, int64_t batch, int level, void* stream) {
    return (int)enqueue(kSobel, d_input, d_output, width, height, channels, batch, 0.0f, 1, level,
                        nullptr, (cudaStream_t)stream);
}

static BandArgs band_args(const uint8_t* above, const uint8_t* below, int64_t y0, int64_t rows,
                          int64_t ra, int64_t rb) {
    BandArgs b; b.above = above; b.below = below; b.y0 = y0; b.rows = rows; b.rows_above = ra; b.rows_below = rb;
    return b;
}
int gip_gaussian_blur_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                           uint8_t* d_output, int64_t width, int64_t height, int channels,
                           int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                           float sigma, int radius, int level, void* stream) {
    BandArgs b = band_args(d_above, d_below, band_y0, band_rows, rows_above, rows_below);
    return (int)enqueue(kGaussian, d_band, d_output, width, height, channels, 1, sigma, radius, level, &b,
                        (cudaStream_t)stream);
}
int gip_box_blur_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                      uint8_t* d_output, int64_t width, int64_t height, int channels,
                      int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                      int radius, int level, void* stream) {
    BandArgs b = band_args(d_above, d_below, band_y0, band_rows, rows_above, rows_below);
    return (int)enqueue(kBox, d_band, d_output, width, height, channels, 1, 0.0f, radius, level, &b,
                        (cudaStream_t)stream);
}
int gip_sobel_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                   uint8_t* d_output, int64_t width, int64_t height, int channels,
                   int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                   int level, void* stream) {
    BandArgs b = band_args(d_above, d_below, band_y0, band_rows, rows_above, rows_below);
    return (int)enqueue(kSobel, d_band, d_output, width, height, channels, 1, 0.0f, 1, level, &b,
                        (cudaStream_t)stream);
}

int gip_gaussian_blur_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                           int channels, int64_t batch, float sigma, int radius, int level,
                           gip_metrics* metrics) {
    return (int)run_host(kGaussian, h_input, h_output, width, height, channels, batch, sigma, radius, level, metrics);
}
int gip_box_blur_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                      int channels, int64_t batch, int radius, int level, gip_metrics* metrics) {
    return (int)run_host(kBox, h_input, h_output, width, height, channels, batch, 0.0f, radius, level, metrics);
}
int gip_sobel_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                   int channels, int64_t batch, int level, gip_metrics* metrics) {
    return (int)run_host(kSobel, h_input, h_output, width, height, channels, batch, 0.0f, 1, level, metrics);
}

int gip_device_alloc(int64_t bytes, void** d_ptr_out) {
    if (bytes <= 0 || !d_ptr_out) return (int)cudaErrorInvalidValue;
    return (int)cudaMalloc(d_ptr_out, (size_t)bytes);
}
int gip_device_free(void* d_ptr) { return (int)cudaFree(d_ptr); }
int gip_ipc_export(const void* d_ptr, uint8_t handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle_out) return (int)cudaErrorInvalidValue;
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr));
    if (err == cudaSuccess) memcpy(handle_out, &h, 64);
    return (int)err;
}
int gip_ipc_open(const uint8_t handle[64], void** d_ptr_out) {
    if (!handle || !d_ptr_out) return (int)cudaErrorInvalidValue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return (int)cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
}
int gip_ipc_close(void* d_ptr) { return (int)cudaIpcCloseMemHandle(d_ptr); }
int gip_enable_peer_access(int peer_device) {
    cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); err = cudaSuccess; }
    return (int)err;
}

int gip_host_alloc(int64_t bytes, void** h_ptr_out) {
    if (bytes <= 0 || !h_ptr_out) return (int)cudaErrorInvalidValue;
    return (int)cudaHostAlloc(h_ptr_out, (size_t)bytes, cudaHostAllocPortable);
}
int gip_host_free(void* h_ptr) { return (int)cudaFreeHost(h_ptr); }

int gip_gaussian_weights(float* weights_out, int radius, float sigma) {
    if (!weights_out || radius < 0 || !(sigma > 0.0f)) return (int)cudaErrorInvalidValue;
    gaussian_weights_host(weights_out, radius, sigma);
    return 0;
}
const char* gip_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }
int64_t gip_launch_count(void) { return g_launches.load(); }
int gip_release_cache(void) {
    std::lock_guard<std::mutex> lock(g_cache.mu);
    g_cache.release();
    int dev = 0;                                   // also hand the library's scratch pool of this device back to the driver
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> plock(g_pool_mu);
        if (g_pools[dev]) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(g_pools[dev], 0);
        }
    }
    cudaGetLastError();
    return 0;
}
const char* gip_version(void) { return "gip_b200 0.1 sm_100a"; }
int gip_set_path(int path) { return g_path.exchange(path); }

}  // extern "C"

// ============================== reference C++ entry points ====================================
// Same mangled names as cuda_lib/include/image_filters.h:46-112.
cudaError_t gaussianBlur(unsigned char* d_input, unsigned char* d_output, int width, int height,
                         int channels, float sigma, int kernelRadius, OptimizationLevel level,
                         PerformanceMetrics* metrics) {
    static_assert(sizeof(PerformanceMetrics) == sizeof(gip_metrics), "metrics layout");
    return run_sync(kGaussian, d_input, d_output, width, height, channels, sigma, kernelRadius, (int)level,
                    reinterpret_cast<gip_metrics*>(metrics));
}
cudaError_t boxBlur(unsigned char* d_input, unsigned char* d_output, int width, int height,
                    int channels, int kernelRadius, OptimizationLevel level, PerformanceMetrics* metrics) {
    return run_sync(kBox, d_input, d_output, width, height, channels, 0.0f, kernelRadius, (int)level,
                    reinterpret_cast<gip_metrics*>(metrics));
}
cudaError_t sobelEdgeDetection(unsigned char* d_input, unsigned char* d_output, int width, int height,
                               int channels, OptimizationLevel level, PerformanceMetrics* metrics) {
    return run_sync(kSobel, d_input, d_output, width, height, channels, 0.0f, 1, (int)level,
                    reinterpret_cast<gip_metrics*>(metrics));
}
