/* gip_b200.h -- C ABI of libgip_b200.so, the B200 (sm_100a) filter library.
 *
 * Plain C: pointers, sizes, ints.  No C++ or torch types cross this boundary, so the same
 * entry points bind from ctypes (Python), cgo, JNI or N-API.  Every function returns a
 * cudaError_t value as int (0 == cudaSuccess); gip_error_string() names it.
 *
 * What each group replaces in the reference (/root/reference):
 *   gip_gaussian_blur / gip_box_blur / gip_sobel
 *       the three entry points of cuda_lib/include/image_filters.h:46-56, :72-81, :103-111
 *       (implemented at cuda_lib/src/image_filters.cu:679, :945, :1603).  Same argument order
 *       and meaning; `level` is the OptimizationLevel value (image_filters.h:24-29) as int.
 *       include/image_filters.h exports the same three with the reference's C++ linkage.
 *   gip_*_host
 *       the body of the pybind11 wrappers, backend/cuda_bindings/bindings.cpp:12-91, :96-163,
 *       :168-237: host buffer in, host buffer out, allocation + H2D + filter + D2H inside.
 *   gip_*_async, gip_*_band
 *       new: stream-ordered batched launches and row-band launches for the multi-GPU
 *       partitioning (the reference has no batch, stream or multi-GPU API).
 *   gip_gaussian_weights
 *       generateGaussianKernel, cuda_lib/src/image_filters.cu:25-39 (without its printf).
 *
 * In-place calls: d_output may overlap d_input (the reference's blurs tolerate that because they go through a
 * temp image, image_filters.cu:760-880); the input is then copied to stream-ordered scratch first.  Band calls
 * return cudaErrorInvalidValue when d_above / d_below overlap d_output.
 *
 * Image layout everywhere: u8, row-major, interleaved channels,
 *   byte(y, x, c) = base[(y*width + x)*channels + c]           (image_filters.cu:95)
 * channels in {1,3,4}.  A batch is `batch` such images back to back.
 *
 * Levels (values of OptimizationLevel): Gaussian accepts 1 and 3, box and Sobel accept 1 and 2,
 * anything else returns cudaErrorNotSupported (801) exactly like the reference.  Gaussian and
 * box compute the same bytes at either level; Sobel level 1 keeps the gray value in float,
 * level 2 rounds it to u8 before the stencil (image_filters.cu:1443-1444).
 */
#ifndef GIP_B200_H
#define GIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIP_LEVEL_NAIVE 1
#define GIP_LEVEL_SHARED_MEMORY 2
#define GIP_LEVEL_TEXTURE_MEMORY 3
#define GIP_MAX_FUSED_RADIUS 31

/* Same layout as PerformanceMetrics (image_filters.h:17-21). */
typedef struct gip_metrics {
    float time_ms;
    float bandwidth_gbps;
    float fps;
} gip_metrics;

/* ---- synchronous, device pointers: drop-in for image_filters.h --------------------------- */
int gip_gaussian_blur(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
                      float sigma, int radius, int level, gip_metrics* metrics);
int gip_box_blur(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
                 int radius, int level, gip_metrics* metrics);
int gip_sobel(const uint8_t* d_input, uint8_t* d_output, int width, int height, int channels,
              int level, gip_metrics* metrics);

/* ---- stream-ordered, batched, device pointers --------------------------------------------
 * `stream` is a cudaStream_t (NULL = legacy default stream).  Returns after enqueueing. */
int gip_gaussian_blur_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                            int channels, int64_t batch, float sigma, int radius, int level, void* stream);
int gip_box_blur_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                       int channels, int64_t batch, int radius, int level, void* stream);
int gip_sobel_async(const uint8_t* d_input, uint8_t* d_output, int64_t width, int64_t height,
                    int channels, int64_t batch, int level, void* stream);

/* ---- row bands of one tall image (multi-GPU partitioning) ---------------------------------
 * The caller owns rows [band_y0, band_y0+band_rows) of an image of `height` rows in d_band
 * and wants the same rows of the filtered image in d_output (band_rows*width*channels bytes).
 * Rows above / below the band that the stencil needs are read from d_above / d_below:
 *   d_above points at image row (band_y0 - rows_above), rows_above rows, same row pitch;
 *   d_below points at image row (band_y0 + band_rows), rows_below rows.
 * They may be peer-GPU memory mapped into this process (NVLink P2P): the kernel loads the halo
 * rows straight through the pointer, there is no separate exchange step.  Pass NULL/0 at the
 * image's own top / bottom edge (clamp-to-edge for the blurs, zero border for Sobel).
 * rows_above/rows_below must be >= min(radius, rows that exist on that side) (1 for Sobel),
 * else cudaErrorInvalidValue. */
int gip_gaussian_blur_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                           uint8_t* d_output, int64_t width, int64_t height, int channels,
                           int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                           float sigma, int radius, int level, void* stream);
int gip_box_blur_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                      uint8_t* d_output, int64_t width, int64_t height, int channels,
                      int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                      int radius, int level, void* stream);
int gip_sobel_band(const uint8_t* d_band, const uint8_t* d_above, const uint8_t* d_below,
                   uint8_t* d_output, int64_t width, int64_t height, int channels,
                   int64_t band_y0, int64_t band_rows, int64_t rows_above, int64_t rows_below,
                   int level, void* stream);

/* ---- host buffers: what bindings.cpp does around each call -------------------------------
 * h_input/h_output are ordinary host memory, width*height*channels*batch bytes.  The call
 * works through cached device buffers in chunks (row bands of one image, image ranges of a
 * batch): upload, kernel and download of successive chunks overlap on three streams.  Pinned
 * caller memory is used directly; pageable memory is staged through cached pinned buffers by
 * helper threads.  Returns when h_output is complete.
 * metrics->time_ms is kernel-only time like the reference's (the sum over the chunks).
 * Threading: thread-safe; the host entry points of one process share one set of cached buffers, streams and
 * helper threads and take turns on it (one mutex around the whole call).  A single call already keeps both PCIe
 * directions busy (1.4-1.5 x the time of the two copies alone on a 64 MiB image), so concurrent callers -- the REST
 * service's thread pool -- queue at the link either way; the device-pointer entry points (gip_*, gip_*_async,
 * gip_*_band) take no library lock and run concurrently on the caller's streams. */
int gip_gaussian_blur_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                           int channels, int64_t batch, float sigma, int radius, int level,
                           gip_metrics* metrics);
int gip_box_blur_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                      int channels, int64_t batch, int radius, int level, gip_metrics* metrics);
int gip_sobel_host(const uint8_t* h_input, uint8_t* h_output, int64_t width, int64_t height,
                   int channels, int64_t batch, int level, gip_metrics* metrics);

/* ---- pinned host memory -------------------------------------------------------------------------
 * The host entry points above run at PCIe speed when the caller's buffers are page-locked (no staging copy).
 * gip_host_alloc returns such memory (cudaHostAlloc, portable); the `gpu_filters` module hands out result arrays
 * from a pool of these.  Replaces the pageable numpy result of bindings.cpp:77-81. */
int gip_host_alloc(int64_t bytes, void** h_ptr_out);
int gip_host_free(void* h_ptr);

/* ---- peer memory for row bands across processes (CUDA IPC) --------------------------------
 * gip_device_alloc returns plain cudaMalloc memory (framework caching allocators sub-allocate, and CUDA IPC
 * exports whole allocations).  gip_ipc_export writes a 64-byte handle for such a base pointer; another process on
 * the same node opens it with gip_ipc_open and may pass the mapped pointer as d_above/d_below.
 * gip_enable_peer_access enables P2P from the current device to `peer_device`. */
int gip_device_alloc(int64_t bytes, void** d_ptr_out);   /* cudaMalloc: a whole allocation, exportable with gip_ipc_export */
int gip_device_free(void* d_ptr);
int gip_ipc_export(const void* d_ptr, uint8_t handle_out[64]);
int gip_ipc_open(const uint8_t handle[64], void** d_ptr_out);
int gip_ipc_close(void* d_ptr);
int gip_enable_peer_access(int peer_device);

/* ---- utilities ---------------------------------------------------------------------------- */
/* 2*radius+1 float32 weights, bit-identical to the reference's host code (image_filters.cu:25-39). */
int gip_gaussian_weights(float* weights_out, int radius, float sigma);
const char* gip_error_string(int err);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
int64_t gip_launch_count(void);
/* Memory the library retains between calls: (1) the host entry points' pinned + device staging buffers (sized for the
 * largest call so far), (2) per device, a library-owned stream-ordered pool (cudaMemPoolCreate, release threshold
 * unlimited) that serves the scratch of the two-kernel and general paths -- up to 1 GiB per call, kept at its high-water
 * mark.  The device's default mempool is not touched.  gip_release_cache() frees (1) and trims (2) on the current device. */
int gip_release_cache(void);
/* "gip_b200 <version> sm_100a" */
const char* gip_version(void);
/* Choose the kernel family: 0 = automatic (fused fast path when eligible), 1 = force the general
 * two-pass path (used by tests to cross-check the two implementations). Returns the old value. */
int gip_set_path(int path);

#ifdef __cplusplus
}
#endif
#endif /* GIP_B200_H */
