/*
 * filters_oracle.c -- CPU restatement of the reference's per-pixel filter math.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpu_image_processing_b200/,
 * include/) may call, link or import this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or the
 * reported CPU baseline.
 *
 * What it restates (all citations are /root/reference/cuda_lib/src/image_filters.cu):
 *   weights            :25-39    host float32 expf, left-to-right sum, divide
 *   Gaussian H / V     :81-103 / :125-143   taps i=-r..r in order, clamp-to-edge,
 *                                 (uchar)(sum + 0.5f), u8 temp between the passes (:759-761)
 *   box H / V          :379-395 / :417-430  float sum (exact integer), sum*inv + 0.5f
 *   Sobel level 1      :1164-1176 (borders -> 0), :1183-1233 (gray), :1238-1313 (colour)
 *   Sobel level 2      :1443-1444 (gray rounded to u8), :1555-1584 (stencil on u8 gray)
 *
 * Floating-point contraction.  The reference is CUDA; nvcc (default -fmad=true) fuses
 * a*b+c.  The fusion pattern below was read off the SASS of the reference compiled
 * with `nvcc -O3 -arch=sm_100a` (cuobjdump -sass):
 *   Gaussian : sum = FFMA(pixel, w, sum)                       (one FFMA per tap)
 *   box      : FFMA(sum, inv, 0.5f),  inv = 1.0f/(float)k      (IEEE division)
 *   gray     : FFMA(B, .114f, FFMA(R, .299f, FMUL(G, .587f)))   (G product is the plain mul)
 *   Sobel    : every tap is a single-rounded add (x1, x2 are exact), row-major order,
 *              mag2 = FFMA(gx, gx, FMUL(gy, gy)), sqrtf correctly rounded
 * This file is compiled with -ffp-contract=off and uses fmaf() exactly where the SASS
 * has an FFMA, so the host compiler cannot add or remove a fusion.
 *
 * Parity pin: tests/test_reference_pin_gpu.py runs the reference's own kernels
 * (oracle/_ref, built from the sources where they lie) on the B200 and compares them with
 * this file byte for byte; tests/golden/ holds hashes minted from that run.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GIPO_API __attribute__((visibility("default")))

static inline int64_t clampi(int64_t v, int64_t lo, int64_t hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

static int pick_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    return nthreads;
#else
    (void)nthreads;
    return 1;
#endif
}

GIPO_API int gipo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* image_filters.cu:25-39 */
GIPO_API void gipo_gaussian_weights(float *kernel, int radius, float sigma) {
    float sum = 0.0f;
    for (int i = -radius; i <= radius; i++) {
        float x = (float)i;
        float value = expf(-(x * x) / (2.0f * sigma * sigma));
        kernel[radius + i] = value;
        sum += value;
    }
    for (int i = 0; i < 2 * radius + 1; i++) kernel[i] /= sum;
}

/* One separable pass over the whole image.  horizontal!=0: taps along x (:81-103),
 * else along y (:125-143).  kernel==NULL selects the box rule (:379-395, :417-430). */
static void blur_pass(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                      const float *kernel, int radius, int horizontal, int nthreads) {
    const int ksize = 2 * radius + 1;
    const float inv = 1.0f / (float)ksize;
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t y = 0; y < h; y++) {
        for (int64_t x = 0; x < w; x++) {
            for (int ch = 0; ch < c; ch++) {
                float sum = 0.0f;
                for (int i = -radius; i <= radius; i++) {
                    int64_t nx = horizontal ? clampi(x + i, 0, w - 1) : x;
                    int64_t ny = horizontal ? y : clampi(y + i, 0, h - 1);
                    float p = (float)in[(ny * w + nx) * c + ch];
                    if (kernel) sum = fmaf(p, kernel[radius + i], sum);
                    else        sum += p;
                }
                float r = kernel ? (sum + 0.5f) : fmaf(sum, inv, 0.5f);
                out[(y * w + x) * c + ch] = (uint8_t)r;
            }
        }
    }
}

/* image_filters.cu:679-939 (level 1 and level 2 compute the same values) */
GIPO_API int gipo_gaussian_blur(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                                float sigma, int radius, int nthreads) {
    if (!in || !out || w <= 0 || h <= 0 || c <= 0 || radius < 0) return 1;
    nthreads = pick_threads(nthreads);
    float *k = (float *)malloc(sizeof(float) * (size_t)(2 * radius + 1));
    uint8_t *tmp = (uint8_t *)malloc((size_t)(w * h * c));
    if (!k || !tmp) { free(k); free(tmp); return 2; }
    gipo_gaussian_weights(k, radius, sigma);
    blur_pass(in, tmp, w, h, c, k, radius, 1, nthreads);
    blur_pass(tmp, out, w, h, c, k, radius, 0, nthreads);
    free(k); free(tmp);
    return 0;
}

/* image_filters.cu:945-1119, float arithmetic exactly as the reference */
GIPO_API int gipo_box_blur(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                           int radius, int nthreads) {
    if (!in || !out || w <= 0 || h <= 0 || c <= 0 || radius < 0) return 1;
    nthreads = pick_threads(nthreads);
    uint8_t *tmp = (uint8_t *)malloc((size_t)(w * h * c));
    if (!tmp) return 2;
    blur_pass(in, tmp, w, h, c, NULL, radius, 1, nthreads);
    blur_pass(tmp, out, w, h, c, NULL, radius, 0, nthreads);
    free(tmp);
    return 0;
}

/* Same filter with pure integer rounding (2S+k)/(2k); tests prove it equals the float
 * rule for every S in [0,255k], k odd <= 63. */
static void box_pass_int(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                         int radius, int horizontal, int nthreads) {
    const int k = 2 * radius + 1;
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t y = 0; y < h; y++)
        for (int64_t x = 0; x < w; x++)
            for (int ch = 0; ch < c; ch++) {
                int s = 0;
                for (int i = -radius; i <= radius; i++) {
                    int64_t nx = horizontal ? clampi(x + i, 0, w - 1) : x;
                    int64_t ny = horizontal ? y : clampi(y + i, 0, h - 1);
                    s += in[(ny * w + nx) * c + ch];
                }
                out[(y * w + x) * c + ch] = (uint8_t)((2 * s + k) / (2 * k));
            }
}

GIPO_API int gipo_box_blur_int(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                               int radius, int nthreads) {
    if (!in || !out || w <= 0 || h <= 0 || c <= 0 || radius < 0) return 1;
    nthreads = pick_threads(nthreads);
    uint8_t *tmp = (uint8_t *)malloc((size_t)(w * h * c));
    if (!tmp) return 2;
    box_pass_int(in, tmp, w, h, c, radius, 1, nthreads);
    box_pass_int(tmp, out, w, h, c, radius, 0, nthreads);
    free(tmp);
    return 0;
}

/* The reference's float box rounding for one window sum (used by the equivalence test). */
GIPO_API int gipo_box_round_float(int sum, int ksize, int fused) {
    float inv = 1.0f / (float)ksize;
    float s = (float)sum;
    float r = fused ? fmaf(s, inv, 0.5f) : (s * inv + 0.5f);
    return (int)(uint8_t)r;
}

/* gray as the reference's SASS evaluates `0.299f*R + 0.587f*G + 0.114f*B` (:1245) */
static inline float gray_f(const uint8_t *px) {
    float g = 0.587f * (float)px[1];
    g = fmaf((float)px[0], 0.299f, g);
    g = fmaf((float)px[2], 0.114f, g);
    return g;
}

static inline uint8_t sobel_mag(float gx, float gy) {
    float m2 = fmaf(gx, gx, gy * gy);
    float m = sqrtf(m2);
    m = fminf(m, 255.0f);
    return (uint8_t)(m + 0.5f);
}

/* image_filters.cu:1152-1315 (level 1) and :1329-1597 (level 2) */
GIPO_API int gipo_sobel(const uint8_t *in, uint8_t *out, int64_t w, int64_t h, int c,
                        int level, int nthreads) {
    if (!in || !out || w <= 0 || h <= 0 || c <= 0) return 1;
    if (level != 1 && level != 2) return 3;
    nthreads = pick_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t y = 0; y < h; y++) {
        for (int64_t x = 0; x < w; x++) {
            uint8_t *o = out + (y * w + x) * c;
            if (x < 1 || x >= w - 1 || y < 1 || y >= h - 1) {      /* :1164-1176 */
                for (int ch = 0; ch < c; ch++) o[ch] = 0;
                continue;
            }
            float g[3][3];
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const uint8_t *px = in + ((y + dy) * w + (x + dx)) * c;
                    float v;
                    if (c == 1) v = (float)px[0];
                    else {
                        v = gray_f(px);
                        if (level == 2) v = (float)(uint8_t)(v + 0.5f);   /* :1443-1444 */
                    }
                    g[dy + 1][dx + 1] = v;
                }
            /* row-major tap order; x1 and x2 taps are single-rounded adds, x0 taps are no-ops */
            float gx = 0.0f, gy = 0.0f;
            gx = gx - g[0][0];            gy = gy - g[0][0];
                                          gy = gy - 2.0f * g[0][1];
            gx = gx + g[0][2];            gy = gy - g[0][2];
            gx = gx - 2.0f * g[1][0];
            gx = gx + 2.0f * g[1][2];
            gx = gx - g[2][0];            gy = gy + g[2][0];
                                          gy = gy + 2.0f * g[2][1];
            gx = gx + g[2][2];            gy = gy + g[2][2];
            uint8_t e = sobel_mag(gx, gy);
            for (int ch = 0; ch < c; ch++) o[ch] = e;             /* :1311-1313 */
        }
    }
    return 0;
}
