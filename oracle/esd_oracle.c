/* CPU ORACLE (test infrastructure only) -- plain C restatement of the integer
 * stages of the scene-scoring path, plus the CPU twin of the synthetic clip
 * generator.  Built by oracle/Makefile into oracle/_build/libesd_oracle.so and
 * loaded with ctypes by oracle/c_oracle.py.  Only tests/, smoke() and bench.py's
 * CPU legs may use it; the product (libesd.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" against the scenedetect package (absent from
 * the reference, see oracle/psd_cv2.py); every function here is pinned
 * bit-exact against cv2 4.13.0 by tests/test_oracle_vs_cv2.py.
 *
 * What each function follows (reference = what /root/reference's scene task is
 * specified to run, ml-service/src/services/model_manager.py:715-835 ->
 * PySceneDetect -> OpenCV; arithmetic per SURVEY.md Appendix A):
 *   orc_axis_tables      A.2  OpenCV resize.cpp, INTER_LINEAR coefficient setup
 *   orc_resize_linear    A.2  HResizeLinear + VResizeLinear (uchar,int,short)
 *   orc_bgr2hsv          A.3  OpenCV color_hsv RGB2HSV_b, hrange 180
 *   orc_bgr2y            A.7  OpenCV color_yuv RGB2YCrCb_i Y row (yuv_shift 14)
 *   orc_score_frames     A.4/A.7 integer part: per-frame sum|dHSV| and Y histogram
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../synthclip/synth_core.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ A.2 */
ORC_API void orc_axis_tables(int src, int dst, int32_t* ofs0, int32_t* ofs1, int16_t* c0, int16_t* c1) {
    double inv_scale = (double)dst / (double)src;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        ofs0[d] = s;
        ofs1[d] = s + 1 < src ? s + 1 : src - 1;
        c0[d] = (int16_t)lrintf((1.f - f) * 2048.f); /* cvRound == round-half-even */
        c1[d] = (int16_t)lrintf(f * 2048.f);
    }
}

ORC_API void orc_resize_linear(const uint8_t* src, int sh, int sw, int64_t pitch, uint8_t* dst, int dh, int dw) {
    if (dh == sh && dw == sw) {
        for (int y = 0; y < sh; ++y) memcpy(dst + (size_t)y * dw * 3, src + y * pitch, (size_t)sw * 3);
        return;
    }
    int32_t* xo0 = malloc(sizeof(int32_t) * dw), *xo1 = malloc(sizeof(int32_t) * dw);
    int32_t* yo0 = malloc(sizeof(int32_t) * dh), *yo1 = malloc(sizeof(int32_t) * dh);
    int16_t* a0 = malloc(sizeof(int16_t) * dw), *a1 = malloc(sizeof(int16_t) * dw);
    int16_t* b0 = malloc(sizeof(int16_t) * dh), *b1 = malloc(sizeof(int16_t) * dh);
    orc_axis_tables(sw, dw, xo0, xo1, a0, a1);
    orc_axis_tables(sh, dh, yo0, yo1, b0, b1);
    for (int y = 0; y < dh; ++y) {
        const uint8_t* r0 = src + yo0[y] * pitch;
        const uint8_t* r1 = src + yo1[y] * pitch;
        uint8_t* o = dst + (size_t)y * dw * 3;
        for (int x = 0; x < dw; ++x) {
            for (int c = 0; c < 3; ++c) {
                int h0 = r0[xo0[x] * 3 + c] * a0[x] + r0[xo1[x] * 3 + c] * a1[x];
                int h1 = r1[xo0[x] * 3 + c] * a0[x] + r1[xo1[x] * 3 + c] * a1[x];
                int v = (((b0[y] * (h0 >> 4)) >> 16) + ((b1[y] * (h1 >> 4)) >> 16) + 2) >> 2;
                o[x * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
    free(xo0); free(xo1); free(yo0); free(yo1); free(a0); free(a1); free(b0); free(b1);
}

/* ------------------------------------------------------------------ A.3 */
static int32_t g_sdiv[256], g_hdiv[256];
static int g_tables_ready = 0;

static void orc_init_tables(void) {
    if (g_tables_ready) return;
    g_sdiv[0] = g_hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        g_sdiv[i] = (int32_t)lrint((255 << 12) / (1. * i));
        g_hdiv[i] = (int32_t)lrint((180 << 12) / (6. * i));
    }
    g_tables_ready = 1;
}

ORC_API void orc_hsv_tables(int32_t* sdiv, int32_t* hdiv) {
    orc_init_tables();
    memcpy(sdiv, g_sdiv, sizeof g_sdiv);
    memcpy(hdiv, g_hdiv, sizeof g_hdiv);
}

ORC_API void orc_bgr2hsv(const uint8_t* src, int64_t n_px, uint8_t* dst) {
    orc_init_tables();
    for (int64_t i = 0; i < n_px; ++i) {
        int b = src[3 * i], g = src[3 * i + 1], r = src[3 * i + 2];
        int v = b > g ? b : g; v = v > r ? v : r;
        int m = b < g ? b : g; m = m < r ? m : r;
        int diff = v - m;
        int s = (diff * g_sdiv[v] + (1 << 11)) >> 12;
        int h = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
        h = (h * g_hdiv[diff] + (1 << 11)) >> 12;
        if (h < 0) h += 180;
        dst[3 * i] = (uint8_t)h; dst[3 * i + 1] = (uint8_t)s; dst[3 * i + 2] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------ A.7 */
ORC_API void orc_bgr2y(const uint8_t* src, int64_t n_px, uint8_t* dst) {
    for (int64_t i = 0; i < n_px; ++i) {
        int b = src[3 * i], g = src[3 * i + 1], r = src[3 * i + 2];
        dst[i] = (uint8_t)((4899 * r + 9617 * g + 1868 * b + 8192) >> 14);
    }
}

/* ------------------------------------------------------------------ A.4 + A.7 integer part
 * frames: n frames of sh x sw x 3 (pitch, frame_stride in bytes).  prev_hsv: dh*dw*3 HSV of the
 * frame before frames[0] or NULL (then sums[0] = 0).  On return prev_hsv_out (may alias a caller
 * buffer of dh*dw*3) holds the last frame's HSV.  sums: int64[n][3]; hist: uint32[n][bins] or NULL. */
ORC_API void orc_score_frames(const uint8_t* frames, int64_t n, int sh, int sw, int64_t pitch, int64_t frame_stride,
                              int dh, int dw, const uint8_t* prev_hsv, uint8_t* prev_hsv_out, int64_t* sums,
                              uint32_t* hist, int bins) {
    size_t px = (size_t)dh * dw;
    uint8_t* small = malloc(px * 3);
    uint8_t* hsv[2] = {malloc(px * 3), malloc(px * 3)};
    uint8_t* yb = hist ? malloc(px) : NULL;
    int have_prev = prev_hsv != NULL;
    if (have_prev) memcpy(hsv[1], prev_hsv, px * 3);
    for (int64_t f = 0; f < n; ++f) {
        uint8_t* cur = hsv[f & 1];
        const uint8_t* prv = hsv[(f & 1) ^ 1];
        orc_resize_linear(frames + f * frame_stride, sh, sw, pitch, small, dh, dw);
        orc_bgr2hsv(small, (int64_t)px, cur);
        int64_t s0 = 0, s1 = 0, s2 = 0;
        if (have_prev) {
            for (size_t i = 0; i < px; ++i) {
                s0 += abs((int)cur[3 * i] - (int)prv[3 * i]);
                s1 += abs((int)cur[3 * i + 1] - (int)prv[3 * i + 1]);
                s2 += abs((int)cur[3 * i + 2] - (int)prv[3 * i + 2]);
            }
        }
        sums[3 * f] = s0; sums[3 * f + 1] = s1; sums[3 * f + 2] = s2;
        if (hist) {
            uint32_t* hf = hist + (size_t)f * bins;
            memset(hf, 0, sizeof(uint32_t) * bins);
            orc_bgr2y(small, (int64_t)px, yb);
            for (size_t i = 0; i < px; ++i) hf[((int)yb[i] * bins) >> 8]++;
        }
        have_prev = 1;
    }
    if (prev_hsv_out && n > 0) memcpy(prev_hsv_out, hsv[(n - 1) & 1], px * 3);
    free(small); free(hsv[0]); free(hsv[1]); free(yb);
}

/* ------------------------------------------------------------------ synthetic clip CPU twin
 * Identical bytes to the CUDA filler (both evaluate synth_core.h); caches scene bases; rows are
 * split over a few pthreads (no OpenMP runtime in this image). */
#include <pthread.h>
#include <unistd.h>

typedef struct row_job {
    void (*fn)(void* ctx, int y0, int y1);
    void* ctx;
    int y0, y1;
} row_job;

static void* row_job_main(void* p) {
    row_job* j = (row_job*)p;
    j->fn(j->ctx, j->y0, j->y1);
    return NULL;
}

static void par_rows(void (*fn)(void*, int, int), void* ctx, int H) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt > 16) nt = 16;
    if (nt < 1) nt = 1;
    if (nt > H) nt = H;
    pthread_t th[16];
    row_job jobs[16];
    for (long t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].y0 = (int)((long)H * t / nt); jobs[t].y1 = (int)((long)H * (t + 1) / nt);
        if (t + 1 < nt) pthread_create(&th[t], NULL, row_job_main, &jobs[t]);
    }
    row_job_main(&jobs[nt - 1]);
    for (long t = 0; t + 1 < nt; ++t) pthread_join(th[t], NULL);
}

typedef struct base_cache { int scene; uint8_t* img; } base_cache;
typedef struct base_ctx { uint32_t seed; int scene, W, H; uint8_t* img; } base_ctx;

static void base_rows(void* p, int y0, int y1) {
    base_ctx* b = (base_ctx*)p;
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < b->W; ++x)
            for (int c = 0; c < 3; ++c)
                b->img[((size_t)y * b->W + x) * 3 + c] = (uint8_t)syn_base(b->seed, b->scene, x, y, c, b->W, b->H);
}

static const uint8_t* get_base(base_cache* cache, int* n_cache, uint32_t seed, int scene, int W, int H) {
    for (int i = 0; i < *n_cache; ++i)
        if (cache[i].scene == scene) return cache[i].img;
    int slot;
    if (*n_cache < 4) slot = (*n_cache)++;
    else { /* evict the oldest */
        free(cache[0].img);
        memmove(cache, cache + 1, sizeof(base_cache) * 3);
        slot = 3;
    }
    base_ctx b = {seed, scene, W, H, malloc((size_t)W * H * 3)};
    par_rows(base_rows, &b, H);
    cache[slot].scene = scene;
    cache[slot].img = b.img;
    return b.img;
}

typedef struct fill_ctx { uint32_t seed; int W, H; const syn_frame_desc* d; const uint8_t *A, *B; uint8_t* o; } fill_ctx;

static void fill_rows(void* p, int y0, int y1) {
    fill_ctx* q = (fill_ctx*)p;
    const syn_frame_desc* d = q->d;
    const int W = q->W;
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < W; ++x) {
            int xa = (x + d->pan_a) % W, xb = (x + d->pan_b) % W;
            for (int c = 0; c < 3; ++c) {
                int a = q->A ? q->A[((size_t)y * W + xa) * 3 + c] : 0;
                int b = q->B ? q->B[((size_t)y * W + xb) * 3 + c] : 0;
                q->o[((size_t)y * W + x) * 3 + c] = (uint8_t)syn_clip255(syn_blend(a, b, d) + syn_noise(q->seed, d, x, y, c, W));
            }
        }
}

ORC_API void orc_synth_frames(uint32_t seed, int W, int H, const syn_frame_desc* descs, int64_t n, uint8_t* out) {
    base_cache cache[4];
    int n_cache = 0;
    for (int64_t f = 0; f < n; ++f) {
        const syn_frame_desc* d = descs + f;
        const uint8_t* A = (d->num < d->den) ? get_base(cache, &n_cache, seed, d->scene_a, W, H) : NULL;
        const uint8_t* B = (d->num > 0) ? get_base(cache, &n_cache, seed, d->scene_b, W, H) : NULL;
        if (A) A = get_base(cache, &n_cache, seed, d->scene_a, W, H); /* in case fetching B evicted it */
        fill_ctx q = {seed, W, H, d, A, B, out + (size_t)f * W * H * 3};
        par_rows(fill_rows, &q, H);
    }
    for (int i = 0; i < n_cache; ++i) free(cache[i].img);
}
