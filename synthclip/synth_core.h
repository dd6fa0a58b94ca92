/* Synthetic clip pixel generator -- shared by the CUDA filler (synthclip/synth_fill.cu) and
 * the CPU twin (oracle/esd_oracle.c) so that both produce identical bytes.
 *
 * This is benchmark/test *input* infrastructure, not part
 * of the scoring path: integer-only, counter-based (no RNG state), so any frame
 * range can be generated independently on any device.
 *
 * pixel(t,x,y,c) = clip( blend(A(x+panA,y,c), B(x+panB,y,c), num, den) + noise(t,x,y,c) )
 * A/B are scene base images: a 17x10 grid of hashed BGR anchors, integer
 * bilinear upsampled to WxH, plus a fixed per-scene texture of -4..+3.
 */
#ifndef ESD_SYNTH_CORE_H
#define ESD_SYNTH_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD static inline
#endif

#define SYN_GRID_W 16 /* cells across; 17 anchors */
#define SYN_GRID_H 9  /* cells down;   10 anchors */
#define SYN_SCENE_BLACK (-1)
#define SYN_SCENE_WHITE (-2)

/* one descriptor per frame, built on the host by synthclip/schedule.py */
typedef struct syn_frame_desc {
    int32_t scene_a; /* scene id >= 0, or SYN_SCENE_BLACK / SYN_SCENE_WHITE */
    int32_t scene_b;
    int32_t num; /* blend numerator: 0 -> all A, den -> all B */
    int32_t den; /* >= 1 */
    int32_t pan_a; /* horizontal roll of A in pixels */
    int32_t pan_b;
    int32_t noise_key; /* per-frame noise stream id (normally the frame number) */
    int32_t noise_amp; /* noise in [-amp, amp]; 0 disables */
} syn_frame_desc;

SYN_HD uint32_t syn_hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

SYN_HD uint32_t syn_hash3(uint32_t seed, uint32_t a, uint32_t b) {
    uint32_t h = syn_hash(seed ^ 0x9e3779b9U);
    h = syn_hash(h + a * 0x85ebca6bU + 0x165667b1U);
    h = syn_hash(h ^ (b * 0xc2b2ae35U + 0x27d4eb2fU));
    return h;
}

SYN_HD int syn_clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* anchor colour of scene k at grid point (gx, gy), channel c */
SYN_HD int syn_anchor(uint32_t seed, int scene, int gx, int gy, int c) {
    uint32_t h = syn_hash3(seed, (uint32_t)scene, (uint32_t)(gy * (SYN_GRID_W + 1) + gx));
    return (int)((h >> (8 * c)) & 255U);
}

/* scene base image value (before blend/noise) */
SYN_HD int syn_base(uint32_t seed, int scene, int x, int y, int c, int W, int H) {
    if (scene == SYN_SCENE_BLACK) return 0;
    if (scene == SYN_SCENE_WHITE) return 255;
    int fx = (x * (SYN_GRID_W * 256)) / W;
    int fy = (y * (SYN_GRID_H * 256)) / H;
    int gx = fx >> 8, rx = fx & 255;
    int gy = fy >> 8, ry = fy & 255;
    int g00 = syn_anchor(seed, scene, gx, gy, c);
    int g01 = syn_anchor(seed, scene, gx + 1, gy, c);
    int g10 = syn_anchor(seed, scene, gx, gy + 1, c);
    int g11 = syn_anchor(seed, scene, gx + 1, gy + 1, c);
    int v = (g00 * (256 - rx) * (256 - ry) + g01 * rx * (256 - ry) +
             g10 * (256 - rx) * ry + g11 * rx * ry + 32768) >> 16;
    uint32_t t = syn_hash3(seed ^ 0x5bd1e995U, (uint32_t)scene, (uint32_t)(y * W + x));
    v += (int)((t >> (8 * c)) & 7U) - 4;
    return syn_clip255(v);
}

SYN_HD int syn_noise(uint32_t seed, const syn_frame_desc* d, int x, int y, int c, int W) {
    if (d->noise_amp <= 0) return 0;
    uint32_t h = syn_hash3(seed ^ 0xa511e9b3U, (uint32_t)d->noise_key, (uint32_t)(y * W + x));
    uint32_t span = (uint32_t)(2 * d->noise_amp + 1);
    return (int)(((h >> (8 * c)) & 255U) % span) - d->noise_amp;
}

/* blend of two already-evaluated base values */
SYN_HD int syn_blend(int a, int b, const syn_frame_desc* d) {
    if (d->num <= 0) return a;
    if (d->num >= d->den) return b;
    return ((d->den - d->num) * a + d->num * b + (d->den >> 1)) / d->den;
}

/* full pixel evaluation (used by the CUDA filler; the CPU twin caches bases) */
SYN_HD int syn_pixel(uint32_t seed, const syn_frame_desc* d, int x, int y, int c, int W, int H) {
    int a = 0, b = 0;
    if (d->num < d->den) {
        int xa = x + d->pan_a; xa %= W;
        a = syn_base(seed, d->scene_a, xa, y, c, W, H);
    }
    if (d->num > 0) {
        int xb = x + d->pan_b; xb %= W;
        b = syn_base(seed, d->scene_b, xb, y, c, W, H);
    }
    return syn_clip255(syn_blend(a, b, d) + syn_noise(seed, d, x, y, c, W));
}

#endif /* ESD_SYNTH_CORE_H */
