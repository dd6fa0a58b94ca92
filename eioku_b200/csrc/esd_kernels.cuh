// esd_kernels.cuh -- device code of libesd.so (sm_100a).
//
// Fused scoring kernel (one launch per pushed batch):
//   HBM --cp.async.bulk (TMA 1-D, mbarrier complete_tx)--> smem ring of source-row pairs
//   -> INTER_LINEAR 11-bit fixed-point 2x2 taps (OpenCV resize.cpp arithmetic)
//   -> BGR->HSV uint8 (OpenCV RGB2HSV_b tables, A.3)  [+ BGR->Y and smem-atomic histogram, A.7]
//   -> |cur - prev| against the previous frame's HSV kept in shared memory (A.4)
//   -> warp REDUX -> per-(frame,rowgroup,warp) partial sums.
// A CTA owns a group of destination rows and walks consecutive frames, so the previous
// frame's HSV never leaves the SM; only the touched source rows are ever read from HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_tables.h"  // struct Unit

namespace esd {

#ifndef ESD_CONSUMER_WARPS
#define ESD_CONSUMER_WARPS 8
#endif
constexpr int kConsumerWarps = ESD_CONSUMER_WARPS;
constexpr int kConsumers = kConsumerWarps * 32;  // 256 pixel threads
constexpr int kThreads = kConsumers + 32;        // + 1 producer warp
constexpr int kMaxStages = 8;
constexpr int kMaxRowsPerStage = 4;  // destination rows staged per pipeline slot

enum : int { F_HALO = 1, F_NOPREV = 2, F_CTXPREV = 4, F_SAVE = 8, F_FRAME_END = 16, F_END = 32 };


struct YRow {  // per destination row
    int row0, row1;    // source rows (full layout); NV12: rows of the Y plane
    int crow0, crow1;  // indices into the compact (touched rows only) layout
    uint32_t b0s, b1s; // vertical coefficients, pre-shifted << 16 for mul.hi
    uint32_t uvrows;   // NV12: rows of the interleaved UV plane, row0 >> 1 | (row1 >> 1) << 16
    uint32_t cuvrows;  // NV12: their indices in the compact layout (Y rows first, then UV rows), packed the same way
};

struct FusedParams {
    const uint8_t* src;
    const uint8_t* src_uv;   // NV12: interleaved UV plane of frame 0 (full layout; the compact layout keeps UV rows behind the Y rows)
    long long frame_stride;  // bytes between frames
    long long row_stride;    // bytes between rows (pitch, or row_bytes in the compact layout)
    long long uv_frame_stride, uv_row_stride;  // NV12, full layout: the same two for the UV plane (a decoder surface: equal to Y's)
    int compact;
    int uv_packed_base;  // >= 0: src_uv holds only the touched UV rows (I420 repack, i420_interleave_kernel): row index = cuvrows - this
    int n_frames;
    int dst_w, dst_h;
    int row_bytes;  // bytes of one source row: src_w * 3 (BGR24) or src_w (NV12: Y rows and UV rows alike)
    int rows_per_group, n_groups;
    int stages;
    int rows_per_stage;  // destination rows per pipeline stage (1..kMaxRowsPerStage)
    int rowbuf;  // bytes reserved per staged source row (multiple of 16)
    int has_prev;
    int bins;
    int want_bgr;  // also accumulate sum(B+G+R) per frame (ThresholdDetector's average_rgb)
    int lane_stride;  // 1, 2, 4 or 8: destination columns between neighbouring lanes of a consumer warp (bank-conflict-free taps)
    const YRow* yrows;
    const uint2* xtab;  // per destination column: {byte offset of tap 0, a0 | a1 << 16}
    const int* sdiv;
    const int* hdiv;
    const Unit* units;
    const int* cta_unit_begin;  // [grid + 1]
    const uint32_t* prev_in;    // packed H|S<<8|V<<16 of the frame before this batch [dst_h][dst_w]
    uint32_t* prev_out;         // same, written from the last frame of this batch
    uint4* part;                // [n_frames][n_groups][kConsumerWarps] {sumH, sumS, sumV, 0}
    uint16_t* hist_part;        // [n_frames][n_groups][bins]
    uint8_t* vplane;            // [n_frames][dst_h][dst_w] V plane for the edge detector, or nullptr
    uint8_t* gplane;            // [n_frames][dst_h][dst_w] BGR2GRAY plane for the hash detector, or nullptr
};

// ----------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// 2 us: no effect on the HBM-bound BGR24 kernel (A/B: 0.957 sustained either way) but in the compute-bound variants
// (NV12, full-resolution) the producer warp's empty-barrier spin stops competing with the consumers for issue slots
#ifndef ESD_WAIT_HINT_NS
#define ESD_WAIT_HINT_NS 2000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
#if ESD_WAIT_HINT_NS > 0
    // suspend-time hint: the warp may stay suspended (not issuing) up to this long before try_wait returns false; it still
    // wakes as soon as the phase completes, so unlike a sleep this only removes spin iterations
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"((uint32_t)ESD_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    return done != 0;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a pipeline bug must trap (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;  // the clock is only read on the (never taken in a healthy run) long-wait path
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0) {
            const unsigned long long t = globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ULL) __trap();
        }
    }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk tensor-memory-accelerator copy global -> shared, completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}

// ----------------------------------------------------------------------------------- pixel math
// unaligned little-endian 32-bit read at shared byte address `a` (reads a & ~3 and the next word)
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t* base, uint32_t a, uint32_t& next_word) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (a & ~3u));
    uint32_t w0 = w[0], w1 = w[1];
    next_word = w1;
    return __funnelshift_r(w0, w1, (a & 3u) * 8u);
}

// OpenCV RGB2HSV_b (uint8, hrange 180), packed H | S << 8 | V << 16
__device__ __forceinline__ uint32_t bgr_to_hsv_packed(int b, int g, int r, const int* __restrict__ sdiv,
                                                      const int* __restrict__ hdiv) {
    int v = max(max(b, g), r);
    int m = min(min(b, g), r);
    int diff = v - m;
    int s = (diff * sdiv[v] + (1 << 11)) >> 12;
    // branch-free select (a divergent branch here cost more issue slots than the two spare subtractions)
    const int hr = g - b, hg = b - r + 2 * diff, hb = r - g + 4 * diff;
    int h = (v == g) ? hg : hb;
    h = (v == r) ? hr : h;
    h = (h * hdiv[diff] + (1 << 11)) >> 12;
    h += (h < 0) ? 180 : 0;
    return (uint32_t)h | ((uint32_t)s << 8) | ((uint32_t)v << 16);
}

// cv2.cvtColor(BGR2GRAY) for uint8: 15-bit coefficients (OpenCV RGB2Gray<uchar>), not the 14-bit Y of BGR2YUV
__device__ __forceinline__ uint8_t bgr_to_gray(int b, int g, int r) {
    return (uint8_t)((9798 * r + 19235 * g + 3735 * b + 16384) >> 15);
}

// OpenCV VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>: b0s/b1s are the coefficients << 16
__device__ __forceinline__ int vresize(uint32_t h0, uint32_t h1, uint32_t b0s, uint32_t b1s) {
    // no saturation needed: a0+a1 and b0+b1 are 2048 (+-1), so the sum of the two terms is <= 1021 and
    // (1021 + 2) >> 2 == 255
    return (int)((__umulhi(b0s, h0 >> 4) + __umulhi(b1s, h1 >> 4) + 2u) >> 2);
}

// cv2.cvtColor(COLOR_YUV2BGR_NV12), OpenCV color_yuv (ITU-R BT.601, 20-bit fixed point); verified over all 2^24 (Y, U, V):
//   yy = max(0, Y - 16) * 1220542;  B = sat((yy + buv) >> 20) ...  with the chroma terms below (1 << 19 is the rounding)
struct Chroma { int ruv, guv, buv; };
__device__ __forceinline__ Chroma nv12_chroma(int u, int v) {
    u -= 128; v -= 128;
    Chroma c;
    c.ruv = (1 << 19) + 1673527 * v;
    c.guv = (1 << 19) - 852492 * v - 409993 * u;
    c.buv = (1 << 19) + 2116026 * u;
    return c;
}
__device__ __forceinline__ void nv12_pixel(int y, const Chroma& c, int& b, int& g, int& r) {
    const int yy = __vimax_s32_relu(y - 16, 0) * 1220542;
    b = __vimin_s32_relu((yy + c.buv) >> 20, 255);
    g = __vimin_s32_relu((yy + c.guv) >> 20, 255);
    r = __vimin_s32_relu((yy + c.ruv) >> 20, 255);
}

// One destination row of one frame for one consumer thread (PXT columns).  SPECIAL = the stage carries a rare flag
// (halo frame, first frame of the video / of the batch, last frame of the batch); the common path has none.
// EXTRAS = this launch needs one of the per-pixel side outputs (sum of B+G+R for ThresholdDetector, V plane for the edge
// detector, gray plane for the hash detector); the common launches compile all three out (9-12 instructions per pixel).
template <bool RESIZE, int PXT, bool CONTENT, bool HIST, bool SPECIAL, bool NV12, bool EXTRAS>
__device__ __forceinline__ void score_row(const FusedParams& p, const uint8_t* __restrict__ row0, const uint8_t* __restrict__ row1,
                                          const uint8_t* __restrict__ uv0, const uint8_t* __restrict__ uv1, uint32_t misuv,
                                          uint32_t mis0, uint32_t mis1, uint32_t b0s, uint32_t b1s, int flags, int rloc,
                                          int row, int frame, int tid, int col, const uint32_t (&xoff)[PXT], const uint32_t (&xa01)[PXT],
                                          const int* __restrict__ s_sdiv, const int* __restrict__ s_hdiv,
                                          uint32_t* __restrict__ s_prev, uint32_t* __restrict__ s_hist_cur,
                                          uint32_t& acc_hv, uint32_t& acc_s, uint32_t& acc_bgr) {
#pragma unroll
    for (int k = 0; k < PXT; ++k) {
        const int d = k * kConsumers + col;  // this thread's destination column (per-thread state stays indexed by tid)
        // No branch around the pixel: threads without a column (d >= dst_w) compute on the row's first bytes (their x offset
        // is 0) and only their side effects are predicated off, so the body is straight-line code and two rows can be
        // interleaved by the scheduler (the consumer warps are latency-bound: ~4 warps per scheduler).
        const bool valid = d < p.dst_w;
        {
            int b, g, r;
            if (RESIZE && NV12) {
                // xoff packs, per destination column: byte offset of the two luma taps in a Y row (bits 0-12), byte offset
                // of the chroma pair of tap 0 in a UV row (bits 13-25) and whether tap 1 uses the next pair (bit 26).
                // Full rows: x0 | (x0 & ~1) << 13 | (x0 & 1) << 26; gathered tap rows: 2 d | 4 d << 13 | 1 << 26.
                uint32_t nx;
                const uint32_t x0 = xoff[k] & 0x1fffu;
                const uint32_t ya = lds_u32_unaligned(row0, x0 + mis0, nx);   // [Y(x0), Y(x0+1), ..] of source row 0
                const uint32_t yb = lds_u32_unaligned(row1, x0 + mis1, nx);   // same for source row 1
                const uint32_t co = (xoff[k] >> 13) & 0x1fffu;
                const uint32_t ca = lds_u32_unaligned(uv0, co + (misuv & 0xffu), nx);         // [U V U' V'] for row 0
                const uint32_t sh = (xoff[k] >> 26) * 16u;  // tap 1 belongs to the next chroma pair
                const Chroma c00 = nv12_chroma(ca & 255u, (ca >> 8) & 255u);
                const Chroma c01 = nv12_chroma((ca >> sh) & 255u, (ca >> (sh + 8u)) & 255u);
                Chroma c10 = c00, c11 = c01;
                if (!(misuv & 0x10000u)) {  // warp-uniform: both source rows share one chroma row half of the time
                    const uint32_t cb = lds_u32_unaligned(uv1, co + ((misuv >> 8) & 0xffu), nx);  // [U V U' V'] for row 1
                    c10 = nv12_chroma(cb & 255u, (cb >> 8) & 255u);
                    c11 = nv12_chroma((cb >> sh) & 255u, (cb >> (sh + 8u)) & 255u);
                }
                int b00, g00, r00, b01, g01, r01, b10, g10, r10, b11, g11, r11;
                nv12_pixel(ya & 255u, c00, b00, g00, r00);
                nv12_pixel((ya >> 8) & 255u, c01, b01, g01, r01);
                nv12_pixel(yb & 255u, c10, b10, g10, r10);
                nv12_pixel((yb >> 8) & 255u, c11, b11, g11, r11);
                const uint32_t a0 = xa01[k] & 0xffffu, a1 = xa01[k] >> 16;
                b = vresize(a0 * b00 + a1 * b01, a0 * b10 + a1 * b11, b0s, b1s);
                g = vresize(a0 * g00 + a1 * g01, a0 * g10 + a1 * g11, b0s, b1s);
                r = vresize(a0 * r00 + a1 * r01, a0 * r10 + a1 * r11, b0s, b1s);
            } else if (RESIZE) {
                uint32_t n0, n1;
                const uint32_t o0 = xoff[k] + mis0, o1 = xoff[k] + mis1;
                const uint32_t lo0 = lds_u32_unaligned(row0, o0, n0);
                const uint32_t w02 = *reinterpret_cast<const uint32_t*>(row0 + (o0 & ~3u) + 8);
                const uint32_t hi0 = __funnelshift_r(n0, w02, (o0 & 3u) * 8u);
                const uint32_t lo1 = lds_u32_unaligned(row1, o1, n1);
                const uint32_t w12 = *reinterpret_cast<const uint32_t*>(row1 + (o1 & ~3u) + 8);
                const uint32_t hi1 = __funnelshift_r(n1, w12, (o1 & 3u) * 8u);
                // lo = [Ab Ag Ar Bb], hi = [Bg Br . .]  ->  [Ab Bb Ag Bg] and [Ar Br . .]
                const uint32_t bg0 = __byte_perm(lo0, hi0, 0x4130), rr0 = __byte_perm(lo0, hi0, 0x0052);
                const uint32_t bg1 = __byte_perm(lo1, hi1, 0x4130), rr1 = __byte_perm(lo1, hi1, 0x0052);
                const uint32_t a = xa01[k];
                // HResizeLinear: tap0 * a0 + tap1 * a1 (scale 2^11)
                const uint32_t hb0 = __dp2a_lo(a, bg0, 0u), hg0 = __dp2a_hi(a, bg0, 0u), hr0 = __dp2a_lo(a, rr0, 0u);
                const uint32_t hb1 = __dp2a_lo(a, bg1, 0u), hg1 = __dp2a_hi(a, bg1, 0u), hr1 = __dp2a_lo(a, rr1, 0u);
                b = vresize(hb0, hb1, b0s, b1s);
                g = vresize(hg0, hg1, b0s, b1s);
                r = vresize(hr0, hr1, b0s, b1s);
            } else {
                uint32_t nx;
                const uint32_t px = lds_u32_unaligned(row0, xoff[k] + mis0, nx);
                b = px & 255u;
                g = (px >> 8) & 255u;
                r = (px >> 16) & 255u;
            }
            if (CONTENT) {
                if (EXTRAS && p.want_bgr && valid && !(SPECIAL && (flags & F_HALO))) acc_bgr += (uint32_t)(b + g + r);
                const uint32_t cur = bgr_to_hsv_packed(b, g, r, s_sdiv, s_hdiv);
                uint32_t* slot = s_prev + (rloc * PXT + k) * kConsumers + tid;
                uint32_t pv;
                if (SPECIAL) {
                    pv = cur;
                    if (!(flags & (F_HALO | F_NOPREV)))
                        pv = (flags & F_CTXPREV) ? (valid ? __ldg(p.prev_in + (size_t)row * p.dst_w + d) : cur) : *slot;
                    if ((flags & F_SAVE) && valid) p.prev_out[(size_t)row * p.dst_w + d] = cur;
                } else {
                    pv = *slot;
                }
                if (EXTRAS && p.vplane && valid && !(SPECIAL && (flags & F_HALO)))
                    p.vplane[((size_t)frame * p.dst_h + row) * p.dst_w + d] = (uint8_t)(cur >> 16);
                const uint32_t diff = valid ? __vabsdiffu4(cur, pv) : 0u;
                acc_hv += diff & 0x00ff00ffu;
                acc_s += (diff >> 8) & 0xffu;
                *slot = cur;
            }
            if (HIST && valid && !(SPECIAL && (flags & F_HALO))) {
                const int y = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
                atomicAdd(&s_hist_cur[(y * p.bins) >> 8], 1u);
            }
            if (EXTRAS && p.gplane && valid && !(SPECIAL && (flags & F_HALO)))
                p.gplane[((size_t)frame * p.dst_h + row) * p.dst_w + d] = bgr_to_gray(b, g, r);
        }
    }
}

// Full-resolution rows (no resize) with 16-byte aligned rows: a thread takes four consecutive pixels = three aligned
// words, so there are no funnel shifts, the previous HSV moves as one 16-byte vector and guards are per quad.
template <int QPT, bool CONTENT, bool HIST, bool SPECIAL, bool EXTRAS>
__device__ __forceinline__ void score_row_quads(const FusedParams& p, const uint8_t* __restrict__ row0, int flags, int rloc,
                                                int row, int frame, int tid, const int* __restrict__ s_sdiv,
                                                const int* __restrict__ s_hdiv, uint32_t* __restrict__ s_prev,
                                                uint32_t* __restrict__ s_hist_cur, uint32_t& acc_hv, uint32_t& acc_s,
                                                uint32_t& acc_bgr) {
    const int quads = (p.dst_w + 3) >> 2;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        const int q = j * kConsumers + tid;
        if (q < quads) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(row0) + 3 * q;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            const uint32_t px[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
            const int nvalid = min(4, p.dst_w - 4 * q);
            uint32_t cur[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int b = px[e] & 255u, g = (px[e] >> 8) & 255u, r = (px[e] >> 16) & 255u;
                cur[e] = CONTENT ? bgr_to_hsv_packed(b, g, r, s_sdiv, s_hdiv) : 0u;
                if (EXTRAS && CONTENT && p.want_bgr && e < nvalid && !(SPECIAL && (flags & F_HALO))) acc_bgr += (uint32_t)(b + g + r);
                if (HIST && e < nvalid && !(SPECIAL && (flags & F_HALO))) {
                    const int y = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
                    atomicAdd(&s_hist_cur[(y * p.bins) >> 8], 1u);
                }
                if (EXTRAS && p.gplane && e < nvalid && !(SPECIAL && (flags & F_HALO)))
                    p.gplane[((size_t)frame * p.dst_h + row) * p.dst_w + 4 * q + e] = bgr_to_gray(b, g, r);
            }
            if (CONTENT) {
                uint4* slot = reinterpret_cast<uint4*>(s_prev) + (rloc * QPT + j) * kConsumers + tid;
                uint32_t pv[4] = {cur[0], cur[1], cur[2], cur[3]};
                if (SPECIAL) {
                    if (!(flags & (F_HALO | F_NOPREV))) {
                        if (flags & F_CTXPREV) {
                            for (int e = 0; e < nvalid; ++e) pv[e] = __ldg(p.prev_in + (size_t)row * p.dst_w + 4 * q + e);
                        } else {
                            const uint4 t = *slot;
                            pv[0] = t.x; pv[1] = t.y; pv[2] = t.z; pv[3] = t.w;
                        }
                    }
                    if (flags & F_SAVE)
                        for (int e = 0; e < nvalid; ++e) p.prev_out[(size_t)row * p.dst_w + 4 * q + e] = cur[e];
                    if (EXTRAS && p.vplane && !(flags & F_HALO))
                        for (int e = 0; e < nvalid; ++e)
                            p.vplane[((size_t)frame * p.dst_h + row) * p.dst_w + 4 * q + e] = (uint8_t)(cur[e] >> 16);
                } else {
                    const uint4 t = *slot;
                    pv[0] = t.x; pv[1] = t.y; pv[2] = t.z; pv[3] = t.w;
                    if (EXTRAS && p.vplane)
                        for (int e = 0; e < nvalid; ++e)
                            p.vplane[((size_t)frame * p.dst_h + row) * p.dst_w + 4 * q + e] = (uint8_t)(cur[e] >> 16);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t diff = (e < nvalid) ? __vabsdiffu4(cur[e], pv[e]) : 0u;
                    acc_hv += diff & 0x00ff00ffu;
                    acc_s += (diff >> 8) & 0xffu;
                }
                *slot = make_uint4(cur[0], cur[1], cur[2], cur[3]);
            }
        }
    }
}

// ALIGNED: every source row starts on a 16-byte boundary (base, pitch and frame stride multiples of 16), so the
// per-row misalignment is zero and each thread's smem word offset / funnel shift are loop invariants.
template <bool RESIZE, int PXT, bool CONTENT, bool HIST, bool ALIGNED, bool NV12 = false, bool EXTRAS = true>
__global__ void __launch_bounds__(kThreads) fused_score_kernel(const FusedParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // ---- carve shared memory (host twin: fused_smem_bytes in esd.cu)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);              // [kMaxStages]
    uint64_t* empty_bar = full_bar + kMaxStages;                               // [kMaxStages]
    int4* meta = reinterpret_cast<int4*>(empty_bar + kMaxStages);              // [kMaxStages] {f, rloc|row<<8, flags, nrows}
    uint4* meta_r = reinterpret_cast<uint4*>(meta + kMaxStages);               // [kMaxStages][kMaxRowsPerStage] {b0s,b1s,mis,0}
    int* s_sdiv = reinterpret_cast<int*>(meta_r + kMaxStages * kMaxRowsPerStage);  // [256]
    int* s_hdiv = s_sdiv + 256;                                                // [256]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_hdiv + 256);              // [2][256]
    uint32_t* s_prev = s_hist + 512;                                           // [rows_per_group][PXT*kConsumers]
    size_t off = (size_t)(reinterpret_cast<uint8_t*>(s_prev + (CONTENT ? p.rows_per_group * PXT * kConsumers : 0)) - smem_raw);
    off = (off + 127) & ~(size_t)127;
    uint8_t* s_stage = smem_raw + off;                                         // [stages][rows_per_stage][1|2][rowbuf]

    const int tid = threadIdx.x;
    const int S = p.stages;
    for (int i = tid; i < 256; i += kThreads) {
        s_sdiv[i] = p.sdiv[i];
        s_hdiv[i] = p.hdiv[i];
        s_hist[i] = 0;
        s_hist[256 + i] = 0;
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    const int u_begin = p.cta_unit_begin[blockIdx.x];
    const int u_end = p.cta_unit_begin[blockIdx.x + 1];
    const int R = p.rows_per_group;
    const int RS = p.rows_per_stage;
    // smem bytes of one destination row's source rows: [row0][row1], NV12: [Y row0][Y row1][UV row0][UV row1]
    const int row_slot = (RESIZE ? (NV12 ? 4 : 2) : 1) * p.rowbuf;
    const int stage_bytes = RS * row_slot;

    if (tid >= kConsumers) {
        // ================================================================ producer warp (one lane)
        if (tid != kConsumers) return;
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t full_base = smem_u32(full_bar), empty_base = smem_u32(empty_bar), stage_base = smem_u32(s_stage);
        int s = 0;
        uint32_t par = 1u;  // empty barriers start "free"
        for (int u = u_begin; u < u_end; ++u) {
            const Unit un = p.units[u];
            const int r_begin = un.rg * R;
            const int r_end = min(r_begin + R, p.dst_h);
            const int f_start = (CONTENT && un.f0 > 0) ? un.f0 - 1 : un.f0;
            for (int f = f_start; f < un.f1; ++f) {
                const uint8_t* frame = p.src + (long long)f * p.frame_stride;
                int fflags = 0;
                if (CONTENT) {
                    if (f < un.f0) fflags |= F_HALO;
                    else if (f == 0) fflags |= p.has_prev ? F_CTXPREV : F_NOPREV;
                    if (f == p.n_frames - 1) fflags |= F_SAVE;
                }
                for (int r = r_begin; r < r_end; r += RS) {
                    const int nr = min(RS, r_end - r);
                    mbar_wait(empty_base + 8u * s, par);
                    const uint8_t* src0[kMaxRowsPerStage];
                    const uint8_t* src1[kMaxRowsPerStage];
                    uint32_t nb0[kMaxRowsPerStage], nb1[kMaxRowsPerStage];
                    const uint8_t* srcu0[kMaxRowsPerStage];
                    const uint8_t* srcu1[kMaxRowsPerStage];
                    uint32_t nbu0[kMaxRowsPerStage], nbu1[kMaxRowsPerStage];
                    uint32_t tx = 0;
#pragma unroll
                    for (int q = 0; q < kMaxRowsPerStage; ++q) {
                        if (q < nr) {
                            const YRow yr = p.yrows[r + q];
                            const uint8_t* a0 = frame + (long long)(p.compact ? yr.crow0 : yr.row0) * p.row_stride;
                            const uint32_t mis0 = ALIGNED ? 0u : (uint32_t)(reinterpret_cast<uintptr_t>(a0) & 15u);
                            nb0[q] = (mis0 + (uint32_t)p.row_bytes + 15u) & ~15u;
                            src0[q] = a0 - mis0;
                            uint32_t mis1 = 0;
                            nb1[q] = 0;
                            src1[q] = a0;
                            if (RESIZE) {
                                const uint8_t* a1 = frame + (long long)(p.compact ? yr.crow1 : yr.row1) * p.row_stride;
                                mis1 = ALIGNED ? 0u : (uint32_t)(reinterpret_cast<uintptr_t>(a1) & 15u);
                                nb1[q] = (mis1 + (uint32_t)p.row_bytes + 15u) & ~15u;
                                src1[q] = a1 - mis1;
                            }
                            tx += nb0[q] + nb1[q];
                            uint32_t misuv = 0;
                            if (NV12) {
                                // UV rows: from the UV plane (full layout) or from the compact frame (absolute indices)
                                const uint8_t* uvbase = p.compact ? frame : p.src_uv + (long long)f * p.uv_frame_stride;
                                const long long uvrs = p.compact ? p.row_stride : p.uv_row_stride;
                                uint32_t rows = p.compact ? yr.cuvrows : yr.uvrows;
                                if (!p.compact && p.uv_packed_base >= 0) rows = yr.cuvrows - (uint32_t)p.uv_packed_base * 0x10001u;
                                const uint8_t* u0 = uvbase + (long long)(rows & 0xffffu) * uvrs;
                                const uint8_t* u1 = uvbase + (long long)(rows >> 16) * uvrs;
                                const uint32_t m0 = ALIGNED ? 0u : (uint32_t)(reinterpret_cast<uintptr_t>(u0) & 15u);
                                const uint32_t m1 = ALIGNED ? 0u : (uint32_t)(reinterpret_cast<uintptr_t>(u1) & 15u);
                                nbu0[q] = (m0 + (uint32_t)p.row_bytes + 15u) & ~15u;
                                srcu0[q] = u0 - m0;
                                const bool same = (rows & 0xffffu) == (rows >> 16);  // both source rows share one chroma row
                                nbu1[q] = same ? 0u : ((m1 + (uint32_t)p.row_bytes + 15u) & ~15u);
                                srcu1[q] = u1 - m1;
                                misuv = m0 | ((same ? m0 : m1) << 8) | (same ? 0x10000u : 0u);
                                tx += nbu0[q] + nbu1[q];
                            }
                            meta_r[s * kMaxRowsPerStage + q] = make_uint4(yr.b0s, yr.b1s, mis0 | (mis1 << 8), misuv);
                        }
                    }
                    const int flags = fflags | ((r + nr >= r_end) ? F_FRAME_END : 0);
                    meta[s] = make_int4(f, (r - r_begin) | (r << 8), flags, nr | (un.rg << 8));  // w: rows in the stage | row group << 8
                    const uint32_t bar = full_base + 8u * s;
                    const uint32_t dst = stage_base + (uint32_t)s * stage_bytes;
                    mbar_arrive_expect_tx(bar, tx);
#pragma unroll
                    for (int q = 0; q < kMaxRowsPerStage; ++q) {
                        if (q < nr) {
                            bulk_g2s(dst + q * row_slot, src0[q], nb0[q], bar, pol);
                            if (RESIZE) bulk_g2s(dst + q * row_slot + p.rowbuf, src1[q], nb1[q], bar, pol);
                            if (NV12) {
                                bulk_g2s(dst + q * row_slot + 2 * p.rowbuf, srcu0[q], nbu0[q], bar, pol);
                                if (nbu1[q]) bulk_g2s(dst + q * row_slot + 3 * p.rowbuf, srcu1[q], nbu1[q], bar, pol);
                            }
                        }
                    }
                    if (++s == S) { s = 0; par ^= 1u; }
                }
            }
        }
        // sentinel: tells the consumers to stop
        mbar_wait(empty_base + 8u * s, par);
        meta[s] = make_int4(0, 0, F_END, 0);
        mbar_arrive(full_base + 8u * s);
        return;
    }

    // ==================================================================== consumer warps
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // Lane -> destination column.  Neighbouring columns' taps lie 3 * scale bytes apart in the staged row (22.5 B at 1080p),
    // so 32 consecutive columns hit some shared-memory banks twice (48 % of the shared-load wavefronts were replays,
    // profiles/r01_fused_full.md).  With lanes `lane_stride` columns apart (host-chosen so that the tap words of a warp fall
    // into distinct banks: 8 columns = 180 B = 45 words at 1080p) the loads are conflict-free; the warps interleave.
    const int ks = p.lane_stride;
    const int col = ks * lane + (warp & (ks - 1)) + 32 * ks * (warp / ks);
    uint32_t xoff[PXT], xa01[PXT];
#pragma unroll
    for (int k = 0; k < PXT; ++k) {
        const int d = k * kConsumers + col;
        if (RESIZE) {
            const uint2 xe = (d < p.dst_w) ? p.xtab[d] : make_uint2(0u, 0u);
            xoff[k] = xe.x;
            xa01[k] = xe.y;
        } else {
            xoff[k] = (d < p.dst_w) ? (uint32_t)(3 * d) : 0u;
            xa01[k] = 0;
        }
    }
    const uint32_t full_base = smem_u32(full_bar), empty_base = smem_u32(empty_bar);
    constexpr bool kQuads = !RESIZE && ALIGNED && PXT >= 4;  // wide full-resolution rows: four pixels per thread step
    uint32_t acc_hv = 0, acc_s = 0;  // per-frame, per-thread: H | V << 16 and S
    uint32_t acc_bgr = 0;            // per-frame, per-thread: sum of B+G+R (only when p.want_bgr)
    int hist_buf = 0;
    int s = 0;
    uint32_t par = 0u;
    for (;;) {
        mbar_wait(full_base + 8u * s, par);
        const int4 m = meta[s];
        const int flags = m.z;
        if (flags & F_END) break;
        const int nrows = m.w & 0xff;
        const int rloc0 = m.y & 0xff;
        const int row_first = m.y >> 8;
        const uint8_t* stage = s_stage + (size_t)s * stage_bytes;
        uint32_t* hist_cur = s_hist + hist_buf * 256;
        if (flags & (F_HALO | F_NOPREV | F_CTXPREV | F_SAVE)) {
            for (int q = 0; q < nrows; ++q) {
                if (kQuads) {
                    score_row_quads<(PXT >= 4 ? PXT / 4 : 1), CONTENT, HIST, true, EXTRAS>(p, stage + q * row_slot, flags, rloc0 + q, row_first + q,
                                                                                   m.x, tid, s_sdiv, s_hdiv, s_prev, hist_cur, acc_hv,
                                                                                   acc_s, acc_bgr);
                    continue;
                }
                const uint4 mr = meta_r[s * kMaxRowsPerStage + q];
                const uint32_t mis0 = ALIGNED ? 0u : (mr.z & 0xffu), mis1 = ALIGNED ? 0u : ((mr.z >> 8) & 0xffu);
                const uint8_t* uvp0 = stage + q * row_slot + 2 * p.rowbuf;
                const uint8_t* uvp1 = (mr.w & 0x10000u) ? uvp0 : uvp0 + p.rowbuf;
                score_row<RESIZE, PXT, CONTENT, HIST, true, NV12, EXTRAS>(p, stage + q * row_slot, stage + q * row_slot + p.rowbuf, uvp0, uvp1,
                                                                  ALIGNED ? (mr.w & 0x10000u) : mr.w, mis0, mis1, mr.x, mr.y, flags, rloc0 + q,
                                                                  row_first + q, m.x, tid, col, xoff, xa01, s_sdiv, s_hdiv, s_prev,
                                                                  hist_cur, acc_hv, acc_s, acc_bgr);
            }
        } else {
            // common path (no rare flag): rows in pairs, so that the two independent dependency chains overlap
            auto fast_row = [&](int q) {
                if (kQuads) {
                    score_row_quads<(PXT >= 4 ? PXT / 4 : 1), CONTENT, HIST, false, EXTRAS>(p, stage + q * row_slot, flags, rloc0 + q, row_first + q,
                                                                                    m.x, tid, s_sdiv, s_hdiv, s_prev, hist_cur, acc_hv,
                                                                                    acc_s, acc_bgr);
                    return;
                }
                const uint4 mr = meta_r[s * kMaxRowsPerStage + q];
                const uint32_t mis0 = ALIGNED ? 0u : (mr.z & 0xffu), mis1 = ALIGNED ? 0u : ((mr.z >> 8) & 0xffu);
                const uint8_t* uvp0 = stage + q * row_slot + 2 * p.rowbuf;
                const uint8_t* uvp1 = (mr.w & 0x10000u) ? uvp0 : uvp0 + p.rowbuf;
                score_row<RESIZE, PXT, CONTENT, HIST, false, NV12, EXTRAS>(p, stage + q * row_slot, stage + q * row_slot + p.rowbuf, uvp0, uvp1,
                                                                   ALIGNED ? (mr.w & 0x10000u) : mr.w, mis0, mis1, mr.x, mr.y, flags, rloc0 + q,
                                                                   row_first + q, m.x, tid, col, xoff, xa01, s_sdiv, s_hdiv, s_prev,
                                                                   hist_cur, acc_hv, acc_s, acc_bgr);
            };
            int q = 0;
            if (!kQuads && !NV12 && PXT == 1) {
                for (; q + 1 < nrows; q += 2) {
                    fast_row(q);
                    fast_row(q + 1);
                }
            }
            for (; q < nrows; ++q) fast_row(q);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_base + 8u * s);  // stage may be refilled
        if (++s == S) { s = 0; par ^= 1u; }

        if (flags & F_FRAME_END) {
            const bool scored = !(flags & F_HALO);
            const int group = m.w >> 8;
            if (CONTENT) {
                if (scored) {
                    const uint32_t sh = __reduce_add_sync(0xffffffffu, acc_hv & 0xffffu);
                    const uint32_t sv = __reduce_add_sync(0xffffffffu, acc_hv >> 16);
                    const uint32_t ss = __reduce_add_sync(0xffffffffu, acc_s);
                    const uint32_t sb = (EXTRAS && p.want_bgr) ? __reduce_add_sync(0xffffffffu, acc_bgr) : 0u;
                    if (lane == 0)
                        p.part[((size_t)m.x * p.n_groups + group) * kConsumerWarps + warp] = make_uint4(sh, ss, sv, sb);
                }
                acc_hv = 0;
                acc_s = 0;
                acc_bgr = 0;
            }
            if (HIST && scored) {
                consumer_bar_sync();  // all smem atomics of this frame have landed
                if (tid < p.bins) {
                    uint32_t* hslot = &s_hist[hist_buf * 256 + tid];
                    p.hist_part[((size_t)m.x * p.n_groups + group) * p.bins + tid] = (uint16_t)*hslot;
                    *hslot = 0;
                }
                hist_buf ^= 1;  // the other buffer was zeroed one frame ago (before the barrier above)
            }
        }
    }
}

// ----------------------------------------------------------------------------------- finalize
struct ScoreWeights {
    double w[4];
    double div;
};

// IEEE double, left to right, no FMA: PySceneDetect ContentDetector._calculate_frame_score (A.4)
__device__ __forceinline__ double content_val_of(const unsigned long long s[3], unsigned long long edge_sum, double npx,
                                                 const ScoreWeights& W) {
    const double dh = __ddiv_rn((double)s[0], npx);
    const double ds = __ddiv_rn((double)s[1], npx);
    const double dv = __ddiv_rn((double)s[2], npx);
    // delta_edges is 0.0 unless the weight is > 0 (PySceneDetect only computes edges then)
    const double de = W.w[3] > 0.0 ? __ddiv_rn((double)edge_sum, npx) : 0.0;
    double acc = 0.0;
    acc = __dadd_rn(acc, __dmul_rn(dh, W.w[0]));
    acc = __dadd_rn(acc, __dmul_rn(ds, W.w[1]));
    acc = __dadd_rn(acc, __dmul_rn(dv, W.w[2]));
    acc = __dadd_rn(acc, __dmul_rn(de, W.w[3]));
    return __ddiv_rn(acc, W.div);
}

// Planar I420 (what software decoders emit: Y plane, U plane, V plane, chroma planes half-size) -> the interleaved UV rows the
// NV12 variant of the fused kernel stages.  One block per (touched chroma row j, frame): the U and V halves of the row go
// through shared memory, so the kernel also works IN PLACE on a ring-slot row that holds [U row][V row] (what the ingest's DMA
// leaves there: a copy engine cannot interleave).  Only the chroma rows the taps touch are converted.
constexpr int kInterleaveThreads = 128;
__global__ void __launch_bounds__(kInterleaveThreads) i420_interleave_kernel(
    const uint8_t* __restrict__ u_plane, const uint8_t* __restrict__ v_plane, long long src_frame_stride, long long src_pitch,
    const int* __restrict__ src_rows /* [gridDim.x] chroma row per j, or nullptr: row j */, uint8_t* __restrict__ dst,
    long long dst_frame_stride, long long dst_pitch, int half_w) {
    extern __shared__ uint8_t s_uv[];  // [2][half_w]
    const int j = blockIdx.x, f = blockIdx.y;
    const long long r = src_rows ? src_rows[j] : j;
    const uint8_t* u = u_plane + f * src_frame_stride + r * src_pitch;
    const uint8_t* v = v_plane + f * src_frame_stride + r * src_pitch;
    for (int x = threadIdx.x; x < half_w; x += kInterleaveThreads) {
        s_uv[x] = u[x];
        s_uv[half_w + x] = v[x];
    }
    __syncthreads();
    uint8_t* out = dst + f * dst_frame_stride + (long long)j * dst_pitch;
    if ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
        for (int x = 2 * threadIdx.x; x < half_w; x += 2 * kInterleaveThreads) {  // two chroma pairs per 32-bit store
            if (x + 1 < half_w) {
                *reinterpret_cast<uint32_t*>(out + 2 * x) = (uint32_t)s_uv[x] | ((uint32_t)s_uv[half_w + x] << 8) |
                                                            ((uint32_t)s_uv[x + 1] << 16) | ((uint32_t)s_uv[half_w + x + 1] << 24);
            } else {
                out[2 * x] = s_uv[x];
                out[2 * x + 1] = s_uv[half_w + x];
            }
        }
    } else {
        for (int x = threadIdx.x; x < half_w; x += kInterleaveThreads) {
            out[2 * x] = s_uv[x];
            out[2 * x + 1] = s_uv[half_w + x];
        }
    }
}

// one warp per frame: reduce the per-(group,warp) partials, emit sums and both content_val flavours
__global__ void finalize_sums_kernel(const uint4* __restrict__ part, int n_frames, int parts_per_frame, double npx,
                                     ScoreWeights wc, ScoreWeights wa, unsigned long long* __restrict__ sums3,
                                     double* __restrict__ content_val, double* __restrict__ adaptive_val,
                                     double* __restrict__ average_rgb, const uint32_t* __restrict__ edge_counts) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    unsigned long long s[3] = {0, 0, 0};
    unsigned long long sb = 0;
    const uint4* pf = part + (size_t)f * parts_per_frame;
    for (int i = lane; i < parts_per_frame; i += 32) {
        const uint4 v = pf[i];
        s[0] += v.x; s[1] += v.y; s[2] += v.z; sb += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
        s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
        s[2] += __shfl_xor_sync(0xffffffffu, s[2], o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) {
        sums3[3 * (size_t)f] = s[0];
        sums3[3 * (size_t)f + 1] = s[1];
        sums3[3 * (size_t)f + 2] = s[2];
        // sum |edges - last_edges| over 0/255 maps = 255 x (number of differing pixels)
        const unsigned long long es = edge_counts ? 255ull * edge_counts[f] : 0ull;
        content_val[f] = content_val_of(s, es, npx, wc);
        adaptive_val[f] = content_val_of(s, es, npx, wa);
        // ThresholdDetector._compute_frame_average: numpy.sum(frame) / float(rows * cols * channels)
        average_rgb[f] = __ddiv_rn((double)sb, __dmul_rn(npx, 3.0));
    }
}

// one block per frame: sum the per-group partial histograms into uint32 counts
__global__ void finalize_hist_counts_kernel(const uint16_t* __restrict__ hist_part, int n_groups, int bins,
                                            uint32_t* __restrict__ counts) {
    const int f = blockIdx.x;
    for (int b = threadIdx.x; b < bins; b += blockDim.x) {
        uint32_t c = 0;
        const uint16_t* hp = hist_part + (size_t)f * n_groups * bins + b;
        for (int g = 0; g < n_groups; ++g) c += hp[(size_t)g * bins];
        counts[(size_t)f * bins + b] = c;
    }
}

// cv2.normalize(hist, hist) (L2) then cv2.compareHist(prev, cur, CORREL) with OpenCV's summation order
// (two interleaved double lanes over blocks of 4, scalar tail), SURVEY.md A.7.  One block per frame.
// counts points at the batch's first frame inside the ctx-wide array; counts[-bins..-1] is the frame before.
__global__ void hist_diff_kernel(const uint32_t* __restrict__ counts, int bins, int first_has_prev,
                                 double* __restrict__ hist_diff) {
    __shared__ float hn[2][256];
    __shared__ double red[2][32];
    __shared__ double lanes_out[10];
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    if (f == 0 && !first_has_prev) {
        if (tid == 0) hist_diff[0] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    // L2 norms (exact integer arithmetic in double) of frame f-1 and f
    for (int which = 0; which < 2; ++which) {
        const uint32_t* c = counts + ((long long)f - 1 + which) * bins;
        double ss = 0.0;
        for (int b = tid; b < bins; b += blockDim.x) {
            const double v = (double)c[b];
            ss += v * v;  // exact: integers below 2^53
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((tid & 31) == 0) red[which][tid >> 5] = ss;
    }
    __syncthreads();
    for (int which = 0; which < 2; ++which) {
        double ss = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ss += red[which][w];
        const double nrm = __dsqrt_rn(ss);
        const double scale = nrm > 2.220446049250313e-16 ? __ddiv_rn(1.0, nrm) : 0.0;
        const float fs = __double2float_rn(scale);
        const uint32_t* c = counts + ((long long)f - 1 + which) * bins;
        for (int b = tid; b < bins; b += blockDim.x) hn[which][b] = __fmul_rn((float)c[b], fs);
    }
    __syncthreads();
    // 5 quantities x 2 lanes: strictly sequential double chains in OpenCV's order
    if (tid < 10) {
        const int q = tid >> 1, ln = tid & 1;
        const int nv = (bins / 4) * 4;
        double acc = 0.0;
        for (int j = ln; j < nv; j += 2) {
            const double a = (double)hn[0][j], b = (double)hn[1][j];
            double t;
            switch (q) {
                case 0: t = __dmul_rn(a, b); break;  // s12
                case 1: t = a; break;                // s1
                case 2: t = __dmul_rn(a, a); break;  // s11
                case 3: t = b; break;                // s2
                default: t = __dmul_rn(b, b); break; // s22
            }
            acc = __dadd_rn(acc, t);
        }
        lanes_out[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double q[5];
        const int nv = (bins / 4) * 4;
        for (int k = 0; k < 5; ++k) q[k] = __dadd_rn(0.0, __dadd_rn(lanes_out[2 * k], lanes_out[2 * k + 1]));
        for (int j = nv; j < bins; ++j) {
            const double a = (double)hn[0][j], b = (double)hn[1][j];
            q[0] = __dadd_rn(q[0], __dmul_rn(a, b));
            q[1] = __dadd_rn(q[1], a);
            q[2] = __dadd_rn(q[2], __dmul_rn(a, a));
            q[3] = __dadd_rn(q[3], b);
            q[4] = __dadd_rn(q[4], __dmul_rn(b, b));
        }
        const double s12 = q[0], s1 = q[1], s11 = q[2], s2 = q[3], s22 = q[4];
        const double scale = __ddiv_rn(1.0, (double)bins);
        const double num = __dadd_rn(s12, -__dmul_rn(__dmul_rn(s1, s2), scale));
        const double d1 = __dadd_rn(s11, -__dmul_rn(__dmul_rn(s1, s1), scale));
        const double d2 = __dadd_rn(s22, -__dmul_rn(__dmul_rn(s2, s2), scale));
        const double den2 = __dmul_rn(d1, d2);
        hist_diff[f] = fabs(den2) > 2.220446049250313e-16 ? __ddiv_rn(num, __dsqrt_rn(den2)) : 1.0;
    }
}

// AdaptiveDetector rolling-window ratio (A.6) of target frame t: val[t - w .. t + w] must exist.
__device__ __forceinline__ double adaptive_ratio_at(const double* __restrict__ val, long long t, int w, double min_content_val) {
    double sum = 0.0;
    bool first = true;
    for (int k = -w; k <= w; ++k) {
        if (k == 0) continue;
        const double s = val[t + k];
        sum = first ? s : __dadd_rn(sum, s);  // int 0 + s == s
        first = false;
    }
    const double avg = __ddiv_rn(sum, __dmul_rn(2.0, (double)w));
    const double target = val[t];
    double r = 0.0;
    if (!(fabs(avg) < 0.00001)) {
        const double q = __ddiv_rn(target, avg);
        r = (255.0 < q) ? 255.0 : q;  // Python min(q, 255.0)
    } else if (target >= min_content_val) {
        r = 255.0;
    }
    return r;
}

// one thread per target frame index t in [t_begin, t_end); val/ratio are indexed from the first frame of the video
__global__ void adaptive_ratio_kernel(const double* __restrict__ val, double* __restrict__ ratio, long long t_begin,
                                      long long t_end, int w, double min_content_val) {
    const long long t = t_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    ratio[t] = adaptive_ratio_at(val, t, w, min_content_val);
}

// Stand-alone flavour for a whole score array (global decision pass of frame-range sharding): thread per frame index,
// NaN where the window is incomplete, so no separate fill pass is needed.
__global__ void adaptive_ratio_full_kernel(const double* __restrict__ val, double* __restrict__ ratio, long long n, int w,
                                           double min_content_val) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    ratio[t] = (t < w || t >= n - w) ? __longlong_as_double(0x7ff8000000000000LL) : adaptive_ratio_at(val, t, w, min_content_val);
}

// ----------------------------------------------------------------------------------- edges
// ContentDetector._detect_edges on the device: numpy.median(lum) -> Canny(low, high) -> dilate(k x k ones).
// One CTA per frame at detector resolution; the whole V plane and the edge map live in shared memory.
// cv2.Canny (aperture 3, L1 gradient) restated from OpenCV canny.cpp: Sobel with replicated borders, |dx|+|dy|,
// non-maximum suppression with the TG22 fixed-point sector test against a zero-bordered magnitude map, hysteresis =
// weak pixels 8-connected to a strong one.  Output: one bit per pixel, [n][words_per_frame].
constexpr int kEdgeThreads = 1024;

__device__ __forceinline__ int sobel_mag(const uint8_t* __restrict__ V, int w, int h, int x, int y, int& dx, int& dy) {
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1);
    const uint8_t* r0 = V + max(y - 1, 0) * w;
    const uint8_t* r1 = V + y * w;
    const uint8_t* r2 = V + min(y + 1, h - 1) * w;
    const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], e = r1[xp], f = r2[xm], g = r2[x], i = r2[xp];
    dx = (c + 2 * e + i) - (a + 2 * d + f);
    dy = (f + 2 * g + i) - (a + 2 * b + c);
    return abs(dx) + abs(dy);
}
__device__ __forceinline__ int mag_at(const uint8_t* __restrict__ V, int w, int h, int x, int y) {
    if (x < 0 || y < 0 || x >= w || y >= h) return 0;  // the magnitude map has a zero border
    int dx, dy;
    return sobel_mag(V, w, h, x, y, dx, dy);
}

// Walks the pixels i = tid, tid + T, tid + 2T, ... of a w-wide image keeping (x, y) without a division per pixel.
struct PixelWalk {
    int x, y, dx, dy, w;
    __device__ __forceinline__ PixelWalk(int first, int step, int width) : w(width) {
        y = first / width; x = first - y * width;
        dy = step / width; dx = step - dy * width;
    }
    __device__ __forceinline__ void next() {
        x += dx; y += dy;
        if (x >= w) { x -= w; ++y; }
    }
};

// Output: one bit per pixel in row-padded order, [n][h][ceil(w / 32)] words (padding bits are zero).
__global__ void __launch_bounds__(kEdgeThreads) edges_kernel(const uint8_t* __restrict__ vplane, int w, int h, int ksize,
                                                             uint32_t* __restrict__ bitmaps, int words_per_frame) {
    extern __shared__ __align__(16) uint8_t esm[];
    const int npx = w * h;
    const int npx_pad = (npx + 31) & ~31;
    uint8_t* V = esm;                  // [npx_pad]
    uint8_t* map = esm + npx_pad;      // [npx_pad]   0 none, 1 weak, 2/4 frontier, 3 edge
    __shared__ uint32_t hist[256];
    __shared__ int s_low, s_high;
    __shared__ int changed[3];
    const int tid = threadIdx.x;
    const uint8_t* src = vplane + (size_t)blockIdx.x * npx;

    for (int i = tid; i < 256; i += kEdgeThreads) hist[i] = 0;
    if (tid < 3) changed[tid] = 0;
    __syncthreads();
    for (int i = tid; i < npx_pad; i += kEdgeThreads) {
        const uint8_t v = i < npx ? src[i] : 0;
        V[i] = v;
        if (i < npx) atomicAdd(&hist[v], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        // numpy.median: middle element, or the mean of the two middle elements for an even count
        const int k_lo = (npx - 1) / 2, k_hi = npx / 2;
        int acc = 0, v_lo = -1, v_hi = -1;
        for (int b = 0; b < 256 && v_hi < 0; ++b) {
            acc += (int)hist[b];
            if (v_lo < 0 && acc > k_lo) v_lo = b;
            if (acc > k_hi) v_hi = b;
        }
        const double median = __ddiv_rn((double)(v_lo + v_hi), 2.0);
        const double sigma = 1.0 / 3.0;
        const double lo = __dmul_rn(1.0 - sigma, median), hi = __dmul_rn(1.0 + sigma, median);
        int low = (int)(lo > 0.0 ? lo : 0.0);
        int high = (int)(hi < 255.0 ? hi : 255.0);
        if (low > high) { const int t = low; low = high; high = t; }
        s_low = low;
        s_high = high;
    }
    __syncthreads();
    const int low = s_low, high = s_high;
    constexpr int TG22 = 13573;  // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
    {
        PixelWalk pw(tid, kEdgeThreads, w);
        for (int i = tid; i < npx_pad; i += kEdgeThreads, pw.next()) {
            uint8_t out = 0;
            if (i < npx) {
                const int x = pw.x, y = pw.y;
                int dx, dy;
                const int m = sobel_mag(V, w, h, x, y, dx, dy);
                if (m > low) {
                    const int xa = abs(dx), ya = abs(dy) << 15;
                    const int tg22x = xa * TG22;
                    bool ok;
                    if (ya < tg22x) ok = m > mag_at(V, w, h, x - 1, y) && m >= mag_at(V, w, h, x + 1, y);
                    else {
                        const int tg67x = tg22x + (xa << 16);
                        if (ya > tg67x) ok = m > mag_at(V, w, h, x, y - 1) && m >= mag_at(V, w, h, x, y + 1);
                        else {
                            const int sgn = ((dx ^ dy) < 0) ? -1 : 1;
                            ok = m > mag_at(V, w, h, x - sgn, y - 1) && m > mag_at(V, w, h, x + sgn, y + 1);
                        }
                    }
                    if (ok) out = (m > high) ? 2 : 1;
                }
            }
            map[i] = out;
        }
    }
    __syncthreads();
    // hysteresis: breadth-first growth of the strong set through weak pixels; frontier label alternates 2 <-> 4
    const uint32_t* map32 = reinterpret_cast<const uint32_t*>(map);
    for (int r = 0;; ++r) {
        const uint32_t cur = (r & 1) ? 4u : 2u, nxt = (r & 1) ? 2u : 4u;
        if (tid == 0) changed[(r + 1) % 3] = 0;
        bool any = false;
        for (int wi = tid; wi < npx_pad / 4; wi += kEdgeThreads) {
            const uint32_t word = map32[wi];
            const uint32_t t = word ^ (cur * 0x01010101u);
            if (!((t - 0x01010101u) & ~t & 0x80808080u)) continue;  // no byte equals `cur`
            for (int b = 0; b < 4; ++b) {
                if (((word >> (8 * b)) & 0xffu) != cur) continue;
                const int i = wi * 4 + b;
                const int y = i / w, x = i - y * w;
                for (int yy = max(y - 1, 0); yy <= min(y + 1, h - 1); ++yy)
                    for (int xx = max(x - 1, 0); xx <= min(x + 1, w - 1); ++xx)
                        if (map[yy * w + xx] == 1) { map[yy * w + xx] = (uint8_t)nxt; any = true; }
                map[i] = 3;
            }
        }
        if (any) changed[r % 3] = 1;
        __syncthreads();
        if (!changed[r % 3]) break;
    }
    // Every edge pixel now carries the label 3.  Dilate with a k x k block of ones (anchor at the centre, outside
    // ignored) on a bit-packed copy: rows of ceil(w / 32) words, horizontal then vertical OR of shifted words.
    const int wpr = (w + 31) >> 5;  // words per row
    uint32_t* bits0 = reinterpret_cast<uint32_t*>(esm + 2 * (size_t)npx_pad);  // [h][wpr] packed edge map
    uint32_t* bits1 = bits0 + h * wpr;                          // [h][wpr] after the horizontal pass
    const int rad = ksize / 2, rad1 = ksize - 1 - rad;          // window [p - rad, p + rad1] (anchor = k / 2)
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int wd = warp; wd < h * wpr; wd += kEdgeThreads / 32) {
            const int y = wd / wpr, j = wd - y * wpr;
            const int x = 32 * j + lane;
            const bool e = x < w && map[y * w + x] == 3;
            const uint32_t bits = __ballot_sync(0xffffffffu, e);
            if (lane == 0) bits0[wd] = bits;
        }
    }
    __syncthreads();
    for (int wd = tid; wd < h * wpr; wd += kEdgeThreads) {
        const int y = wd / wpr, j = wd - y * wpr;
        const uint32_t c = bits0[wd];
        const uint32_t l = j > 0 ? bits0[wd - 1] : 0u, r = j + 1 < wpr ? bits0[wd + 1] : 0u;
        uint32_t o = c;
        // output bit x is set when any input bit in [x - rad, x + rad1] is set
        for (int d = 1; d <= rad; ++d) o |= (c << d) | (l >> (32 - d));    // input at x - d
        for (int d = 1; d <= rad1; ++d) o |= (c >> d) | (r << (32 - d));   // input at x + d
        if (j == wpr - 1 && (w & 31)) o &= (1u << (w & 31)) - 1u;           // keep the padding bits zero
        bits1[wd] = o;
    }
    __syncthreads();
    for (int wd = tid; wd < h * wpr; wd += kEdgeThreads) {
        const int y = wd / wpr;
        uint32_t o = 0;
        for (int yy = max(y - rad, 0); yy <= min(y + rad1, h - 1); ++yy) o |= bits1[wd + (yy - y) * wpr];
        bitmaps[(size_t)blockIdx.x * words_per_frame + wd] = o;
    }
}

// number of pixels whose dilated edge bit differs from the previous frame's (block per frame);
// bitmaps points at the batch's first frame, prev_last at the last frame of the previous batch (or nullptr)
__global__ void edge_delta_kernel(const uint32_t* __restrict__ bitmaps, const uint32_t* __restrict__ prev_last,
                                  int words_per_frame, int first_has_prev, uint32_t* __restrict__ counts) {
    __shared__ uint32_t red[32];
    const int f = blockIdx.x;
    const uint32_t* cur = bitmaps + (size_t)f * words_per_frame;
    const uint32_t* prv = f > 0 ? cur - words_per_frame : (first_has_prev ? prev_last : nullptr);
    uint32_t c = 0;
    if (prv)
        for (int i = threadIdx.x; i < words_per_frame; i += blockDim.x) c += __popc(cur[i] ^ prv[i]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        counts[f] = t;
    }
}

// ----------------------------------------------------------------------------------- perceptual hash
// HashDetector.hash_frame on the device, one CTA per frame, reading the BGR2GRAY plane the fused kernel wrote:
//   cv2.resize(gray, (S, S), INTER_AREA)  -- OpenCV resize.cpp restated exactly: ResizeAreaFast_ (integral scales:
//       integer box sum, saturate_cast(sum * float(1/area)); 2x2 is (sum + 2) >> 2) or ResizeArea_ (float32 row
//       buffer `buf += S * alpha` in table order, rows combined as `sum = beta * buf` / `sum += beta * buf`, no FMA)
//   x = float32(v) / max(v)   (max == 0 -> 1)
//   DCT-II of x, low-frequency hs x hs block, evaluated in float64 and narrowed to float32.  cv2.dct's own float32
//       rounding is build-dependent (its IPP and plain paths differ by 1-2 ulp), so this stage is tolerance-parity.
//   bits = coefficient > numpy.median(block)   (float32; even count: (a + b) / 2 in float32)
struct HashParams {
    int w, h;          // detector-resolution frame
    int S, hs;         // DCT size (hash_size * lowpass) and hash size
    int fast;          // both scales integral (ResizeAreaFast_)
    int isx, isy;      // integral scales (fast path)
    float fast_scale;  // float32(1 / (isx * isy))
    int xsum_int;      // horizontal pass as an exact integer byte sum: fast path, or every x weight is the same power of two
    int xwords;        // > 0: every column sums 4 * xwords bytes from a 4-byte boundary (hash_kernel<WORDS>)
    float xalpha;      // that weight
    int n_xent, n_yent, n_bands;
    // computeResizeAreaTab tables, grouped per destination index, in one block each (staged into shared memory):
    const int* itab;   // xb[S + 1] | xs0[S] (first source column) | yb[S + 1] | ysrc[n_yent] | bands[n_bands] {dy0, dy1, r0, nrows}
    const float* wtab; // xw[n_xent] | yw[n_yent]
    const double* C;   // [hs][S] orthonormal DCT-II rows
    int words;         // ceil(hs * hs / 32)
    int band_rows;     // source rows of horizontal sums the shared-memory band buffer holds
    float* margin_out; // [n] smallest |coefficient - median| of each frame of this launch (how far its weakest bit is from flipping)
};
constexpr int kHashThreads = 256;

// order-preserving float -> uint key (for the radix select of the median)
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of keys[0..n) by 4 passes of an 8-bit radix histogram.  All threads of the CTA call it;
// hist is 256 words of shared memory, sel two more.  Returns the key to every thread.
__device__ uint32_t block_radix_select(const uint32_t* __restrict__ keys, int n, int k, uint32_t* hist, uint32_t* sel) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;  // kHashThreads == 256 bins
        __syncthreads();
        for (int i = tid; i < n; i += kHashThreads) {
            const uint32_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t b[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { b[q] = hist[8 * tid + q]; s += b[q]; }
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            uint32_t cum = incl - s;
            if ((uint32_t)k >= cum && (uint32_t)k < incl) {  // exactly one lane
                int q = 0;
                while (cum + b[q] <= (uint32_t)k) { cum += b[q]; ++q; }
                sel[0] = (uint32_t)(8 * tid + q);
                sel[1] = (uint32_t)k - cum;
            }
        }
        __syncthreads();
        prefix |= sel[0] << shift;
        mask |= 0xffu << shift;
        k = (int)sel[1];
    }
    return prefix;
}

// Persistent over frames (grid-stride), tables staged once per CTA.  Shared memory:
//   C[hs*S] f64 | xd[S*S] f32 | union { hrow[band_rows][S] f32 (area resize) ; T[hs*S] f64 + keys[hs*hs] u32 } | wtab f32 | itab i32
// WORDS > 0: every destination column sums exactly 4 * WORDS bytes starting on a 4-byte boundary (e.g. 256 -> 32 columns:
// 8 bytes), so the horizontal pass is WORDS aligned word loads and DP4As; WORDS == 0 is the general code.
template <int WORDS>
__global__ void __launch_bounds__(kHashThreads, 8) hash_kernel(const uint8_t* __restrict__ gplane, int n_frames, HashParams P,
                                                            uint8_t* __restrict__ small_out, uint32_t* __restrict__ bits_out) {
    extern __shared__ __align__(16) uint8_t hsm[];
    const int S = P.S, hs = P.hs, N = hs * hs;
    double* Cs = reinterpret_cast<double*>(hsm);                // [hs * S]
    float* xd = reinterpret_cast<float*>(Cs + hs * S);          // [S * S] normalised thumbnail (S * S is even)
    uint8_t* region = reinterpret_cast<uint8_t*>(xd + S * S);
    const size_t region_bytes = max((size_t)P.band_rows * S * sizeof(float), (size_t)hs * S * sizeof(double) + (size_t)N * sizeof(uint32_t));
    float* hrow = reinterpret_cast<float*>(region);             // [band_rows][S] horizontal sums
    double* T = reinterpret_cast<double*>(region);              // [hs * S]
    uint32_t* keys = reinterpret_cast<uint32_t*>(T + hs * S);   // [N] order-preserving coefficient keys
    float* s_xw = reinterpret_cast<float*>(region + ((region_bytes + 7) & ~(size_t)7));  // [n_xent]
    float* s_yw = s_xw + P.n_xent;                              // [n_yent]
    int* s_xb = reinterpret_cast<int*>(s_yw + P.n_yent);        // [S + 1]
    int* s_xs0 = s_xb + S + 1;                                  // [S]
    int* s_yb = s_xs0 + S;                                      // [S + 1]
    int* s_ysrc = s_yb + S + 1;                                 // [n_yent]
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_sel[2];
    __shared__ int s_max;
    __shared__ uint32_t s_cnt, s_min, s_margin;
    const int tid = threadIdx.x;
    const int n_itab = 3 * S + 2 + P.n_yent;
    for (int i = tid; i < n_itab; i += kHashThreads) s_xb[i] = P.itab[i];
    for (int i = tid; i < P.n_xent + P.n_yent; i += kHashThreads) s_xw[i] = P.wtab[i];
    for (int i = tid; i < hs * S; i += kHashThreads) Cs[i] = P.C[i];
    const int4* bands = reinterpret_cast<const int4*>(P.itab + ((n_itab + 3) & ~3));  // a few uniform loads per frame

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const uint8_t* g = gplane + (size_t)f * P.w * P.h;
        __syncthreads();  // tables staged / previous frame done with every buffer
        if (tid == 0) { s_max = 0; s_cnt = 0; s_min = 0xffffffffu; s_margin = 0x7f800000u; }
        // ---- INTER_AREA, separable exactly like OpenCV: per source row the horizontal sums (phase 1), then the
        //      vertical combination per destination pixel (phase 2), in bands of destination rows whose source rows fit
        //      the shared-memory buffer.
        int vmax = 0;
        for (int bi = 0; bi < P.n_bands; ++bi) {
            const int4 band = bands[bi];  // {dy0, dy1, r0, nrows}
            const int r0 = band.z;
            if (bi > 0) __syncthreads();  // previous band's phase 2 is done with hrow
            {
                PixelWalk pw(tid, kHashThreads, S);
                for (int it = tid; it < band.w * S; it += kHashThreads, pw.next()) {
                    const int dx = pw.x;
                    const int xb = s_xb[dx], cnt = s_xb[dx + 1] - xb;
                    const uint8_t* px = g + (size_t)(r0 + pw.y) * P.w + s_xs0[dx];  // the taps of a column are consecutive pixels
                    if (WORDS > 0) {
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(px);
                        uint32_t sum = 0;
#pragma unroll
                        for (int j = 0; j < WORDS; ++j) sum = __dp4a(wp[j], 0x01010101u, sum);
                        hrow[it] = P.fast ? __uint_as_float(sum) : __fmul_rn((float)sum, P.xalpha);
                    } else if (P.xsum_int) {
                        // exact integer sum of cnt bytes at any alignment: masked words through DP4A
                        const uintptr_t a = reinterpret_cast<uintptr_t>(px);
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                        const int lead = (int)(a & 3), total = lead + cnt, nw = (total + 3) >> 2;
                        uint32_t sum = 0;
                        for (int j = 0; j < nw; ++j) {
                            uint32_t word = wp[j];
                            if (j == 0) word &= 0xffffffffu << (8 * lead);
                            if (j == nw - 1 && (total & 3)) word &= 0xffffffffu >> (8 * (4 - (total & 3)));
                            sum = __dp4a(word, 0x01010101u, sum);
                        }
                        // fast path keeps the integer; otherwise sum * 2^-k equals the float accumulation exactly
                        hrow[it] = P.fast ? __uint_as_float(sum) : __fmul_rn((float)sum, P.xalpha);
                    } else {
                        float buf = 0.f;
                        for (int k = 0; k < cnt; ++k) buf = __fadd_rn(buf, __fmul_rn((float)px[k], s_xw[xb + k]));
                        hrow[it] = buf;
                    }
                }
            }
            __syncthreads();
            {
                const int first = band.x * S + tid;
                PixelWalk pw(first, kHashThreads, S);
                for (int p = first; p < band.y * S; p += kHashThreads, pw.next()) {
                    const int dx = pw.x, dy = pw.y;
                    const int yb = s_yb[dy], ye = s_yb[dy + 1];
                    int v;
                    if (P.fast) {
                        int sum = 0;
                        for (int ky = yb; ky < ye; ++ky) sum += __float_as_int(hrow[(s_ysrc[ky] - r0) * S + dx]);
                        if (P.isx == 2 && P.isy == 2) v = (sum + 2) >> 2;
                        else v = min(255, max(0, __float2int_rn(__fmul_rn((float)sum, P.fast_scale))));
                    } else {
                        float sum = 0.f;
                        for (int ky = yb; ky < ye; ++ky) {
                            const float t = __fmul_rn(s_yw[ky], hrow[(s_ysrc[ky] - r0) * S + dx]);
                            sum = (ky == yb) ? t : __fadd_rn(sum, t);
                        }
                        v = min(255, max(0, __float2int_rn(sum)));
                    }
                    xd[p] = (float)v;
                    vmax = max(vmax, v);
                    if (small_out) small_out[(size_t)f * S * S + p] = (uint8_t)v;
                }
            }
        }
        vmax = __reduce_max_sync(0xffffffffu, vmax);
        if ((tid & 31) == 0) atomicMax(&s_max, vmax);
        __syncthreads();
        const float fmax_v = (float)(s_max == 0 ? 1 : s_max);
        for (int p = tid; p < S * S; p += kHashThreads) xd[p] = __fdiv_rn(xd[p], fmax_v);
        __syncthreads();
        // ---- T = C[:hs] . x   then   D = T . C[:hs]^T   (float64, narrowed to float32 at the end)
        {
            PixelWalk pw(tid, kHashThreads, S);
            for (int idx = tid; idx < hs * S; idx += kHashThreads, pw.next()) {
                const double* cu = Cs + pw.y * S;
                const float* xj = xd + pw.x;
                double acc = 0.0;
#pragma unroll 8
                for (int i = 0; i < S; ++i) acc = fma(cu[i], (double)xj[i * S], acc);
                T[idx] = acc;
            }
        }
        __syncthreads();
        float mine[4];  // this thread's coefficients (N <= 1024)
        {
            PixelWalk pw(tid, kHashThreads, hs);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = q * kHashThreads + tid;
                mine[q] = 0.f;
                if (idx < N) {
                    const double* tu = T + pw.y * S;
                    const double* cv = Cs + pw.x * S;
                    double acc = 0.0;
#pragma unroll 8
                    for (int j = 0; j < S; ++j) acc = fma(tu[j], cv[j], acc);
                    mine[q] = (float)acc;
                }
                pw.next();
            }
        }
        __syncthreads();  // every T read is done before keys (behind T) and the next frame's hrow are written
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = q * kHashThreads + tid;
            if (idx < N) keys[idx] = float_key(mine[q]);
        }
        __syncthreads();
        // ---- numpy.median: the middle element, or the float32 mean of the two middle elements
        const int k_lo = (N - 1) / 2, k_hi = N / 2;
        const uint32_t key_lo = block_radix_select(keys, N, k_lo, s_hist, s_sel);
        float med = key_float(key_lo);
        if (k_hi != k_lo) {
            // the next order statistic: key_lo again if it occurs often enough, else the smallest key above it
            uint32_t cnt = 0, mn = 0xffffffffu;
            for (int i = tid; i < N; i += kHashThreads) {
                const uint32_t key = keys[i];
                cnt += key <= key_lo ? 1u : 0u;
                if (key > key_lo) mn = min(mn, key);
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            mn = __reduce_min_sync(0xffffffffu, mn);
            if ((tid & 31) == 0) { atomicAdd(&s_cnt, cnt); atomicMin(&s_min, mn); }
            __syncthreads();
            const float hi = (s_cnt > (uint32_t)k_hi) ? med : key_float(s_min);
            med = __fmul_rn(__fadd_rn(med, hi), 0.5f);
        }
        {   // the frame's weakest bit: min |coefficient - median| (non-negative floats order like their bit patterns)
            uint32_t mg = 0x7f800000u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q * kHashThreads + tid < N) mg = min(mg, __float_as_uint(fabsf(__fsub_rn(mine[q], med))));
            mg = __reduce_min_sync(0xffffffffu, mg);
            if ((tid & 31) == 0) atomicMin(&s_margin, mg);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = q * kHashThreads + tid;
            if (q * kHashThreads < P.words * 32) {  // uniform per iteration: whole warps vote
                const bool bit = idx < N && mine[q] > med;
                const uint32_t word = __ballot_sync(0xffffffffu, bit);
                if ((tid & 31) == 0 && (idx >> 5) < P.words) bits_out[(size_t)f * P.words + (idx >> 5)] = word;
            }
        }
        __syncthreads();  // every warp's atomicMin into s_margin has landed
        if (tid == 0 && P.margin_out) P.margin_out[f] = __uint_as_float(s_margin);
    }
}

// hash_dist_norm = count_nonzero(cur != prev) / float(size * size); hashes points at the batch's first frame inside
// the ctx-wide array, so hashes[-words..-1] is the last frame of the previous batch.  NaN = no previous frame.
__global__ void hash_dist_kernel(const uint32_t* __restrict__ hashes, int words, int n, int first_has_prev, double size_sq,
                                 double* __restrict__ hash_dist) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    if (f == 0 && !first_has_prev) {
        hash_dist[0] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const uint32_t* cur = hashes + (long long)f * words;
    const uint32_t* prv = cur - words;
    int cnt = 0;
    for (int i = 0; i < words; ++i) cnt += __popc(cur[i] ^ prv[i]);
    hash_dist[f] = __ddiv_rn((double)cnt, size_sq);
}

// ----------------------------------------------------------------------------------- decision passes
struct DecisionState {
    long long c_last_above, c_merge_start;
    int c_init, c_merge_enabled, c_merge_triggered, pad0;
    long long a_last_cut;
    int a_init, pad1;
    long long h_last_cut;  // 0 == "not set" (Python falsiness of `if not self._last_scene_cut`)
    long long n_cuts[5];
    int overflow, pad2;
    // ThresholdDetector: last_scene_cut, last_fade {frame, type}, processed_frame
    long long t_last_scene_cut, t_fade_frame;
    int t_init, t_processed, t_fade_out, pad3;
    // HashDetector: _last_scene_cut
    long long x_last_cut;
    int x_init, pad4;
};

struct DecisionParams {
    int detectors;
    int content_min_scene_len, content_filter_mode;
    int adaptive_w, adaptive_min_scene_len;
    int hist_min_scene_len;
    double content_threshold;
    double adaptive_threshold, adaptive_min_content_val;
    double hist_threshold;  // already clamp(1 - t, 0, 1)
    long long max_cuts;
    double thresh_threshold;  // int(threshold) as double
    double thresh_fade_bias;
    int thresh_min_scene_len, thresh_method;  // method 0 FLOOR, 1 CEILING
    double hash_threshold;
    int hash_min_scene_len, pad;
    long long cuts_stride;  // cut list of detector d starts at cuts + d * cuts_stride (max_cuts; 0 = one shared list, stand-alone pass)
    int fresh_state, pad5;  // stand-alone pass: start from the zero state and do not touch `st` (no memset launch needed)
};

struct CutSink {
    long long* cuts;
    long long n, max_cuts;
    int overflow;
    __device__ __forceinline__ void emit(long long v) {
        if (n < max_cuts) cuts[n] = v;
        else overflow = 1;
        ++n;
    }
};

constexpr int kDecideThreads = 1024;

// grid = 5 blocks (content, adaptive, hist, threshold, hash) of kDecideThreads; frames [i_begin, i_end) are indices from first_frame_num.
// The threshold tests run in parallel into a shared bitmask; one thread then walks the sequential
// FlashFilter / min_scene_len state machine (A.5-A.7) with its state in registers, visiting only
// frames that can change it (set bits, or every frame while a MERGE burst is open).
__global__ void __launch_bounds__(kDecideThreads) decide_kernel(DecisionParams P, DecisionState* __restrict__ st, long long* __restrict__ cuts,
                              const double* __restrict__ content_val, const double* __restrict__ adaptive_val,
                              const double* __restrict__ adaptive_ratio, const double* __restrict__ hist_diff,
                              const double* __restrict__ average_rgb, const double* __restrict__ hash_dist,
                              long long first_frame_num, long long i_begin, long long i_end,
                              long long* __restrict__ mailbox, long long ticket) {
    constexpr int CH = 16384;
    __shared__ uint32_t bits[CH / 32];
    // frames of the chunk whose test is true, ascending: the sequential walk steps through this list instead of scanning the
    // bitmask a word at a time (one thread's dependent smem load + branches cost ~300 cycles per empty word: 100 of the 116 us
    // an 18 000-frame pass took, profiles/r02_decide_ncu.md)
    __shared__ uint16_t s_idx[CH];
    __shared__ int s_warp_sum[kDecideThreads / 32];
    __shared__ int s_count;
    const int det = blockIdx.x;
    if (!(P.detectors & (1 << det)) || i_end <= i_begin) return;
    const int tid = threadIdx.x;

    // thread 0's register copy of this detector's state
    CutSink sink{cuts + (size_t)det * P.cuts_stride, 0, P.max_cuts, 0};
    long long last = 0, merge_start = 0;
    int init = 0, merge_enabled = 0, merge_triggered = 0;
    if (tid == 0 && P.fresh_state) {
        if (!(det == 0 && !(P.content_min_scene_len > 0)) && det != 2) { init = 1; last = first_frame_num + i_begin; }
    } else if (tid == 0) {
        sink.n = st->n_cuts[det];
        if (det == 0) {
            last = st->c_last_above; merge_start = st->c_merge_start; init = st->c_init;
            merge_enabled = st->c_merge_enabled; merge_triggered = st->c_merge_triggered;
            if (!init && P.content_min_scene_len > 0) { init = 1; last = first_frame_num + i_begin; }
        } else if (det == 1) {
            last = st->a_last_cut; init = st->a_init;
            if (!init) { init = 1; last = first_frame_num + i_begin; }
        } else if (det == 2) {
            last = st->h_last_cut;
        } else if (det == 3) {
            last = st->t_last_scene_cut; init = st->t_init;
            if (!init) { init = 1; last = first_frame_num + i_begin; }
        } else {
            last = st->x_last_cut; init = st->x_init;
            if (!init) { init = 1; last = first_frame_num + i_begin; }
        }
    }
    const int L = det == 0 ? P.content_min_scene_len : det == 1 ? P.adaptive_min_scene_len
                : det == 2 ? P.hist_min_scene_len : det == 3 ? P.thresh_min_scene_len : P.hash_min_scene_len;
    // ThresholdDetector state (thread 0): merge_start doubles as last_fade.frame
    int t_processed = 0, t_fade_out = 0;
    if (tid == 0 && det == 3 && !P.fresh_state) { t_processed = st->t_processed; t_fade_out = st->t_fade_out; merge_start = st->t_fade_frame; }

    for (long long c0 = i_begin; c0 < i_end; c0 += CH) {
        const long long c1 = (c0 + CH < i_end) ? c0 + CH : i_end;
        // threshold tests of the chunk, eight frames per thread per pass with the loads issued before the first use
        // (a 256-thread block walking 8192 frames one load at a time was latency-bound: ~30 us per chunk)
        for (int j0 = 0; j0 < CH; j0 += 8 * (int)blockDim.x) {
            bool bit[8];
            if (det == 1) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long long i = c0 + j0 + u * (int)blockDim.x + tid;
                    bit[u] = false;
                    if (i < c1 && i >= 2LL * P.adaptive_w) {
                        const long long t = i - P.adaptive_w;
                        const double r = adaptive_ratio ? adaptive_ratio[t] : adaptive_ratio_at(adaptive_val, t, P.adaptive_w, P.adaptive_min_content_val);
                        bit[u] = r >= P.adaptive_threshold && adaptive_val[t] >= P.adaptive_min_content_val;
                    }
                }
            } else {
                const double* __restrict__ src = det == 0 ? content_val : det == 2 ? hist_diff : det == 4 ? hash_dist : average_rgb;
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long long i = c0 + j0 + u * (int)blockDim.x + tid;
                    v[u] = i < c1 ? src[i] : __longlong_as_double(0x7ff8000000000000LL);  // NaN: every test below is false
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    bit[u] = det == 0 ? v[u] >= P.content_threshold
                           : det == 2 ? v[u] <= P.hist_threshold   // NaN (no previous frame) compares false
                           : det == 4 ? v[u] >= P.hash_threshold   // NaN (no previous frame) compares false
                                      : v[u] < P.thresh_threshold; // "below the fade threshold"
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * (int)blockDim.x + tid;
                const uint32_t word = __ballot_sync(0xffffffffu, bit[u]);
                if ((tid & 31) == 0 && j < CH) bits[j >> 5] = word;
            }
        }
        __syncthreads();
        if (det != 3) {
            // compaction: thread t owns word t (CH / 32 = 512 words <= kDecideThreads): block-wide exclusive scan of the popcounts
            static_assert(CH / 32 <= kDecideThreads, "one word per thread");
            const int nj_all = (int)(c1 - c0);
            uint32_t word = 0;
            if (tid < CH / 32) {
                word = bits[tid];
                const int first = tid * 32;
                if (first >= nj_all) word = 0;
                else if (first + 32 > nj_all) word &= (1u << (nj_all - first)) - 1u;   // frames beyond the chunk's end carry no test
            }
            const int cnt = __popc(word);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += v;
            }
            if ((tid & 31) == 31) s_warp_sum[tid >> 5] = incl;
            __syncthreads();
            int base = 0;
            for (int w = 0; w < (tid >> 5); ++w) base += s_warp_sum[w];
            int o = base + incl - cnt;
            while (word) {
                const int bpos = __ffs(word) - 1;
                s_idx[o++] = (uint16_t)(tid * 32 + bpos);
                word &= word - 1;
            }
            if (tid == kDecideThreads - 1) s_count = base + incl;
            __syncthreads();
        }
        if (tid == 0) {
            const int nj = (int)(c1 - c0);
            int j = 0;
            if (det == 3) {
                // ThresholdDetector.process_frame: fades are crossings of the threshold; FLOOR fades out below it,
                // CEILING fades out at/above it.  Only frames where the "faded out" predicate flips matter.
                while (j < nj) {
                    const bool below = (bits[j >> 5] >> (j & 31)) & 1u;
                    const long long fn = first_frame_num + c0 + j;
                    if (!t_processed) {
                        merge_start = 0;        // last_fade.frame
                        t_fade_out = below;     // first frame: type is 'out' iff frame_avg < threshold (both methods)
                        t_processed = 1;
                        ++j;
                        continue;
                    }
                    const bool out_now = (P.thresh_method == 0) ? below : !below;
                    if (out_now == (bool)t_fade_out) {  // no transition: skip ahead to the next flip of `below`
                        uint32_t word = bits[j >> 5] >> (j & 31);
                        if (!below) { if (word == 0) { j = (j | 31) + 1; continue; } j += __ffs(word) - 1; }
                        else {
                            word = ~word & (0xffffffffu >> (j & 31));
                            if (word == 0) { j = (j | 31) + 1; continue; }
                            j += __ffs(word) - 1;
                        }
                        continue;
                    }
                    if (!t_fade_out) {  // was 'in', now faded out: remember where
                        t_fade_out = 1;
                        merge_start = fn;
                    } else {            // was 'out', now faded in: emit the split point if the scene is long enough
                        if ((fn - last) >= L) {
                            const long long f_out = merge_start;
                            const long long bias = (long long)(P.thresh_fade_bias * (double)(fn - f_out));
                            sink.emit((long long)((double)(fn + f_out + bias) / 2.0));
                            last = fn;
                        }
                        t_fade_out = 0;
                        merge_start = fn;
                    }
                    ++j;
                }
                j = nj;
            }
            const int m = det != 3 ? s_count : 0;   // above-threshold frames of this chunk
            int pnext = 0;                            // next entry of s_idx; invariant: s_idx[pnext] >= j
            while (j < nj) {
                if (det == 0 && merge_triggered) {
                    // Open MERGE burst: nothing changes between above-threshold frames, and the burst closes at
                    // the first below-threshold frame fstar = last + L once (last - merge_start) >= L.
                    const int ja = pnext < m ? (int)s_idx[pnext] : nj;
                    const long long fn_a = first_frame_num + c0 + ja;  // next above frame (or end of chunk)
                    if ((last - merge_start) >= L) {
                        const long long fn_j = first_frame_num + c0 + j;
                        const long long fstar = (last + L > fn_j) ? last + L : fn_j;
                        if (fstar < fn_a) {
                            merge_triggered = 0;
                            sink.emit(last);
                            j = (int)(fstar - first_frame_num - c0) + 1;
                            continue;
                        }
                    }
                    if (ja >= nj) break;
                    last = fn_a;  // above frame inside the burst only advances last_above
                    j = ja + 1;
                    ++pnext;
                    continue;
                }
                const bool every_frame = (det == 2 && last == 0);
                if (!every_frame) {  // jump to the next frame whose test is true
                    if (pnext >= m) break;
                    j = (int)s_idx[pnext];
                }
                if (pnext < m && (int)s_idx[pnext] == j) ++pnext;
                const bool above = (bits[j >> 5] >> (j & 31)) & 1u;
                const long long fn = first_frame_num + c0 + j;
                ++j;
                if (det == 0) {
                    if (!(L > 0)) { if (above) sink.emit(fn); continue; }
                    const bool met = (fn - last) >= L;
                    if (P.content_filter_mode == 1) {  // SUPPRESS (legacy min_scene_len)
                        if (above && met) { last = fn; sink.emit(fn); }
                        continue;
                    }
                    if (above) last = fn;
                    if (merge_triggered) {
                        if (met && !above && (last - merge_start) >= L) { merge_triggered = 0; sink.emit(last); }
                        continue;
                    }
                    if (!above) continue;
                    if (met) { merge_enabled = 1; sink.emit(fn); continue; }
                    if (merge_enabled) { merge_triggered = 1; merge_start = fn; }
                } else if (det == 1) {
                    if (above && (fn - last) >= L) { last = fn - P.adaptive_w; sink.emit(last); }
                } else if (det == 2) {
                    if (last == 0) last = fn;  // `if not self._last_scene_cut` (0 is falsy)
                    if (above && (fn - last) >= L) { sink.emit(fn); last = fn; }
                } else {  // HashDetector: `_last_scene_cut is None` -> first frame; plain min_scene_len rule
                    if (above && (fn - last) >= L) { sink.emit(fn); last = fn; }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (!P.fresh_state) {
            st->n_cuts[det] = sink.n;
            if (sink.overflow) st->overflow = 1;
        }
        // per-frame path: cut count and overflow flag straight into host-mapped pinned memory (no D2H copy), then this
        // detector's completion ticket -- the host polls the tickets instead of paying a stream synchronise
        if (mailbox) {
            mailbox[det] = sink.n;
            if (sink.overflow) mailbox[5] = 1;
            __threadfence_system();
            *reinterpret_cast<volatile long long*>(mailbox + 8 + det) = ticket;
        }
        if (P.fresh_state) return;
        if (det == 0) {
            st->c_last_above = last; st->c_merge_start = merge_start; st->c_init = init;
            st->c_merge_enabled = merge_enabled; st->c_merge_triggered = merge_triggered;
        } else if (det == 1) {
            st->a_last_cut = last; st->a_init = init;
        } else if (det == 2) {
            st->h_last_cut = last;
        } else if (det == 3) {
            st->t_last_scene_cut = last; st->t_init = init; st->t_processed = t_processed;
            st->t_fade_out = t_fade_out; st->t_fade_frame = merge_start;
        } else {
            st->x_last_cut = last; st->x_init = init;
        }
    }
}

__global__ void fill_nan_kernel(double* p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace esd
