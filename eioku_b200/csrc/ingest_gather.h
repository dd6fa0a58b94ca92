// ingest_gather.h -- host-side tap gather of the ingest path (no CUDA in here: unit-tested on the CPU by
// tests/test_host_gather.py through tests/gather_shim.cpp).
//
// For every touched source row the fused kernel only reads, per destination column, the two horizontal taps:
//   BGR24: 6 bytes [tap0 B G R, tap1 B G R] at byte 3 * x0         -> written at byte 6 * d of the gathered row
//   NV12 Y row: 2 bytes [Y(x0), Y(x0 + 1)]                          -> written at byte 2 * d
//   NV12 UV row: the chroma pairs of the two taps [U V, U' V']      -> written at byte 4 * d
//   I420 (planar chroma, contiguous frame: Y rows of `pitch`, then src_height / 2 U rows and as many V rows of pitch / 2):
//        Y rows as NV12's; a chroma row reads U[c], V[c] from the two planes and writes the same interleaved pairs at 4 * d,
//        so the device sees exactly the NV12 tap layout
// (x0 = source column of tap 0; a clamped last column repeats the last pixel -- its tap 1 has weight 0).
// The rows of a frame are laid out in the order of the touched-row list with one pitch (`tap_row_bytes`).
#pragma once
#include <emmintrin.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

namespace esd {

struct GatherSpec {
    int dst_w;              // destination columns
    int tap_row_bytes;      // pitch of a gathered row (multiple of 16)
    int row_bytes;          // bytes of a source row (src_w * 3, or src_w for NV12)
    const int* off;         // [dst_w] byte offset of tap 0 in a BGR24 row (3 * x0), or x0 for NV12
    const int32_t* touched; // [n_touched] source rows, ascending (NV12: Y rows, then UV rows as src_height + r)
    int64_t n_touched;
    int n_touched_y;        // NV12: how many of them are Y rows
    bool nv12;
    bool i420;              // planar chroma (nv12 is set too: the gathered layout is NV12's)
    int src_height;         // I420: rows of the Y plane
    int prefetch_bytes;     // rolling software prefetch distance (0 = touch the next row's pages only)
    bool nt_stores;         // assemble each row in L1 and stream it out with non-temporal stores
    int streams;            // BGR24: rows gathered in lock-step per thread (1 = one after another; 8 = the default of the ingest path)
    int prefetch_bytes_multi;  // prefetch distance of the multi-stream loop (each stream runs 1/streams as fast: shorter than prefetch_bytes)
    int prefetch_hint;      // experiment knob of the BGR24 loop: 0 = into L1 (prefetcht0), 1 = L2 (prefetcht1), 2 = L3 (prefetcht2), 3 = non-temporal
};

constexpr int kMaxTapRow = 6 * 1024 + 16;  // dst_w <= 1024 in resizing contexts
constexpr int kGatherStreams = 8;          // rows in lock-step in the multi-stream BGR24 loop
constexpr int kGatherStreamGap = 2;        // consecutive items per stream: one block = 16 items

// BGR24, multi-stream: one core sustains ~6.5 GB/s walking ONE row at a time (a dozen cache misses in flight, the hardware
// prefetchers restarting at every row); gathering eight rows in lock-step keeps eight independent miss streams in flight and
// measured +27 % per core (scripts/gather_probe.py, profiles/r02_gather_streams.log).  Items [lo, lo + 16) of the item sequence:
// stream s takes items lo + 2 s and lo + 2 s + 1.  Rows are assembled in L1 and streamed out (non-temporal), as in the loop below.
inline void gather_tap_rows_bgr_block(const GatherSpec& g, const uint8_t* src, int64_t frame_stride, int64_t pitch, uint8_t* dst_base,
                                      int64_t lo) {
    const int dw = g.dst_w, trb = g.tap_row_bytes, rb = g.row_bytes, pf = g.prefetch_bytes_multi;
    const int* off = g.off;
    // the assembled rows lie tstride apart, NOT a multiple of 4 KB plus a few bytes: with a stride of kMaxTapRow (6 160 B) the eight
    // stores of a column alias in their low 12 address bits and the loop ran 2.3x slower than one row at a time
    alignas(64) uint8_t tmp_all[kGatherStreams * (kMaxTapRow + 192)];
    const int tstride = ((trb + 63) & ~63) + 192;
    uint8_t* tmp[kGatherStreams];
    for (int s = 0; s < kGatherStreams; ++s) tmp[s] = tmp_all + s * tstride;
    int dfast = 0;  // columns whose 8-byte move stays inside the row
    while (dfast < dw - 1 && off[dfast] + 8 <= rb) ++dfast;
    for (int j = 0; j < kGatherStreamGap; ++j) {
        const uint8_t* sr[kGatherStreams];
        for (int s = 0; s < kGatherStreams; ++s) {
            const int64_t it = lo + s * kGatherStreamGap + j;
            const int64_t f = it / g.n_touched, i = it - f * g.n_touched;
            sr[s] = src + f * frame_stride + (int64_t)g.touched[i] * pitch;
        }
        for (int d = 0; d < dfast; ++d) {
            if (pf > 0 && (d & 3) == 0)
                for (int s = 0; s < kGatherStreams; ++s) __builtin_prefetch(sr[s] + off[d] + pf, 0, 2);  // runs on into the next row: harmless
            for (int s = 0; s < kGatherStreams; ++s) {
                uint64_t v;
                memcpy(&v, sr[s] + off[d], 8);
                memcpy(tmp[s] + 6 * d, &v, 8);  // the 2 spare bytes are overwritten by column d + 1
            }
        }
        for (int s = 0; s < kGatherStreams; ++s) {
            for (int d = dfast; d < dw; ++d) {
                const int nbytes = std::min(6, rb - off[d]);  // a clamped last column only has tap 0 (tap 1 weighs 0)
                memcpy(tmp[s] + 6 * d, sr[s] + off[d], (size_t)nbytes);
                if (nbytes < 6) memset(tmp[s] + 6 * d + nbytes, 0, (size_t)(6 - nbytes));
            }
            if (trb > 6 * dw) memset(tmp[s] + 6 * dw, 0, (size_t)(trb - 6 * dw));
            uint8_t* out_row = dst_base + (lo + s * kGatherStreamGap + j) * trb;
            for (int b = 0; b < trb; b += 16)
                _mm_stream_si128(reinterpret_cast<__m128i*>(out_row + b), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp[s] + b)));
        }
    }
}

// Gathers items [lo, hi) of the (frame-major, touched-row-minor) item sequence of `src` (frames `frame_stride` apart, rows
// `pitch` apart) into dst_base + item * tap_row_bytes.  dst_base must be 16-byte aligned when nt_stores is set.
inline void gather_tap_rows(const GatherSpec& g, const uint8_t* src, int64_t frame_stride, int64_t pitch, uint8_t* dst_base,
                            int64_t lo, int64_t hi) {
    const int dw = g.dst_w, trb = g.tap_row_bytes, rb = g.row_bytes, pf = g.prefetch_bytes;
    const int* off = g.off;
    const bool nt = g.nt_stores && trb <= kMaxTapRow;
    if (!g.nv12 && nt && g.streams == kGatherStreams) {
        constexpr int64_t kBlock = kGatherStreams * kGatherStreamGap;
        for (; lo + kBlock <= hi; lo += kBlock) gather_tap_rows_bgr_block(g, src, frame_stride, pitch, dst_base, lo);
        if (lo >= hi) { _mm_sfence(); return; }
    }
    alignas(64) uint8_t tmp_row[kMaxTapRow];
    for (int64_t it = lo; it < hi; ++it) {
        const int64_t f = it / g.n_touched, i = it - f * g.n_touched;
        const uint8_t* sr = src + f * frame_stride + (int64_t)g.touched[i] * pitch;
        const bool planar_chroma = g.i420 && i >= g.n_touched_y;
        if (planar_chroma)  // U row of chroma row r; the V row lies src_height / 2 chroma rows further
            sr = src + f * frame_stride + (int64_t)g.src_height * pitch + (int64_t)(g.touched[i] - g.src_height) * (pitch / 2);
        uint8_t* const out_row = dst_base + it * trb;
        // the ring slot is written once and read by the DMA engine: no read-for-ownership of the gathered bytes
        uint8_t* dr = nt ? tmp_row : out_row;
        const uint8_t* nx = sr;  // where the prefetch runs on to after this row
        if (it + 1 < hi) {
            const int64_t f1 = (it + 1) / g.n_touched, i1 = (it + 1) - f1 * g.n_touched;
            nx = src + f1 * frame_stride + (int64_t)g.touched[i1] * pitch;
            if (pf == 0)  // hardware prefetchers stop at 4 KB boundaries: touch the next row's pages early
                for (int b = 0; b < rb; b += 2048) __builtin_prefetch(nx + b, 0, 1);
        }
        int d = 0;
        if (g.nv12 && i < g.n_touched_y) {
            // luma: one unaligned 16-bit move per column
            for (; d < dw && off[d] + 2 <= rb; ++d) {
                if (pf > 0 && (d & 7) == 0) {
                    const int o = off[d] + pf;
                    __builtin_prefetch(o < rb ? sr + o : nx + (o - rb), 0, 3);
                }
                uint16_t v;
                memcpy(&v, sr + off[d], 2);
                memcpy(dr + 2 * d, &v, 2);
            }
            for (; d < dw; ++d) dr[2 * d] = dr[2 * d + 1] = sr[rb - 1];  // clamped last column
            if (nt) memset(dr + 2 * dw, 0, (size_t)(trb - 2 * dw));
        } else if (planar_chroma) {
            const uint8_t* sv = sr + (int64_t)(g.src_height / 2) * (pitch / 2);
            const int cw = rb / 2;
            for (; d < dw; ++d) {
                const int x0 = off[d];
                if (pf > 0 && (d & 15) == 0) {
                    const int o = (x0 + pf) >> 1;
                    if (o < cw) { __builtin_prefetch(sr + o, 0, 3); __builtin_prefetch(sv + o, 0, 3); }
                }
                const int c0 = x0 >> 1, c1 = ((x0 & 1) && c0 + 1 < cw) ? c0 + 1 : c0;
                const uint32_t v = (uint32_t)sr[c0] | ((uint32_t)sv[c0] << 8) | ((uint32_t)sr[c1] << 16) | ((uint32_t)sv[c1] << 24);
                memcpy(dr + 4 * d, &v, 4);
            }
        } else if (g.nv12) {
            // chroma: an odd x0 takes two consecutive pairs (one 32-bit move), an even x0 the same pair twice
            for (; d < dw; ++d) {
                const int x0 = off[d];
                if (pf > 0 && (d & 7) == 0) {
                    const int o = x0 + pf;
                    __builtin_prefetch(o < rb ? sr + o : nx + (o - rb), 0, 3);
                }
                const int c0 = x0 & ~1;
                uint32_t v;
                if ((x0 & 1) && c0 + 4 <= rb) {
                    memcpy(&v, sr + c0, 4);
                } else {
                    uint16_t a;
                    memcpy(&a, sr + c0, 2);
                    v = (uint32_t)a | ((uint32_t)a << 16);
                }
                memcpy(dr + 4 * d, &v, 4);
            }
        } else {
            for (; d < dw - 1 && off[d] + 8 <= rb; ++d) {  // 8-byte moves; the 2 spare bytes are overwritten by d + 1
                if (pf > 0 && (d & 3) == 0) {
                    const int o = off[d] + pf;
                    const uint8_t* pa = o < rb ? sr + o : nx + (o - rb);
                    switch (g.prefetch_hint) {
                        case 1: __builtin_prefetch(pa, 0, 2); break;
                        case 2: __builtin_prefetch(pa, 0, 1); break;
                        case 3: __builtin_prefetch(pa, 0, 0); break;
                        default: __builtin_prefetch(pa, 0, 3); break;
                    }
                }
                uint64_t v;
                memcpy(&v, sr + off[d], 8);
                memcpy(dr + 6 * d, &v, 8);
            }
            for (; d < dw; ++d) {
                const int nbytes = std::min(6, rb - off[d]);  // a clamped last column only has tap 0 (tap 1 weighs 0)
                memcpy(dr + 6 * d, sr + off[d], (size_t)nbytes);
                if (nbytes < 6) memset(dr + 6 * d + nbytes, 0, (size_t)(6 - nbytes));
            }
        }
        if (nt)
            for (int b = 0; b < trb; b += 16)
                _mm_stream_si128(reinterpret_cast<__m128i*>(out_row + b), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp_row + b)));
    }
    if (nt) _mm_sfence();
}

}  // namespace esd
