// jpeg_parse.h -- host-side JPEG header walk for libesd_decode.so's own decoder (CUDA-free; also used by tests/jpeg_shim.cpp).
// Accepts what OpenCV / ffmpeg write into Motion-JPEG AVI files: baseline (SOF0) or extended-sequential Huffman (SOF1), 8-bit,
// three components, 4:2:0 (luma 2x2, chroma 1x1), one interleaved scan, tables present in the picture.  Everything else is
// reported as unsupported and the caller uses another decoder.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>

#include "jpeg_core.h"

namespace esdjpeg {

struct JpegHeader {
    int width = 0, height = 0;
    uint16_t quant[4][64];   // natural order
    bool have_quant[4] = {false, false, false, false};
    int tq[3] = {0, 0, 0};   // quantisation table of each component
    int td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    ScanTables huff;
    bool have_dc[2] = {false, false}, have_ac[2] = {false, false};
    size_t dht_pos[2][2] = {{0, 0}, {0, 0}};  // [class: 0 DC, 1 AC][table id]: offset of the 16 code-length counts inside the picture
    int dht_nvals[2][2] = {{0, 0}, {0, 0}};   // number of symbol values that follow the counts
    int restart_interval = 0;
    size_t scan_offset = 0, scan_len = 0;  // entropy-coded data inside the picture's bytes
    // raw DHT payload bytes concatenated: pictures of one file are compared by these to share device tables
    std::string dht_bytes;
};

// build_huff = false skips the construction of the Huffman decoding tables (the caller already has the set `dht_bytes` names)
inline bool parse_jpeg(const uint8_t* d, size_t n, JpegHeader* h, std::string* err, bool build_huff = true) {
    auto fail = [&](const char* m) { if (err) *err = m; return false; };
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("no SOI");
    size_t p = 2;
    bool sof = false;
    int comp_id[3] = {0, 0, 0};
    while (p + 4 <= n) {
        if (d[p] != 0xFF) return fail("marker expected");
        while (p < n && d[p] == 0xFF) ++p;
        if (p >= n) break;
        const int m = d[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return fail("EOI before SOS");
        if (p + 2 > n) break;
        const size_t len = ((size_t)d[p] << 8) | d[p + 1];
        if (len < 2 || p + len > n) return fail("truncated segment");
        const uint8_t* s = d + p + 2;
        const size_t sl = len - 2;
        if (m == 0xDB) {  // DQT
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, t = s[q] & 15;
                ++q;
                if (t > 3 || q + (pq ? 128 : 64) > sl) return fail("bad DQT");
                for (int k = 0; k < 64; ++k) {
                    const int v = pq ? ((s[q] << 8) | s[q + 1]) : s[q];
                    q += pq ? 2 : 1;
                    h->quant[t][kNaturalOrderHost[k]] = (uint16_t)v;
                }
                h->have_quant[t] = true;
            }
        } else if (m == 0xC4) {  // DHT
            h->dht_bytes.append((const char*)s, sl);
            size_t q = 0;
            while (q + 17 <= sl) {
                const int tc = s[q] >> 4, t = s[q] & 15;
                int total = 0;
                for (int i = 0; i < 16; ++i) total += s[q + 1 + i];
                if (tc > 1 || t > 1 || q + 17 + total > sl) return fail("bad DHT (baseline allows tables 0 and 1)");
                {   // cheap validity check (Kraft inequality) for tables that are built elsewhere (on the device)
                    int code = 0;
                    for (int i = 0; i < 16; ++i) { code = (code + s[q + 1 + i]) << 1; if (code > (2 << (i + 1))) return fail("over-subscribed Huffman table"); }
                    if (total > 256) return fail("bad DHT");
                }
                h->dht_pos[tc][t] = (size_t)(s - d) + q + 1;
                h->dht_nvals[tc][t] = total;
                HuffTable* tb = tc ? &h->huff.ac[t] : &h->huff.dc[t];
                if (build_huff && !build_huff_table(s + q + 1, s + q + 17, total, tb)) return fail("invalid Huffman table");
                (tc ? h->have_ac : h->have_dc)[t] = true;
                q += 17 + total;
            }
        } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
            if (sl < 15 || s[0] != 8) return fail("only 8-bit samples");
            h->height = (s[1] << 8) | s[2];
            h->width = (s[3] << 8) | s[4];
            if (s[5] != 3) return fail("only three-component pictures");
            for (int c = 0; c < 3; ++c) {
                comp_id[c] = s[6 + 3 * c];
                const int hs = s[7 + 3 * c] >> 4, vs = s[7 + 3 * c] & 15;
                h->tq[c] = s[8 + 3 * c] & 3;
                if (c == 0 ? (hs != 2 || vs != 2) : (hs != 1 || vs != 1)) return fail("only 4:2:0 subsampling (2x2, 1x1, 1x1)");
            }
            sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return fail("progressive / lossless / arithmetic JPEG");
        } else if (m == 0xDD) {  // DRI
            if (sl < 2) return fail("bad DRI");
            h->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {  // SOS
            if (!sof) return fail("SOS before SOF");
            if (sl < 10 || s[0] != 3) return fail("only one interleaved three-component scan");
            for (int c = 0; c < 3; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) return fail("scan components out of order");
                h->td[c] = s[2 + 2 * c] >> 4;
                h->ta[c] = s[2 + 2 * c] & 15;
                if (h->td[c] > 1 || h->ta[c] > 1 || !h->have_dc[h->td[c]] || !h->have_ac[h->ta[c]]) return fail("scan names a missing Huffman table");
                if (!h->have_quant[h->tq[c]]) return fail("missing quantisation table");
            }
            if (s[7] != 0 || s[8] != 63 || s[9] != 0) return fail("not a sequential scan");
            h->scan_offset = p + len;
            h->scan_len = n - h->scan_offset;
            if (h->width < 1 || h->height < 1) return fail("empty picture");
            return true;
        }
        p += len;
    }
    return fail("no SOS");
}

inline FrameGeometry geometry_of(const JpegHeader& h) {
    FrameGeometry g;
    g.width = h.width; g.height = h.height;
    g.mcus_x = (h.width + 15) / 16; g.mcus_y = (h.height + 15) / 16;
    g.yblocks_x = 2 * g.mcus_x; g.cblocks_x = g.mcus_x;
    g.restart_interval = h.restart_interval;
    return g;
}

}  // namespace esdjpeg
