// Test shim (CPU): runs eioku_b200/csrc/jpeg_core.h + jpeg_parse.h -- the arithmetic of libesd_decode.so's own JPEG decoder --
// on the host so that tests/test_host_jpeg.py can compare it with cv2.imdecode bit for bit.  Not part of the product.
#include "../eioku_b200/csrc/jpeg_parse.h"

#include <vector>

using namespace esdjpeg;

static int g_last_rounds = 0;
extern "C" {
int shim_last_rounds() { return g_last_rounds; }
// returns 0 and fills width / height, or -1 (message in err, 256 bytes)
int shim_jpeg_info(const uint8_t* data, long n, int* width, int* height, char* err) {
    JpegHeader h;
    std::string e;
    if (!parse_jpeg(data, (size_t)n, &h, &e)) { strncpy(err, e.c_str(), 255); err[255] = 0; return -1; }
    *width = h.width; *height = h.height;
    return 0;
}
// decodes into out[height][width][3] BGR.  flat != 0: the flat scan decoder (what the GPU's fast path runs: unstuffed input, one
// loop, coefficient buffer + separate IDCT pass); flat == 0: the nested decode_block form.
int shim_jpeg_decode(const uint8_t* data, long n, uint8_t* out, char* err, int flat) {
    JpegHeader h;
    std::string e;
    if (!parse_jpeg(data, (size_t)n, &h, &e)) { strncpy(err, e.c_str(), 255); err[255] = 0; return -1; }
    const FrameGeometry g = geometry_of(h);
    const int yw = g.yblocks_x * 8, yh = g.mcus_y * 16, cw = g.cblocks_x * 8, ch = g.mcus_y * 8;
    std::vector<uint8_t> Y((size_t)yw * yh), Cb((size_t)cw * ch), Cr((size_t)cw * ch);
    if (flat) {
        std::vector<uint32_t> clean_words(h.scan_len / 4 + 4, 0u);
        bool clean = true;
        unstuff_scan(data + h.scan_offset, h.scan_len, reinterpret_cast<uint8_t*>(clean_words.data()), &clean);
        if (!clean || g.restart_interval) { strncpy(err, "restart markers: not for the flat decoder", 255); return -2; }
        const size_t nblocks = (size_t)6 * g.mcus_x * g.mcus_y;
        std::vector<int16_t> coef((nblocks + 1) * 64, 0);  // + the spare block
        // the library's rule: a list of 32 entries per block holds any picture whose scan has at most that many BITS (a symbol that
        // carries a coefficient is at least one bit long); larger scans keep the dense hand-off
        if (flat < 0 && clean_words.size() * 32 > nblocks * 32) flat = -flat;
        if (flat < 0) {    // the same scheme with the SPARSE hand-off (entry list + block offsets), expanded block by block: thread budget = -flat
            const int cap = (int)nblocks * 32;
            std::vector<uint32_t> list((size_t)cap + 1, 0xdeadbeefu), bstart(nblocks + 2, 0xffffffffu);
            g_last_rounds = decode_scan_parallel_host(clean_words.data(), (int)clean_words.size(), h.huff, h.td, h.ta, kNaturalOrderHost,
                                                      g.mcus_x * g.mcus_y, reinterpret_cast<int16_t*>(list.data()), -flat, bstart.data(), cap);
            std::vector<int16_t> ref((nblocks + 1) * 64, 0);
            decode_scan_flat(clean_words.data(), (int)clean_words.size(), h.huff, h.td, h.ta, kNaturalOrderHost, g.mcus_x * g.mcus_y, ref.data());
            for (size_t b = 0; b < nblocks; ++b) {
                expand_block(list.data(), bstart.data(), (int)b, cap, &coef[b * 64]);
                if (memcmp(&ref[b * 64], &coef[b * 64], 128) != 0) { snprintf(err, 255, "sparse hand-off differs from the flat loop at block %zu", b); return -4; }
            }
        } else
        if (flat >= 2) {   // the many-threads-per-picture scheme (host statement), flat = the thread budget; coefficients must equal the flat loop's
            g_last_rounds = decode_scan_parallel_host(clean_words.data(), (int)clean_words.size(), h.huff, h.td, h.ta, kNaturalOrderHost,
                                                      g.mcus_x * g.mcus_y, coef.data(), flat);
            std::vector<int16_t> ref((nblocks + 1) * 64, 0);
            decode_scan_flat(clean_words.data(), (int)clean_words.size(), h.huff, h.td, h.ta, kNaturalOrderHost, g.mcus_x * g.mcus_y, ref.data());
            if (memcmp(ref.data(), coef.data(), nblocks * 64 * sizeof(int16_t)) != 0) { strncpy(err, "parallel decode differs from the flat loop", 255); return -3; }
        } else
        decode_scan_flat(clean_words.data(), (int)clean_words.size(), h.huff, h.td, h.ta, kNaturalOrderHost, g.mcus_x * g.mcus_y, coef.data());
        for (size_t b = 0; b < nblocks; ++b) {  // coefficients are in decoding order: block_position maps them into the planes
            int comp, bx, by;
            block_position((int)b, g.mcus_x, &comp, &bx, &by);
            const int bw = comp == 0 ? g.yblocks_x : g.cblocks_x;
            uint8_t* plane = comp == 0 ? Y.data() : (comp == 1 ? Cb.data() : Cr.data());
            idct_islow(&coef[b * 64], h.quant[h.tq[comp]], plane + (size_t)by * 8 * (bw * 8) + bx * 8, bw * 8);
        }
    } else {
    BitReader br;
    br.init(data + h.scan_offset, (int)h.scan_len);
    int pred[3] = {0, 0, 0};
    int16_t coef[64];
    int mcu = 0;
    for (int my = 0; my < g.mcus_y; ++my)
        for (int mx = 0; mx < g.mcus_x; ++mx, ++mcu) {
            if (g.restart_interval && mcu && mcu % g.restart_interval == 0) { br.restart(); pred[0] = pred[1] = pred[2] = 0; }
            for (int b = 0; b < 4; ++b) {
                memset(coef, 0, sizeof coef);
                decode_block(br, h.huff.dc[h.td[0]], h.huff.ac[h.ta[0]], kNaturalOrderHost, pred[0], coef);
                idct_islow(coef, h.quant[h.tq[0]], &Y[(size_t)(my * 16 + (b >> 1) * 8) * yw + mx * 16 + (b & 1) * 8], yw);
            }
            for (int c = 1; c < 3; ++c) {
                memset(coef, 0, sizeof coef);
                decode_block(br, h.huff.dc[h.td[c]], h.huff.ac[h.ta[c]], kNaturalOrderHost, pred[c], coef);
                idct_islow(coef, h.quant[h.tq[c]], &(c == 1 ? Cb : Cr)[(size_t)(my * 8) * cw + mx * 8], cw);
            }
        }
    }
    const int rcw = (h.width + 1) / 2, rch = (h.height + 1) / 2;  // real (downsampled) chroma size
    for (int y = 0; y < h.height; ++y)
        for (int x = 0; x < h.width; ++x)
            ycc_to_bgr(Y[(size_t)y * yw + x], fancy_chroma(Cb.data(), cw, rcw, rch, x, y), fancy_chroma(Cr.data(), cw, rcw, rch, x, y),
                       out + ((size_t)y * h.width + x) * 3);
    return 0;
}
}
