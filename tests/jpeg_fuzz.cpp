// Fuzz harness (CPU, AddressSanitizer + UBSan): the entropy-decoding code libesd_decode.so runs on the device (csrc/jpeg_core.h:
// decode_scan_flat, decode_span in all three modes, expand_block) fed with DAMAGED scans -- bit flips, random spans, truncation,
// pure noise -- behind a real picture's headers and tables, in buffers of exactly the sizes the library allocates.  Garbage in,
// garbage out is fine; an out-of-bounds read or write, a shift by a negative count or an endless loop is not.  compute-sanitizer
// being closed on the GPU pool, this is where out-of-bounds READS of the decoder would show (the red-zone allocator sees writes).
// Not part of the product.  usage: jpeg_fuzz <picture.jpg> <seed> <iterations>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <vector>

#include "../eioku_b200/csrc/jpeg_parse.h"

using namespace esdjpeg;

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> jpg(1 << 24);
    jpg.resize(fread(jpg.data(), 1, jpg.size(), f));
    fclose(f);
    std::mt19937 rng((unsigned)atoi(argv[2]));
    const int iters = atoi(argv[3]);
    JpegHeader h;
    std::string err;
    if (!parse_jpeg(jpg.data(), jpg.size(), &h, &err)) { fprintf(stderr, "parse: %s\n", err.c_str()); return 2; }
    const FrameGeometry g = geometry_of(h);
    const int n_mcus = g.mcus_x * g.mcus_y, nblocks = 6 * n_mcus;
    long long total_rounds = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> scan(jpg.begin() + h.scan_offset, jpg.begin() + h.scan_offset + h.scan_len);
        const int mode = it % 5;
        if (mode == 0) {  // a few bit flips
            for (int k = 0; k < 1 + (int)(rng() % 8) && !scan.empty(); ++k) scan[rng() % scan.size()] ^= (uint8_t)(1u << (rng() % 8));
        } else if (mode == 1 && scan.size() > 8) {  // a random span
            const size_t a = rng() % scan.size(), n = 1 + rng() % std::min<size_t>(512, scan.size() - a);
            for (size_t i = 0; i < n; ++i) scan[a + i] = (uint8_t)rng();
        } else if (mode == 2) {  // truncation
            scan.resize(rng() % (scan.size() + 1));
        } else if (mode == 3) {  // noise of the same length (FF bytes included: markers appear mid-scan)
            for (auto& b : scan) b = (uint8_t)rng();
        } else {  // runs of one byte value
            const uint8_t v = (uint8_t)rng();
            const size_t a = scan.empty() ? 0 : rng() % scan.size();
            for (size_t i = a; i < scan.size() && i < a + 4096; ++i) scan[i] = v;
        }
        // staged exactly like esd_mjpeg_read: unstuffed, 4-byte words, two zero words behind the data -- and not a byte more
        std::unique_ptr<uint8_t[]> clean(new uint8_t[scan.size() + 16]);
        bool ok = true;
        const size_t nb = unstuff_scan(scan.data(), scan.size(), clean.get(), &ok);
        const size_t padded = (nb + 3 + 8) & ~(size_t)3;
        std::unique_ptr<uint32_t[]> words(new uint32_t[padded / 4]);   // exact size: ASan guards both ends
        memset(words.get(), 0, padded);
        memcpy(words.get(), clean.get(), nb);
        const int nwords = (int)(padded / 4);
        {   // flat loop
            std::unique_ptr<int16_t[]> coef(new int16_t[(size_t)(nblocks + 1) * 64]());
            decode_scan_flat(words.get(), nwords, h.huff, h.td, h.ta, kNaturalOrderHost, n_mcus, coef.get());
        }
        for (int budget : {1024, 37}) {
            std::unique_ptr<int16_t[]> coef(new int16_t[(size_t)(nblocks + 1) * 64]());
            total_rounds += decode_scan_parallel_host(words.get(), nwords, h.huff, h.td, h.ta, kNaturalOrderHost, n_mcus, coef.get(), budget);
            if ((size_t)nwords * 32 <= (size_t)nblocks * 32) {   // the library's rule for the sparse hand-off
                const int cap = nblocks * 32;
                std::unique_ptr<uint32_t[]> list(new uint32_t[(size_t)(nblocks + 1) * 32]);   // the coefficient scratch, reused
                std::unique_ptr<uint32_t[]> bstart(new uint32_t[(size_t)nblocks + 2]);
                for (int i = 0; i < (nblocks + 1) * 32; ++i) list[i] = rng();            // stale contents of an earlier batch
                for (int i = 0; i < nblocks + 2; ++i) bstart[i] = rng();
                decode_scan_parallel_host(words.get(), nwords, h.huff, h.td, h.ta, kNaturalOrderHost, n_mcus,
                                          reinterpret_cast<int16_t*>(list.get()), budget, bstart.get(), cap);
                int16_t blk[64];
                for (int b = 0; b < nblocks; ++b) expand_block(list.get(), bstart.get(), b, cap, blk);
            }
        }
    }
    // damaged HEADERS: the host parser faces file contents; whatever it accepts is decoded with the tables it built
    int accepted = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> pic(jpg);
        const int mode = it % 3;
        const size_t hdr = std::min(pic.size(), h.scan_offset + 16);
        if (mode == 0) { for (int k = 0; k < 1 + (int)(rng() % 6); ++k) pic[rng() % hdr] ^= (uint8_t)(1u << (rng() % 8)); }
        else if (mode == 1) { const size_t a = rng() % hdr, n = 1 + rng() % 24; for (size_t i = a; i < std::min(hdr, a + n); ++i) pic[i] = (uint8_t)rng(); }
        else pic.resize(rng() % (pic.size() + 1));
        std::unique_ptr<uint8_t[]> exact(new uint8_t[pic.size() ? pic.size() : 1]);   // exact size: reads past the end are caught
        memcpy(exact.get(), pic.data(), pic.size());
        JpegHeader hh;
        std::string why;
        if (!parse_jpeg(exact.get(), pic.size(), &hh, &why)) continue;
        ++accepted;
        const FrameGeometry gg = geometry_of(hh);
        if (gg.restart_interval || (long long)gg.mcus_x * gg.mcus_y > 20000) continue;   // (other kernel / absurd size: the library caps pictures too)
        const int mcus = gg.mcus_x * gg.mcus_y, nblk = 6 * mcus;
        std::unique_ptr<uint8_t[]> clean(new uint8_t[hh.scan_len + 16]);
        bool ok = true;
        const size_t nb = unstuff_scan(exact.get() + hh.scan_offset, hh.scan_len, clean.get(), &ok);
        const size_t padded = (nb + 3 + 8) & ~(size_t)3;
        std::unique_ptr<uint32_t[]> words(new uint32_t[padded / 4]);
        memset(words.get(), 0, padded);
        memcpy(words.get(), clean.get(), nb);
        std::unique_ptr<int16_t[]> coef(new int16_t[(size_t)(nblk + 1) * 64]());
        decode_scan_flat(words.get(), (int)(padded / 4), hh.huff, hh.td, hh.ta, kNaturalOrderHost, mcus, coef.get());
        std::unique_ptr<int16_t[]> coef2(new int16_t[(size_t)(nblk + 1) * 64]());
        decode_scan_parallel_host(words.get(), (int)(padded / 4), hh.huff, hh.td, hh.ta, kNaturalOrderHost, mcus, coef2.get(), 64);
    }
    printf("fuzz ok: %d damaged scans, %lld rounds; %d damaged headers, %d accepted by the parser\n", iters, total_rounds, iters, accepted);
    return 0;
}
