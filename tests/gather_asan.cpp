// AddressSanitizer harness for the host tap gather (csrc/ingest_gather.h): every source format and loop variant over frames and
// destination buffers of EXACTLY the advertised sizes, so that a read past a row's / frame's end or a write past the ring slot
// shows.  Geometry as the library computes it (OpenCV INTER_LINEAR taps).  Not part of the product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../eioku_b200/csrc/ingest_gather.h"

static void taps(int src, int dst, std::vector<int>& o0) {
    o0.resize(dst);
    const double s = (double)src / dst;
    for (int d = 0; d < dst; ++d) {
        int x = (int)std::floor((d + 0.5) * s - 0.5);
        if (x < 0) x = 0;
        if (x > src - 1) x = src - 1;
        o0[d] = x;
    }
}

int main() {
    struct Case { int w, h, dw, dh; } cases[] = {{1920, 1080, 256, 144}, {642, 362, 256, 144}, {300, 200, 256, 171}, {34, 18, 17, 9}, {2050, 40, 1024, 20}};
    int runs = 0;
    for (const Case& c : cases) {
        std::vector<int> xo, yo;
        taps(c.w, c.dw, xo);
        taps(c.h, c.dh, yo);
        std::vector<char> used(c.h, 0);
        for (int y : yo) { used[y] = 1; used[std::min(y + 1, c.h - 1)] = 1; }
        for (int fmt = 0; fmt < 3; ++fmt) {  // 0 BGR24, 1 NV12, 2 I420
            if (fmt && ((c.w | c.h) & 1)) continue;
            std::vector<int32_t> touched;
            for (int r = 0; r < c.h; ++r) if (used[r]) touched.push_back(r);
            const int n_y = (int)touched.size();
            if (fmt) {
                std::vector<char> cu(c.h / 2, 0);
                for (int r = 0; r < c.h; ++r) if (used[r]) cu[r >> 1] = 1;
                for (int r = 0; r < c.h / 2; ++r) if (cu[r]) touched.push_back(c.h + r);
            }
            std::vector<int> off(c.dw);
            for (int d = 0; d < c.dw; ++d) off[d] = (fmt ? 1 : 3) * xo[d];
            const int rb = fmt ? c.w : c.w * 3;
            const int trb = ((fmt ? 4 : 6) * c.dw + 15) & ~15;
            const int n = 3;
            const size_t pitch = (size_t)rb;   // dense rows: the tightest case for an over-read
            const size_t frame = fmt ? pitch * c.h * 3 / 2 : pitch * c.h;
            std::unique_ptr<uint8_t[]> src(new uint8_t[frame * n]);
            for (size_t i = 0; i < frame * n; ++i) src[i] = (uint8_t)(i * 131u);
            const int64_t items = (int64_t)n * (int64_t)touched.size();
            uint8_t* dst = nullptr;   // 16-byte aligned (non-temporal stores), exactly items * trb bytes
            if (posix_memalign((void**)&dst, 64, (size_t)items * trb) != 0) return 2;
            for (int streams : {1, 8}) {
                for (int nt = 0; nt < 2; ++nt) {
                    esd::GatherSpec g{};
                    g.dst_w = c.dw; g.tap_row_bytes = trb; g.row_bytes = rb; g.off = off.data(); g.touched = touched.data();
                    g.n_touched = (int64_t)touched.size(); g.n_touched_y = n_y; g.nv12 = fmt != 0; g.i420 = fmt == 2; g.src_height = c.h;
                    g.prefetch_bytes = 4096 < rb ? 4096 : rb; g.nt_stores = nt != 0; g.streams = streams; g.prefetch_bytes_multi = 384; g.prefetch_hint = 0;
                    esd::gather_tap_rows(g, src.get(), (int64_t)frame, (int64_t)pitch, dst, 0, 5);
                    esd::gather_tap_rows(g, src.get(), (int64_t)frame, (int64_t)pitch, dst, 5, items);
                    ++runs;
                }
            }
            free(dst);
        }
    }
    printf("gather asan ok: %d runs\n", runs);
    return 0;
}
