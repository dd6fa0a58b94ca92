// host_tables.h -- CUDA-free host logic of libesd: OpenCV's resize coefficient tables and the fused kernel's work plan.
// Unit-tested on the CPU (tests/test_host_tables.py through tests/tables_shim.cpp) against the oracle's tables.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace esd {

// Work unit of the fused kernel: frames [f0, f1) (batch-relative) of destination-row group rg.
struct Unit {
    int rg, f0, f1, pad;
};

// OpenCV resize.cpp INTER_LINEAR coefficient setup: float32 fraction from a double
// scale, cvRound (half-even) to 11-bit fixed point.
inline void axis_tables(int src, int dst, std::vector<int>& o0, std::vector<int>& o1, std::vector<int>& c0,
                        std::vector<int>& c1) {
    o0.resize(dst); o1.resize(dst); c0.resize(dst); c1.resize(dst);
    const double inv_scale = (double)dst / (double)src;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        volatile double pos = (d + 0.5) * scale;  // volatile: no FMA contraction / excess precision
        pos = pos - 0.5;
        float f = (float)pos;
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        o0[d] = s;
        o1[d] = std::min(s + 1, src - 1);
        volatile float w0 = (1.f - f) * 2048.f, w1 = f * 2048.f;
        c0[d] = (int)lrintf(w0);
        c1[d] = (int)lrintf(w1);
    }
}

// OpenCV resize.cpp computeResizeAreaTab for one axis (cn = 1), grouped per destination index: the entries of
// destination d are [begin[d], begin[d + 1]) of (src, wt), in the order the C++ loop visits them.
inline void area_axis_table(int ssize, int dsize, std::vector<int>& begin, std::vector<int>& src, std::vector<float>& wt) {
    const double scale = (double)ssize / dsize;
    begin.assign(1, 0);
    src.clear();
    wt.clear();
    for (int d = 0; d < dsize; ++d) {
        volatile double fsx1 = d * scale;
        volatile double fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) { src.push_back(sx1 - 1); wt.push_back((float)((sx1 - fsx1) / cell)); }
        for (int x = sx1; x < sx2; ++x) { src.push_back(x); wt.push_back((float)(1.0 / cell)); }
        if (fsx2 - sx2 > 1e-3) { src.push_back(sx2); wt.push_back((float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)); }
        begin.push_back((int)src.size());
    }
}

// Shared-memory wavefronts one staged-row read costs the eight consumer warps when lanes are `ks` destination columns apart
// (column of (warp, lane) = ks * lane + warp % ks + 32 * ks * (warp / ks)): every thread reads `n_words` consecutive 32-bit
// words starting at the word that holds byte `byte_off[column]`; a warp-wide load takes as many passes as the busiest of
// the 32 banks has DISTINCT words.  Columns >= dst_w (threads without a pixel) do not load.
inline int64_t tap_load_wavefronts(const std::vector<uint32_t>& byte_off, int dst_w, int n_words, int ks, int warps = 8) {
    int64_t total = 0;
    for (int w = 0; w < warps; ++w)
        for (int j = 0; j < n_words; ++j) {
            std::vector<uint32_t> seen[32];
            int worst = 0;
            for (int l = 0; l < 32; ++l) {
                const int col = ks * l + (w & (ks - 1)) + 32 * ks * (w / ks);
                if (col >= dst_w) continue;
                const uint32_t word = (byte_off[col] >> 2) + (uint32_t)j;
                auto& v = seen[word & 31u];
                if (std::find(v.begin(), v.end(), word) == v.end()) v.push_back(word);
                worst = std::max(worst, (int)v.size());
            }
            total += worst;
        }
    return total;
}

// The lane stride (1, 2, 4, 8) with the fewest wavefronts; ties go to the smaller stride (better-coalesced side outputs).
inline int choose_lane_stride(const std::vector<uint32_t>& byte_off, int dst_w, int n_words) {
    int best = 1;
    int64_t best_cost = tap_load_wavefronts(byte_off, dst_w, n_words, 1);
    for (int ks = 2; ks <= 8; ks *= 2) {
        const int64_t c = tap_load_wavefronts(byte_off, dst_w, n_words, ks);
        if (c < best_cost) { best_cost = c; best = ks; }
    }
    return best;
}

inline int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

// Work decomposition of one fused launch over n_groups row groups x n frames.  A unit is (row group, frame range); units
// that do not start at frame 0 re-read one halo frame to rebuild the previous HSV, so fewer / longer units are cheaper.
// mode 1 (strips): the (group-major, frame-minor) item sequence cut into `max_ctas` equal contiguous pieces;
// mode 2 (chunks): frame chunks shared by all row groups, dealt round-robin.  Returns the grid size; CTA g works on units
// [begin[g], begin[g + 1]).
inline int build_unit_plan(int n_groups, int64_t n, int64_t max_ctas, int mode, std::vector<Unit>& units, std::vector<int>& begin) {
    const int G = n_groups;
    const int64_t total = (int64_t)G * n;  // (group, frame) items
    int grid = (int)std::min<int64_t>(max_ctas, total);
    units.clear();
    begin.assign(grid + 1, 0);
    if (mode != 2) {
        for (int g = 0; g < grid; ++g) {
            int64_t lo = total * g / grid, hi = total * (g + 1) / grid;
            begin[g] = (int)units.size();
            while (lo < hi) {
                const int rg = (int)(lo / n);
                const int64_t f0 = lo % n;
                const int64_t f1 = std::min<int64_t>(n, f0 + (hi - lo));
                units.push_back(Unit{rg, (int)f0, (int)f1, 0});
                lo += f1 - f0;
            }
        }
        begin[grid] = (int)units.size();
    } else {
        int64_t n_chunks = std::max<int64_t>(1, ((int64_t)grid * 4 + G - 1) / G);
        n_chunks = std::min<int64_t>(n_chunks, std::max<int64_t>(1, n / 16));
        // make the unit count a multiple of the grid when possible
        const int64_t per = grid / gcd64(grid, G);
        if (n_chunks >= per) n_chunks = n_chunks / per * per;
        std::vector<Unit> all;
        for (int64_t ch = 0; ch < n_chunks; ++ch)
            for (int rg = 0; rg < G; ++rg)
                all.push_back(Unit{rg, (int)(n * ch / n_chunks), (int)(n * (ch + 1) / n_chunks), 0});
        grid = (int)std::min<int64_t>(grid, (int64_t)all.size());
        begin.assign(grid + 1, 0);
        for (int g = 0; g < grid; ++g) {
            begin[g] = (int)units.size();
            for (size_t u = g; u < all.size(); u += grid) units.push_back(all[u]);
        }
        begin[grid] = (int)units.size();
    }
    return grid;
}

}  // namespace esd
