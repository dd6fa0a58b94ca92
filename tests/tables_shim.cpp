// Test shim (CPU): exposes eioku_b200/csrc/host_tables.h through a C ABI for tests/test_host_tables.py.  Not part of the product.
#include "../eioku_b200/csrc/host_tables.h"

#include <string.h>

extern "C" {
void shim_axis_tables(int src, int dst, int* o0, int* o1, int* c0, int* c1) {
    std::vector<int> a, b, c, d;
    esd::axis_tables(src, dst, a, b, c, d);
    memcpy(o0, a.data(), sizeof(int) * dst);
    memcpy(o1, b.data(), sizeof(int) * dst);
    memcpy(c0, c.data(), sizeof(int) * dst);
    memcpy(c1, d.data(), sizeof(int) * dst);
}
// returns the number of entries; begin has dsize + 1 ints, src / wt room for cap entries
int shim_area_table(int ssize, int dsize, int* begin, int* src, float* wt, int cap) {
    std::vector<int> b, s;
    std::vector<float> w;
    esd::area_axis_table(ssize, dsize, b, s, w);
    if ((int)s.size() > cap) return -1;
    memcpy(begin, b.data(), sizeof(int) * b.size());
    memcpy(src, s.data(), sizeof(int) * s.size());
    memcpy(wt, w.data(), sizeof(float) * w.size());
    return (int)s.size();
}
// returns the grid; units as int[n_units][4], begin as int[grid + 1]; *n_units receives the unit count
int shim_unit_plan(int n_groups, long long n, long long max_ctas, int mode, int* units, int units_cap, int* begin, int begin_cap, int* n_units) {
    std::vector<esd::Unit> u;
    std::vector<int> b;
    const int grid = esd::build_unit_plan(n_groups, n, max_ctas, mode, u, b);
    *n_units = (int)u.size();
    if ((int)u.size() > units_cap || (int)b.size() > begin_cap) return -1;
    memcpy(units, u.data(), sizeof(esd::Unit) * u.size());
    memcpy(begin, b.data(), sizeof(int) * b.size());
    return grid;
}
// wavefronts of one staged-row read for lane stride ks; *best receives choose_lane_stride's pick
long long shim_tap_wavefronts(const unsigned* byte_off, int dst_w, int n_words, int ks, int* best) {
    std::vector<uint32_t> off(byte_off, byte_off + dst_w);
    if (best) *best = esd::choose_lane_stride(off, dst_w, n_words);
    return esd::tap_load_wavefronts(off, dst_w, n_words, ks);
}
}
