// Test shim (CPU): exposes eioku_b200/csrc/ingest_gather.h through a C ABI so tests/test_host_gather.py can compare the
// ingest path's host tap gather with a numpy statement of the layout.  Built by the test with g++; not part of the product.
#include "../eioku_b200/csrc/ingest_gather.h"

extern "C" void shim_gather(int dst_w, int tap_row_bytes, int row_bytes, const int* off, const int32_t* touched, int64_t n_touched,
                            int n_touched_y, int nv12, int prefetch_bytes, int nt_stores, const uint8_t* src, int64_t frame_stride,
                            int64_t pitch, uint8_t* dst, int64_t lo, int64_t hi, int src_height, int streams) {
    esd::GatherSpec g{};
    g.dst_w = dst_w;
    g.tap_row_bytes = tap_row_bytes;
    g.row_bytes = row_bytes;
    g.off = off;
    g.touched = touched;
    g.n_touched = n_touched;
    g.n_touched_y = n_touched_y;
    g.nv12 = nv12 != 0;
    g.i420 = nv12 == 2;  // nv12: 0 = BGR24, 1 = NV12, 2 = planar I420 of src_height rows
    g.src_height = src_height;
    g.streams = streams;
    g.prefetch_bytes_multi = prefetch_bytes > 0 ? 384 : 0;
    g.prefetch_bytes = prefetch_bytes;
    g.nt_stores = nt_stores != 0;
    esd::gather_tap_rows(g, src, frame_stride, pitch, dst, lo, hi);
}
