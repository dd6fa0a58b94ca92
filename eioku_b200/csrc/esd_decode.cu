// esd_decode.cu -- libesd_decode.so: Motion-JPEG in AVI -> device-resident BGR24 frames through nvJPEG (include/esd_decode.h).
//
// Replaces the host decode loop of the reference (cv2.VideoCapture.read(), ml-service/src/services/model_manager.py:237-263;
// the ffmpeg child of the scene task, :736-755): the host walks the RIFF container and hands compressed pictures to the GPU.
// No CPU decode fallback: without a CUDA device / nvJPEG every entry point fails.
#include "../../include/esd_decode.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <nvjpeg.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "guard_alloc.h"
#include "jpeg_parse.h"

namespace {
thread_local std::string g_open_error;

struct Picture {
    uint64_t offset;
    uint32_t size;
};
}  // namespace

struct esd_mjpeg {
    std::string err;
    int device = 0;
    // container
    int fd = -1;
    const uint8_t* map = nullptr;
    size_t map_bytes = 0;
    std::vector<Picture> pics;
    int width = 0, height = 0, fps_num = 30, fps_den = 1;
    int64_t compressed_bytes = 0;
    int64_t pos = 0;
    // decoder
    nvjpegHandle_t nj = nullptr;
    nvjpegJpegState_t state = nullptr;
    int backend = ESD_JPEG_DEFAULT;
    unsigned hw_engines = 0;
    int batch = 0;
    int initialized_batch = 0;  // batch size of the last nvjpegDecodeBatchedInitialize
    // double-buffered output + pinned bitstream staging
    uint8_t* d_out[2] = {nullptr, nullptr};
    uint8_t* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_bytes[2] = {0, 0};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool in_flight[2] = {false, false};
    // ESD_DEC_TIMING=1: CUDA events around the stages of every batch, summarised on stderr when the handle closes
    bool timing = getenv("ESD_DEC_TIMING") != nullptr;
    cudaEvent_t tev[2][4] = {};
    bool tev_armed[2] = {false, false};
    double t_sum[4] = {0, 0, 0, 0};   // ms: staging copy + clear, entropy, idct + colour, host staging
    int64_t t_batches = 0, t_pictures = 0;
    int timeline = getenv("ESD_DEC_TIMING") ? atoi(getenv("ESD_DEC_TIMING")) : 0;   // 2: per-batch timeline against a process-wide origin
    double host_t[2][3] = {};         // ms since the origin: read entered, staging begins (slot free), staging ends
    int64_t reads = 0;
    std::vector<const unsigned char*> ptrs;
    std::vector<size_t> lens;
    std::vector<nvjpegImage_t> imgs;
    // native decoder (ESD_JPEG_NATIVE): this library's own kernels over csrc/jpeg_core.h
    esdjpeg::FrameGeometry geo{};
    int tq[3] = {0, 0, 0}, td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int blocks_per_frame = 0;
    // ESD_DEC_ENTROPY=flat keeps the thread-per-picture loop for pictures without restart markers (A/B switch); the default decodes
    // each picture with a block of threads
    bool parallel_entropy = !(getenv("ESD_DEC_ENTROPY") && strcmp(getenv("ESD_DEC_ENTROPY"), "flat") == 0);
    int stage_threads = getenv("ESD_DEC_STAGE_THREADS") ? std::max(1, atoi(getenv("ESD_DEC_STAGE_THREADS"))) : 4;   // host threads staging a batch
    bool flat = false;                           // no restart interval: the host removes the byte stuffing and the flat scan decoder runs
    size_t plane_bytes = 0;                      // Y + Cb + Cr sample planes of one frame (MCU-padded)
    // Two decode lanes: batch k runs on lane k & 1 (a stream of the library's own, with its own staging mirror, coefficient and
    // plane scratch), so the entropy stage of batch k + 1 overlaps batch k's instead of queueing behind it on the caller's stream
    // -- a session keeps 2 x batch pictures in flight, which is what the decoder's throughput is made of.  The caller's stream
    // waits for the lane's `done` event; the lane waits for `consumed`, recorded on the caller's stream at the next read (when the
    // work that reads this slot's previous contents has been enqueued).  ESD_DEC_LANES=1 restores the single-stream order.
    int lanes = getenv("ESD_DEC_LANES") && atoi(getenv("ESD_DEC_LANES")) == 1 ? 1 : 2;
    cudaStream_t lane[2] = {nullptr, nullptr};
    cudaEvent_t consumed[2] = {nullptr, nullptr};
    bool have_consumed[2] = {false, false};
    uint8_t* d_comp[2] = {nullptr, nullptr};     // compressed pictures of the batch
    size_t d_comp_bytes[2] = {0, 0};
    struct NativeDesc { uint32_t off, len; uint32_t dht[4]; uint16_t nvals[4]; };  // dht / nvals: DC0, DC1, AC0, AC1 (offset of the 16 counts in the staging block; 0 = table absent)
    NativeDesc* d_desc = nullptr;                // [batch]
    uint16_t* d_quant = nullptr;                 // [batch][3][64] natural order, per component
    int16_t* d_coef[2] = {nullptr, nullptr};     // [batch][blocks_per_frame][64], per lane
    uint32_t* d_bstart[2] = {nullptr, nullptr};  // [batch][blocks_per_frame + 1]: sparse hand-off, where each block's entries begin
    bool sparse_handoff = !(getenv("ESD_DEC_SPARSE") && atoi(getenv("ESD_DEC_SPARSE")) == 0);   // A/B switch
    bool coef_dirty[2] = {true, true};           // needs a full clear before the next batch (first use, or a batch that failed half-way)
    uint8_t* d_planes[2] = {nullptr, nullptr};   // [batch][plane_bytes], per lane
    std::vector<uint8_t> h_meta[2];              // pinned-free host staging of descriptors + quant tables (copied with the batch)
};

namespace {

int fail(esd_mjpeg* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_open_error = buf;
    return code;
}

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
bool tag(const uint8_t* p, const char* t) { return memcmp(p, t, 4) == 0; }

// RIFF walk: collects the video chunks ("##dc" / "##db" of stream 0) of every 'movi' list (also inside 'rec ' lists and in the
// extra 'RIFF AVIX' segments OpenDML files have) and reads the first video stream's header.
struct AviWalk {
    esd_mjpeg* h;
    bool have_vids = false, is_mjpeg = false;
    int cur_stream = -1, video_stream = -1;
    char fourcc[5] = {0, 0, 0, 0, 0};

    void list(uint64_t off, uint64_t end, bool in_movi) {
        const uint8_t* b = h->map;
        while (off + 8 <= end) {
            const uint8_t* c = b + off;
            const uint64_t sz = rd32(c + 4);
            const uint64_t body = off + 8;
            if (body + sz > end + 1 && !(tag(c, "RIFF") || tag(c, "LIST"))) break;  // truncated file: stop at the last whole chunk
            if (tag(c, "RIFF") || tag(c, "LIST")) {
                if (body + 4 > end) break;
                const uint8_t* typ = b + body;
                const uint64_t lend = std::min<uint64_t>(end, body + sz);
                if (tag(typ, "strl")) ++cur_stream;
                list(body + 4, lend, in_movi || tag(typ, "movi"));
            } else if (tag(c, "strh") && sz >= 32) {
                if (tag(b + body, "vids") && !have_vids) {
                    have_vids = true;
                    video_stream = cur_stream;
                    memcpy(fourcc, b + body + 4, 4);
                    const uint32_t scale = rd32(b + body + 20), rate = rd32(b + body + 24);
                    if (scale && rate) { h->fps_num = (int)rate; h->fps_den = (int)scale; }
                }
            } else if (tag(c, "strf") && sz >= 20 && have_vids && cur_stream == video_stream && h->width == 0) {
                h->width = (int)rd32(b + body + 4);
                const int32_t hh = (int32_t)rd32(b + body + 8);
                h->height = hh < 0 ? -hh : hh;
                char comp[5] = {0, 0, 0, 0, 0};
                memcpy(comp, b + body + 16, 4);
                auto mj = [](const char* f) {
                    return !strncasecmp(f, "MJPG", 4) || !strncasecmp(f, "AVI1", 4) || !strncasecmp(f, "JPEG", 4) || !strncasecmp(f, "IJPG", 4);
                };
                is_mjpeg = mj(comp) || mj(fourcc);
                if (!is_mjpeg) memcpy(fourcc, comp, 4);
            } else if (in_movi && sz > 0 && (c[2] == 'd') && (c[3] == 'c' || c[3] == 'b') && c[0] >= '0' && c[0] <= '9' && c[1] >= '0' && c[1] <= '9') {
                const int stream = (c[0] - '0') * 10 + (c[1] - '0');
                if (stream == std::max(0, video_stream)) {
                    h->pics.push_back(Picture{body, (uint32_t)sz});
                    h->compressed_bytes += (int64_t)sz;
                }
            }
            off = body + sz + (sz & 1);
        }
    }
};

nvjpegBackend_t nj_backend(int b) {
    return b == ESD_JPEG_HARDWARE ? NVJPEG_BACKEND_HARDWARE : b == ESD_JPEG_GPU_HYBRID ? NVJPEG_BACKEND_GPU_HYBRID : NVJPEG_BACKEND_DEFAULT;
}

}  // namespace


// =========================================================================== native baseline-JPEG decoder (ESD_JPEG_NATIVE)
// Three kernels per batch, all arithmetic in csrc/jpeg_core.h (checked bit-exact against cv2.imdecode on the CPU):
//   jpeg_entropy_kernel  one thread per PICTURE walks its Huffman-coded scan (the only sequential part of JPEG; pictures
//                        written by ffmpeg / OpenCV carry no restart markers, so the parallelism is across the pictures of the
//                        batch and across decoder sessions) and scatters the non-zero coefficients into a zeroed buffer;
//   jpeg_idct_kernel     one thread per 8x8 block: dequantise + libjpeg's ISLOW IDCT -> MCU-padded Y / Cb / Cr planes;
//   jpeg_color_kernel    one thread per four pixels of two rows: fancy h2v2 chroma upsampling + JFIF YCbCr -> BGR24, written in the
//                        dense layout esd_push_frames reads.
namespace {
__constant__ uint8_t c_natural_order[64] = ESD_JPEG_NATURAL_ORDER;

struct NativeLayout {
    int n;                       // pictures in the batch
    int mcus_x, mcus_y, restart_interval;
    int width, height;
    int blocks_per_frame;        // 6 * mcus_x * mcus_y
    int td[3], ta[3];
    unsigned long long plane_bytes;
};

// ffmpeg's encoder optimises the Huffman tables PER PICTURE, so every thread builds the decoding tables of its own picture from
// the picture's DHT segments -- into its own 5.7 KB slice of shared memory (32 pictures = 178 KB per block, one block per SM):
// every symbol costs a dependent table look-up, and from global memory that latency would dominate the kernel.
constexpr int kEntropyThreads = 32;
__global__ void __launch_bounds__(kEntropyThreads) jpeg_entropy_kernel(NativeLayout L, const uint8_t* __restrict__ stage,
                                                                       const esd_mjpeg::NativeDesc* __restrict__ desc, int16_t* __restrict__ coef) {
    extern __shared__ __align__(16) uint8_t ent_smem[];
    __shared__ uint8_t s_nat[64];
    esdjpeg::ScanTables* tabs = reinterpret_cast<esdjpeg::ScanTables*>(ent_smem);
    for (int i = threadIdx.x; i < 64; i += kEntropyThreads) s_nat[i] = c_natural_order[i];
    __syncthreads();
    const int f = blockIdx.x * kEntropyThreads + threadIdx.x;
    if (f >= L.n) return;
    const esd_mjpeg::NativeDesc d = desc[f];
    esdjpeg::ScanTables& T = tabs[threadIdx.x];
    bool ok = true;
    for (int t = 0; t < 4; ++t) {
        esdjpeg::HuffTable* tb = t < 2 ? &T.dc[t] : &T.ac[t - 2];
        if (d.dht[t]) ok = esdjpeg::build_huff_table(stage + d.dht[t], stage + d.dht[t] + 16, (int)d.nvals[t], tb) && ok;
    }
    if (!ok) return;  // the host validated the tables; a picture that still fails decodes to zero coefficients (mid-grey)
    esdjpeg::BitReader br;
    br.init(stage + d.off, (int)d.len);
    int pred[3] = {0, 0, 0};
    int16_t* cf = coef + (size_t)f * L.blocks_per_frame * 64;  // decoding order: six blocks per MCU
    const esdjpeg::HuffTable& ydc = T.dc[L.td[0]];
    const esdjpeg::HuffTable& yac = T.ac[L.ta[0]];
    const esdjpeg::HuffTable& bdc = T.dc[L.td[1]];
    const esdjpeg::HuffTable& bac = T.ac[L.ta[1]];
    const esdjpeg::HuffTable& rdc = T.dc[L.td[2]];
    const esdjpeg::HuffTable& rac = T.ac[L.ta[2]];
    int mcu = 0;
    for (int my = 0; my < L.mcus_y; ++my)
        for (int mx = 0; mx < L.mcus_x; ++mx, ++mcu) {
            if (L.restart_interval && mcu && mcu % L.restart_interval == 0) { br.restart(); pred[0] = pred[1] = pred[2] = 0; }
            int16_t* m = cf + (size_t)mcu * 6 * 64;
#pragma unroll
            for (int b = 0; b < 4; ++b) esdjpeg::decode_block(br, ydc, yac, s_nat, pred[0], m + b * 64);
            esdjpeg::decode_block(br, bdc, bac, s_nat, pred[1], m + 4 * 64);
            esdjpeg::decode_block(br, rdc, rac, s_nat, pred[2], m + 5 * 64);
        }
}

// Fast path (no restart markers, byte stuffing removed by the host while it stages the picture): the flat, branch-uniform scan
// decoder of jpeg_core.h -- a warp walks 32 pictures in lock-step, one Huffman symbol per iteration.
__global__ void __launch_bounds__(kEntropyThreads) jpeg_entropy_flat_kernel(NativeLayout L, const uint8_t* __restrict__ stage,
                                                                            const esd_mjpeg::NativeDesc* __restrict__ desc,
                                                                            int16_t* __restrict__ coef) {
    extern __shared__ __align__(16) uint8_t ent_smem[];
    __shared__ uint8_t s_nat[64];
    esdjpeg::ScanTables* tabs = reinterpret_cast<esdjpeg::ScanTables*>(ent_smem);
    for (int i = threadIdx.x; i < 64; i += kEntropyThreads) s_nat[i] = c_natural_order[i];
    __syncthreads();
    const int f = blockIdx.x * kEntropyThreads + threadIdx.x;
    if (f >= L.n) return;
    const esd_mjpeg::NativeDesc d = desc[f];
    esdjpeg::ScanTables& T = tabs[threadIdx.x];
    bool ok = true;
    for (int t = 0; t < 4; ++t) {
        esdjpeg::HuffTable* tb = t < 2 ? &T.dc[t] : &T.ac[t - 2];
        if (d.dht[t]) ok = esdjpeg::build_huff_table(stage + d.dht[t], stage + d.dht[t] + 16, (int)d.nvals[t], tb) && ok;
    }
    if (!ok) return;
    const int td[3] = {L.td[0], L.td[1], L.td[2]}, ta[3] = {L.ta[0], L.ta[1], L.ta[2]};
    esdjpeg::decode_scan_flat(reinterpret_cast<const uint32_t*>(stage + d.off), (int)d.len, T, td, ta, s_nat, L.mcus_x * L.mcus_y,
                              coef + (size_t)f * L.blocks_per_frame * 64);
}

// Many threads per picture (jpeg_core.h: decode_span -- speculative sub-sequences that self-synchronise): one block per picture,
// a thread per sub-sequence of the scan.  Rounds until no thread's start state changes (2-8 for camera / encoder content, at worst
// one per sub-sequence, which is the sequential decoding at the old cost), block-wide prefix sums of the block counts and DC sums,
// one writing pass.  A 1080p picture: ~1 000 threads x ~100 symbols per round instead of one thread x 275 000 symbols.
constexpr int kParThreads = 1024;
template <bool SPARSE>
__global__ void __launch_bounds__(kParThreads) jpeg_entropy_parallel_kernel(NativeLayout L, const uint8_t* __restrict__ stage,
                                                                            const esd_mjpeg::NativeDesc* __restrict__ desc,
                                                                            int16_t* __restrict__ coef, uint32_t* __restrict__ bstart) {
    __shared__ esdjpeg::ScanTables T;
    __shared__ uint8_t s_nat[64];
    __shared__ esdjpeg::SpanState s_end[kParThreads];
    __shared__ int s_warp[5][kParThreads / 32];
    __shared__ int s_bad;
    const int tid = threadIdx.x, f = blockIdx.x;
    const esd_mjpeg::NativeDesc d = desc[f];
    if (tid == 0) s_bad = 0;
    if (tid < 64) s_nat[tid] = c_natural_order[tid];
    __syncthreads();
    if (tid < 4) {   // DC0, DC1, AC0, AC1: the picture's own tables, one thread each
        esdjpeg::HuffTable* tb = tid < 2 ? &T.dc[tid] : &T.ac[tid - 2];
        if (d.dht[tid] && !esdjpeg::build_huff_table(stage + d.dht[tid], stage + d.dht[tid] + 16, (int)d.nvals[tid], tb)) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) return;
    const int td[3] = {L.td[0], L.td[1], L.td[2]}, ta[3] = {L.ta[0], L.ta[1], L.ta[2]};
    const uint32_t* words = reinterpret_cast<const uint32_t*>(stage + d.off);
    const int nwords = (int)d.len;
    const int total_blocks = 6 * L.mcus_x * L.mcus_y;
    const uint32_t S = esdjpeg::span_bits_for(nwords, total_blocks, kParThreads), nbits = (uint32_t)nwords * 32u;
    const int nt = (int)((nbits + S - 1) / S);
    const bool active = tid < nt;
    const unsigned long long e64 = (unsigned long long)(tid + 1) * S;
    const uint32_t limit = (uint32_t)(e64 < nbits ? e64 : nbits);
    esdjpeg::SpanState in{(uint32_t)tid * S, 0, 0};
    esdjpeg::SpanResult res;
    res.end = in; res.n_blocks = 0; res.dc[0] = res.dc[1] = res.dc[2] = 0;
    if (active) res = esdjpeg::decode_span<esdjpeg::SPAN_COUNT>(words, nwords, T, td, ta, s_nat, in, limit, nullptr, 0, nullptr, total_blocks);
    for (;;) {
        s_end[tid] = res.end;
        __syncthreads();
        bool changed = false;
        if (active && tid > 0) {
            const esdjpeg::SpanState nin = s_end[tid - 1];
            if (!esdjpeg::same_state(nin, in)) { in = nin; changed = true; }
        }
        if (!__syncthreads_or(changed ? 1 : 0)) break;   // also: every read of s_end is done before the next round's writes
        if (changed) res = esdjpeg::decode_span<esdjpeg::SPAN_COUNT>(words, nwords, T, td, ta, s_nat, in, limit, nullptr, 0, nullptr, total_blocks);
    }
    // exclusive prefix sums over the threads: blocks completed and DC sums per component before this sub-sequence
    int v[5] = {active ? res.n_blocks : 0, active ? res.dc[0] : 0, active ? res.dc[1] : 0, active ? res.dc[2] : 0, active ? res.n_coefs : 0};
    int incl[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        int x = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((tid & 31) >= o) x += y;
        }
        incl[q] = x;
        if ((tid & 31) == 31) s_warp[q][tid >> 5] = x;
    }
    __syncthreads();
    int base[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < (tid >> 5); ++w) {
#pragma unroll
        for (int q = 0; q < 5; ++q) base[q] += s_warp[q][w];
    }
    if (!active) return;
    const int blk0 = base[0] + incl[0] - v[0];
    const int pred[3] = {base[1] + incl[1] - v[1], base[2] + incl[2] - v[2], base[3] + incl[3] - v[3]};
    int16_t* mine = coef + (size_t)f * L.blocks_per_frame * 64;
    if (SPARSE)   // entry list in the picture's slice of the coefficient scratch (32 entries per block + the spare block as the sink)
        esdjpeg::decode_span<esdjpeg::SPAN_SPARSE>(words, nwords, T, td, ta, s_nat, in, limit, mine, blk0, pred, total_blocks,
                                                   bstart + (size_t)f * (L.blocks_per_frame + 1), base[4] + incl[4] - v[4], total_blocks * 32);
    else
        esdjpeg::decode_span<esdjpeg::SPAN_DENSE>(words, nwords, T, td, ta, s_nat, in, limit, mine, blk0, pred, total_blocks);
}

// IDCT from the sparse hand-off: a thread expands its block's entries into a column of a shared-memory tile ([64 positions][128
// threads] int16: a warp's accesses to one position fall into 16 distinct words, conflict-free whatever the positions), then
// runs the same ISLOW arithmetic.  Blocks with at most their DC entry take the short cut.
__global__ void __launch_bounds__(128) jpeg_idct_sparse_kernel(NativeLayout L, const int16_t* __restrict__ coef, const uint32_t* __restrict__ bstart,
                                                               const uint16_t* __restrict__ quant, uint8_t* __restrict__ planes) {
    __shared__ int16_t tile[64][128];
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    const int total_blocks = L.blocks_per_frame - 1;
    const bool live = blk < total_blocks;
    int comp = 0, bx = 0, by = 0;
    if (live) esdjpeg::block_position(blk, L.mcus_x, &comp, &bx, &by);
    const int bw = comp == 0 ? 2 * L.mcus_x : L.mcus_x;
    const int stride = bw * 8;
    const size_t ysz = (size_t)(2 * L.mcus_x * 8) * (L.mcus_y * 16), csz = (size_t)(L.mcus_x * 8) * (L.mcus_y * 8);
    uint8_t* plane = planes + (size_t)f * L.plane_bytes + (comp == 0 ? 0 : (comp == 1 ? ysz : ysz + csz));
    uint8_t* out = plane + (size_t)(by * 8) * stride + bx * 8;
    const uint32_t* list = reinterpret_cast<const uint32_t*>(coef + (size_t)f * L.blocks_per_frame * 64);
    const uint32_t* bs = bstart + (size_t)f * (L.blocks_per_frame + 1);
    const uint32_t cap = (uint32_t)total_blocks * 32u;
    uint32_t st = 0, en = 0;
    if (live) {
        st = bs[blk]; en = bs[blk + 1];
        if (en > cap) en = cap;
        if (st > en) st = en;
        if (en - st > 64u) en = st + 64u;
    }
    const uint16_t* q = quant + ((size_t)f * 3 + comp) * 64;
    if (live && en - st <= 1u) {
        // DC only (or an empty block): both passes of the ISLOW IDCT reduce to (4 * dc * q0 + 16) >> 5 for every sample
        const int dc = en > st ? (int)(int16_t)(list[st] & 0xffffu) : 0;
        const int32_t v = esdjpeg::descale((dc * (int32_t)q[0]) << 2, 5);
        const uint32_t w4 = (uint32_t)esdjpeg::range_limit(v) * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(out + (size_t)r * stride) = make_uint2(w4, w4);
    }
    const bool full = live && en - st > 1u;
    if (!__syncthreads_or(full ? 1 : 0)) return;   // nobody in this CTA needs the tile
    if (full) {
#pragma unroll
        for (int i = 0; i < 64; ++i) tile[i][threadIdx.x] = 0;
        for (uint32_t e = st; e < en; ++e) {
            const uint32_t w = list[e];
            tile[(w >> 16) & 63u][threadIdx.x] = (int16_t)(w & 0xffffu);
        }
        int16_t c[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) c[i] = tile[i][threadIdx.x];
        uint16_t ql[64];
#pragma unroll
        for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(ql)[i] = reinterpret_cast<const uint4*>(q)[i];
        union { uint2 v[8]; uint8_t b[64]; } o;
        esdjpeg::idct_islow(c, ql, o.b, 8);
#pragma unroll
        for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(out + (size_t)r * stride) = o.v[r];
    }
}

// The kernel also RESTORES THE ZEROS: the entropy stage writes only non-zero coefficients into a buffer that must be clear, and
// clearing 6.3 MB per 1080p picture ahead of every batch was a fifth of the batch's time.  Each thread zeroes the 16-byte pieces of
// its block that held anything (typically 1-3 of 8), so the buffer is clean again when the kernel ends and only ~a third of it is
// ever written back.
__global__ void __launch_bounds__(128) jpeg_idct_kernel(NativeLayout L, int16_t* __restrict__ coef, const uint16_t* __restrict__ quant,
                                                        uint8_t* __restrict__ planes) {
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (blk >= L.blocks_per_frame - 1) return;  // the last block of a picture's buffer is the entropy stage's scratch
    int comp, bx, by;
    esdjpeg::block_position(blk, L.mcus_x, &comp, &bx, &by);  // coefficients are stored in decoding order
    const int bw = comp == 0 ? 2 * L.mcus_x : L.mcus_x;       // blocks per row of this component
    const int stride = bw * 8;
    const size_t ysz = (size_t)(2 * L.mcus_x * 8) * (L.mcus_y * 16), csz = (size_t)(L.mcus_x * 8) * (L.mcus_y * 8);
    uint8_t* plane = planes + (size_t)f * L.plane_bytes + (comp == 0 ? 0 : (comp == 1 ? ysz : ysz + csz));
    uint8_t* out = plane + (size_t)(by * 8) * stride + bx * 8;
    uint4* src = reinterpret_cast<uint4*>(coef + ((size_t)f * L.blocks_per_frame + blk) * 64);
    union { uint4 v[8]; int16_t c[64]; } u;
    uint32_t ac = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u.v[i] = src[i];
        ac |= u.v[i].y | u.v[i].z | u.v[i].w | (i ? u.v[i].x : (u.v[i].x & 0xffff0000u));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (u.v[i].x | u.v[i].y | u.v[i].z | u.v[i].w) src[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint16_t* q = quant + ((size_t)f * 3 + comp) * 64;
    if (ac == 0) {
        // DC only: both passes of the ISLOW IDCT reduce to (4 * dc * q0 + 16) >> 5 for every sample (same rounding)
        const int32_t v = esdjpeg::descale(((int32_t)u.c[0] * (int32_t)q[0]) << 2, 5);
        const uint32_t s = esdjpeg::range_limit(v);
        const uint32_t w4 = s * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(out + (size_t)r * stride) = make_uint2(w4, w4);
        return;
    }
    uint16_t ql[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 t = reinterpret_cast<const uint4*>(q)[i];
        reinterpret_cast<uint4*>(ql)[i] = t;
    }
    union { uint2 v[8]; uint8_t b[64]; } o;
    esdjpeg::idct_islow(u.c, ql, o.b, 8);
#pragma unroll
    for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(out + (size_t)r * stride) = o.v[r];
}

// Four pixels of TWO picture rows (2r, 2r + 1) per thread: the rows share their chroma row cy = r, so the column sums of the
// fancy h2v2 upsampling (3 * C[cy][j] + C[neighbour row][j], jdsample.c) are formed once per chroma column -- four columns
// (j = i-1 .. i+2) serve the eight pixels; the first version recomputed two of them for every pixel and channel (1.66 ms of the
// 3.6 ms a 256-picture batch spent behind the entropy stage).  Arithmetic and results are fancy_chroma's / ycc_to_bgr's.
__global__ void __launch_bounds__(256) jpeg_color_kernel(NativeLayout L, const uint8_t* __restrict__ planes, uint8_t* __restrict__ bgr) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int r = blockIdx.y, f = blockIdx.z;
    const int y0 = 2 * r;
    if (x4 >= L.width || y0 >= L.height) return;
    const int yw = 2 * L.mcus_x * 8, cw = L.mcus_x * 8;   // padded strides
    const size_t ysz = (size_t)yw * (L.mcus_y * 16), csz = (size_t)cw * (L.mcus_y * 8);
    const uint8_t* Y = planes + (size_t)f * L.plane_bytes;
    const int rcw = (L.width + 1) >> 1, rch = (L.height + 1) >> 1;   // real chroma size
    const int cy = r;
    const int nb0 = cy > 0 ? cy - 1 : 0;                 // neighbour chroma row of the even picture row / of the odd one
    const int nb1 = cy + 1 < rch ? cy + 1 : rch - 1;
    const int i = x4 >> 1;                               // even: (i, i + 1) is one aligned 16-bit load
    const int jm = i > 0 ? i - 1 : 0, j2 = i + 2 < cw ? i + 2 : cw - 1;
    int up[2][2][4];                                     // [channel][picture row parity][pixel]: upsampled chroma
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint8_t* C = Y + ysz + (size_t)c * csz;
        const uint8_t* rc = C + (size_t)cy * cw;
        const uint8_t* ra = C + (size_t)nb0 * cw;
        const uint8_t* rb = C + (size_t)nb1 * cw;
        const uint32_t mc = *reinterpret_cast<const uint16_t*>(rc + i), ma = *reinterpret_cast<const uint16_t*>(ra + i),
                       mb = *reinterpret_cast<const uint16_t*>(rb + i);
        const int c3[4] = {3 * (int)rc[jm], 3 * (int)(mc & 255u), 3 * (int)(mc >> 8), 3 * (int)rc[j2]};
        const int na[4] = {(int)ra[jm], (int)(ma & 255u), (int)(ma >> 8), (int)ra[j2]};
        const int nb[4] = {(int)rb[jm], (int)(mb & 255u), (int)(mb >> 8), (int)rb[j2]};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int* nn = q ? nb : na;
            const int s0 = c3[0] + nn[0], s1 = c3[1] + nn[1], s2 = c3[2] + nn[2], s3 = c3[3] + nn[3];   // columns i-1, i, i+1, i+2
            up[c][q][0] = i == 0 ? (s1 * 4 + 8) >> 4 : (s1 * 3 + s0 + 8) >> 4;
            up[c][q][1] = i == rcw - 1 ? (s1 * 4 + 7) >> 4 : (s1 * 3 + s2 + 7) >> 4;
            up[c][q][2] = (s2 * 3 + s1 + 8) >> 4;
            up[c][q][3] = i + 1 == rcw - 1 ? (s2 * 4 + 7) >> 4 : (s2 * 3 + s3 + 7) >> 4;
        }
    }
    const bool words = x4 + 4 <= L.width && ((L.width * 3) & 3) == 0;   // 12 bytes on a 4-byte boundary
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int y = y0 + q;
        if (y >= L.height) break;
        const uint32_t yy = *reinterpret_cast<const uint32_t*>(Y + (size_t)y * yw + x4);  // yw is a multiple of 16, x4 of 4
        uint8_t px[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) esdjpeg::ycc_to_bgr((int)((yy >> (8 * k)) & 255u), up[0][q][k], up[1][q][k], px + 3 * k);
        uint8_t* o = bgr + ((size_t)f * L.height + y) * (size_t)L.width * 3 + (size_t)x4 * 3;
        if (words) {
            uint32_t* ow = reinterpret_cast<uint32_t*>(o);
            ow[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
            ow[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
            ow[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
        } else {
            for (int k = 0; k < 4 && x4 + k < L.width; ++k) { o[3 * k] = px[3 * k]; o[3 * k + 1] = px[3 * k + 1]; o[3 * k + 2] = px[3 * k + 2]; }
        }
    }
}

}  // namespace

extern "C" {

int esd_decode_abi_version(void) { return ESD_DECODE_ABI_VERSION; }

const char* esd_mjpeg_last_error(const esd_mjpeg* h) { return h ? h->err.c_str() : g_open_error.c_str(); }

// process-wide origin of the ESD_DEC_TIMING=2 timelines: a device event and the host clock taken together after a synchronise
static cudaEvent_t g_origin_ev = nullptr;
static double g_origin_ms = 0.0;
static double host_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static void ensure_origin() {
    static std::once_flag once;
    std::call_once(once, [] {
        cudaEventCreate(&g_origin_ev);
        cudaDeviceSynchronize();
        cudaEventRecord(g_origin_ev, 0);
        cudaEventSynchronize(g_origin_ev);
        g_origin_ms = host_ms();
    });
}

static void collect_timing(esd_mjpeg* h, int b) {
    if (!h->timing || !h->tev_armed[b]) return;
    h->tev_armed[b] = false;
    if (h->timeline >= 2) {
        float t[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], g_origin_ev, h->tev[b][k]);
        fprintf(stderr, "[esd_decode timeline %p] host: enter %.1f stage %.1f..%.1f | device: copy %.1f entropy %.1f..%.1f done %.1f\n", (void*)h,
                h->host_t[b][0], h->host_t[b][1], h->host_t[b][2], t[0], t[1], t[2], t[3]);
    }
    for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->tev[b][k], h->tev[b][k + 1]) == cudaSuccess) h->t_sum[k] += ms;
    }
    h->t_batches++;
    cudaGetLastError();
}

void esd_mjpeg_close(esd_mjpeg* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int b = 0; b < 2; ++b) {
        if (h->done[b]) { cudaEventSynchronize(h->done[b]); cudaEventDestroy(h->done[b]); }
        collect_timing(h, b);
        for (int k = 0; k < 4; ++k) if (h->tev[b][k]) cudaEventDestroy(h->tev[b][k]);
        esdguard::gfree(h->d_out[b]);
        if (h->h_stage[b]) cudaFreeHost(h->h_stage[b]);
    }
    if (h->timing && h->t_batches)
        fprintf(stderr, "[esd_decode timing] %lld batches, %lld pictures: host staging %.2f ms, copy+clear %.2f ms, entropy %.2f ms, idct+colour %.2f ms per batch\n",
                (long long)h->t_batches, (long long)h->t_pictures, h->t_sum[3] / h->t_batches, h->t_sum[0] / h->t_batches, h->t_sum[1] / h->t_batches,
                h->t_sum[2] / h->t_batches);
    for (int b = 0; b < 2; ++b) {
        if (h->lane[b]) { cudaStreamSynchronize(h->lane[b]); cudaStreamDestroy(h->lane[b]); }
        if (h->consumed[b]) cudaEventDestroy(h->consumed[b]);
        esdguard::gfree(h->d_comp[b]); esdguard::gfree(h->d_coef[b]); esdguard::gfree(h->d_planes[b]); esdguard::gfree(h->d_bstart[b]);
    }
    if (h->state) nvjpegJpegStateDestroy(h->state);
    if (h->nj) nvjpegDestroy(h->nj);
    if (h->map) munmap(const_cast<uint8_t*>(h->map), h->map_bytes);
    if (h->fd >= 0) close(h->fd);
    cudaGetLastError();
    delete h;
}

int esd_mjpeg_open(esd_mjpeg** out, const char* path, int device, int32_t batch_frames, int32_t backend) {
    if (!out || !path) return fail(nullptr, ESD_DEC_ERR_INVALID, "esd_mjpeg_open: null argument");
    *out = nullptr;
    if (batch_frames < 1 || batch_frames > 4096) return fail(nullptr, ESD_DEC_ERR_INVALID, "esd_mjpeg_open: batch_frames must be in 1..4096");
    if (backend < ESD_JPEG_AUTO || backend > ESD_JPEG_NATIVE) return fail(nullptr, ESD_DEC_ERR_INVALID, "esd_mjpeg_open: unknown backend %d", backend);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, ESD_DEC_ERR_CUDA, "no CUDA device available (libesd_decode has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, ESD_DEC_ERR_INVALID, "esd_mjpeg_open: device %d out of range", device);
    esd_mjpeg* h = new esd_mjpeg();
    h->device = device;
    h->batch = batch_frames;
    auto bail = [&](int rc) {
        g_open_error = h->err;
        esd_mjpeg_close(h);
        return rc;
    };
    h->fd = open(path, O_RDONLY);
    if (h->fd < 0) { fail(h, ESD_DEC_ERR_IO, "Failed to open video: %s", path); return bail(ESD_DEC_ERR_IO); }
    struct stat st;
    if (fstat(h->fd, &st) != 0 || st.st_size < 12) { fail(h, ESD_DEC_ERR_IO, "%s: not a readable file", path); return bail(ESD_DEC_ERR_IO); }
    h->map_bytes = (size_t)st.st_size;
    void* m = mmap(nullptr, h->map_bytes, PROT_READ, MAP_PRIVATE, h->fd, 0);
    if (m == MAP_FAILED) { h->map = nullptr; fail(h, ESD_DEC_ERR_IO, "%s: mmap failed", path); return bail(ESD_DEC_ERR_IO); }
    h->map = static_cast<const uint8_t*>(m);
    if (!tag(h->map, "RIFF") || !tag(h->map + 8, "AVI ")) { fail(h, ESD_DEC_ERR_FORMAT, "%s: not a RIFF/AVI file", path); return bail(ESD_DEC_ERR_FORMAT); }
    AviWalk walk{h};
    walk.list(0, h->map_bytes, false);
    if (!walk.have_vids || h->width <= 0 || h->height <= 0) { fail(h, ESD_DEC_ERR_FORMAT, "%s: no video stream header", path); return bail(ESD_DEC_ERR_FORMAT); }
    if (!walk.is_mjpeg) {
        fail(h, ESD_DEC_ERR_UNSUPPORTED, "%s: video codec '%s' is not Motion-JPEG (the only codec this build decodes on the GPU: NVDEC is closed to the container)", path, walk.fourcc);
        return bail(ESD_DEC_ERR_UNSUPPORTED);
    }
    if (h->pics.empty()) { fail(h, ESD_DEC_ERR_FORMAT, "%s: no pictures in the movi list", path); return bail(ESD_DEC_ERR_FORMAT); }

    if (cudaSetDevice(device) != cudaSuccess) { fail(h, ESD_DEC_ERR_CUDA, "cudaSetDevice(%d) failed", device); return bail(ESD_DEC_ERR_CUDA); }
    const size_t frame_bytes = (size_t)h->width * h->height * 3;
    // back end: what the caller asked for, or the best one available.  NATIVE first: it needs nothing but the stream to qualify.
    if (backend == ESD_JPEG_AUTO || backend == ESD_JPEG_NATIVE) {
        esdjpeg::JpegHeader jh;
        std::string why;
        const Picture& p0 = h->pics[0];
        const bool ok = p0.offset + p0.size <= h->map_bytes && esdjpeg::parse_jpeg(h->map + p0.offset, p0.size, &jh, &why) &&
                        jh.width == h->width && jh.height == h->height;
        if (ok) {
            h->geo = esdjpeg::geometry_of(jh);
            for (int c = 0; c < 3; ++c) { h->tq[c] = jh.tq[c]; h->td[c] = jh.td[c]; h->ta[c] = jh.ta[c]; }
            h->blocks_per_frame = 6 * h->geo.mcus_x * h->geo.mcus_y + 1;  // + one spare block per picture (scratch of the flat decoder)
            h->plane_bytes = (size_t)(h->geo.mcus_x * 16) * (h->geo.mcus_y * 16) * 3 / 2;
            h->flat = jh.restart_interval == 0;
            cudaError_t e = cudaFuncSetAttribute(jpeg_entropy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kEntropyThreads * sizeof(esdjpeg::ScanTables)));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(jpeg_entropy_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kEntropyThreads * sizeof(esdjpeg::ScanTables)));
            for (int b = 0; b < h->lanes && e == cudaSuccess; ++b) {
                e = esdguard::gmalloc(&h->d_coef[b], (size_t)h->batch * h->blocks_per_frame * 64 * sizeof(int16_t));
                if (e == cudaSuccess) e = esdguard::gmalloc(&h->d_planes[b], (size_t)h->batch * h->plane_bytes);
                if (e == cudaSuccess) e = esdguard::gmalloc(&h->d_bstart[b], (size_t)h->batch * (h->blocks_per_frame + 1) * sizeof(uint32_t));
                if (e == cudaSuccess) e = cudaMemset(h->d_bstart[b], 0, (size_t)h->batch * (h->blocks_per_frame + 1) * sizeof(uint32_t));
                if (e == cudaSuccess && h->lanes > 1) e = cudaStreamCreateWithFlags(&h->lane[b], cudaStreamNonBlocking);
                if (e == cudaSuccess && h->lanes > 1) e = cudaEventCreateWithFlags(&h->consumed[b], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) {
                fail(h, ESD_DEC_ERR_CUDA, "native decoder: device buffers for %d frames could not be allocated: %s", h->batch, cudaGetErrorString(e));
                return bail(ESD_DEC_ERR_CUDA);
            }
            h->backend = ESD_JPEG_NATIVE;
        } else if (backend == ESD_JPEG_NATIVE) {
            fail(h, ESD_DEC_ERR_UNSUPPORTED, "native decoder: %s", why.empty() ? "picture size differs from the stream header" : why.c_str());
            return bail(ESD_DEC_ERR_UNSUPPORTED);
        }
    }
    const int order_auto[3] = {ESD_JPEG_HARDWARE, ESD_JPEG_GPU_HYBRID, ESD_JPEG_DEFAULT};
    const int order_one[1] = {backend};
    const int* order = backend == ESD_JPEG_AUTO ? order_auto : order_one;
    const int n_order = h->backend == ESD_JPEG_NATIVE ? 0 : (backend == ESD_JPEG_AUTO ? 3 : 1);
    nvjpegStatus_t js = NVJPEG_STATUS_NOT_INITIALIZED;
    for (int i = 0; i < n_order; ++i) {
        js = nvjpegCreateEx(nj_backend(order[i]), nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &h->nj);
        if (js == NVJPEG_STATUS_SUCCESS) {
            js = nvjpegJpegStateCreate(h->nj, &h->state);
            if (js == NVJPEG_STATUS_SUCCESS) js = nvjpegDecodeBatchedInitialize(h->nj, h->state, h->batch, 1, NVJPEG_OUTPUT_BGRI);
            if (js == NVJPEG_STATUS_SUCCESS) { h->backend = order[i]; h->initialized_batch = h->batch; break; }
            if (h->state) { nvjpegJpegStateDestroy(h->state); h->state = nullptr; }
            nvjpegDestroy(h->nj);
        }
        h->nj = nullptr;
        cudaGetLastError();
    }
    if (!h->nj && h->backend != ESD_JPEG_NATIVE) { fail(h, ESD_DEC_ERR_NVJPEG, "nvJPEG: no usable back end (last status %d)", (int)js); return bail(ESD_DEC_ERR_NVJPEG); }
    if (h->backend == ESD_JPEG_HARDWARE) {
        unsigned cores = 0;
        if (nvjpegGetHardwareDecoderInfo(h->nj, &h->hw_engines, &cores) != NVJPEG_STATUS_SUCCESS) h->hw_engines = 0;
    }
    for (int b = 0; b < 2; ++b) {
        if (esdguard::gmalloc(&h->d_out[b], frame_bytes * h->batch) != cudaSuccess || cudaEventCreateWithFlags(&h->done[b], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) {
            fail(h, ESD_DEC_ERR_CUDA, "device buffer of %d frames (%zu bytes) could not be allocated: %s", h->batch, frame_bytes * h->batch,
                 cudaGetErrorString(cudaGetLastError()));
            return bail(ESD_DEC_ERR_CUDA);
        }
    }
    h->ptrs.resize(h->batch);
    h->lens.resize(h->batch);
    h->imgs.resize(h->batch);
    *out = h;
    return ESD_DEC_OK;
}

int esd_mjpeg_get_info(const esd_mjpeg* h, esd_mjpeg_info* o) {
    if (!h || !o) return ESD_DEC_ERR_INVALID;
    memset(o, 0, sizeof *o);
    o->width = h->width; o->height = h->height;
    o->fps_num = h->fps_num; o->fps_den = h->fps_den;
    o->n_frames = (int64_t)h->pics.size();
    o->compressed_bytes = h->compressed_bytes;
    o->backend = h->backend;
    o->hw_engines = (int32_t)h->hw_engines;
    o->batch_frames = h->batch;
    return ESD_DEC_OK;
}

int esd_mjpeg_seek(esd_mjpeg* h, int64_t frame) {
    if (!h) return ESD_DEC_ERR_INVALID;
    if (frame < 0 || frame > (int64_t)h->pics.size()) return fail(h, ESD_DEC_ERR_INVALID, "seek: frame %lld outside [0, %zu]", (long long)frame, h->pics.size());
    h->pos = frame;
    return ESD_DEC_OK;
}

int esd_mjpeg_read(esd_mjpeg* h, int64_t max_frames, void* stream, uint8_t** d_bgr, int64_t* n_frames) {
    if (!h || !d_bgr || !n_frames) return ESD_DEC_ERR_INVALID;
    *d_bgr = nullptr;
    *n_frames = 0;
    const int64_t n = std::min<int64_t>(std::min<int64_t>(h->batch, max_frames), (int64_t)h->pics.size() - h->pos);
    if (n <= 0) return ESD_DEC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int b = (int)(h->reads & 1);
    const bool laned = h->backend == ESD_JPEG_NATIVE && h->lanes > 1;
    const int ln = laned ? b : 0;                 // scratch set of this batch
    cudaStream_t ds = laned ? h->lane[b] : st;    // where the decode runs
    if (laned && h->reads >= 1) {
        // everything the caller enqueued since the previous read -- the work that consumes that batch -- lies before this point
        if (cudaEventRecord(h->consumed[b ^ 1], st) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "cudaEventRecord failed");
        h->have_consumed[b ^ 1] = true;
    }
    if (h->timeline >= 2) ensure_origin();
    const double t_enter = h->timeline >= 2 ? host_ms() - g_origin_ms : 0.0;
    // the pinned staging of this slot was last read by the decode two reads ago (the event blocks instead of spinning: a job runs
    // a session per ~256 pictures in flight, and spinning waiters would take the cores the staging threads need)
    if (h->in_flight[b]) {
        cudaError_t e = cudaEventSynchronize(h->done[b]);
        if (e != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "decode of an earlier batch failed: %s", cudaGetErrorString(e));
        h->in_flight[b] = false;
        collect_timing(h, b);
    }
    timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    const bool native = h->backend == ESD_JPEG_NATIVE;
    // pinned staging of the batch: [descriptors n x 16][quantisation tables n x 3 x 64 x u16][pictures, 64-byte aligned]
    const size_t meta = native ? (((size_t)n * sizeof(esd_mjpeg::NativeDesc) + (size_t)n * 3 * 64 * sizeof(uint16_t) + 63) & ~(size_t)63) : 0;
    size_t total = meta;
    for (int64_t i = 0; i < n; ++i) total += ((size_t)h->pics[h->pos + i].size + 16 + 63) & ~(size_t)63;
    total += 64;  // the entropy kernel's four-byte look-ahead may read past the last picture
    if (total > h->h_stage_bytes[b]) {
        if (h->h_stage[b]) cudaFreeHost(h->h_stage[b]);
        h->h_stage[b] = nullptr;
        h->h_stage_bytes[b] = 0;
        const size_t want = total + total / 4 + 4096;
        if (cudaHostAlloc(&h->h_stage[b], want, cudaHostAllocDefault) != cudaSuccess)
            return fail(h, ESD_DEC_ERR_CUDA, "pinned staging of %zu bytes could not be allocated", want);
        h->h_stage_bytes[b] = want;
    }
    const size_t frame_bytes = (size_t)h->width * h->height * 3;
    size_t off = meta;
    esd_mjpeg::NativeDesc* hdesc = reinterpret_cast<esd_mjpeg::NativeDesc*>(h->h_stage[b]);
    uint16_t* hquant = reinterpret_cast<uint16_t*>(h->h_stage[b] + (size_t)n * sizeof(esd_mjpeg::NativeDesc));
    // where each picture is staged (64-byte aligned): known from the sizes, so the pictures can be staged in parallel
    std::vector<size_t> offs((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        offs[(size_t)i] = off;
        off += ((size_t)h->pics[h->pos + i].size + 16 + 63) & ~(size_t)63;
    }
    // stage_one: 0 or an error code, message in *msg.  Touches only picture i's slots of the staging block.
    auto stage_one = [&](int64_t i, std::string* msg) -> int {
        const Picture& p = h->pics[h->pos + i];
        const size_t off = offs[(size_t)i];
        char buf[160];
        if (p.offset + p.size > h->map_bytes) { snprintf(buf, sizeof buf, "picture %lld lies outside the file", (long long)(h->pos + i)); *msg = buf; return ESD_DEC_ERR_FORMAT; }
        if (!native) {
            memcpy(h->h_stage[b] + off, h->map + p.offset, p.size);
            h->ptrs[i] = h->h_stage[b] + off;
            h->lens[i] = p.size;
            memset(&h->imgs[i], 0, sizeof(nvjpegImage_t));
            h->imgs[i].channel[0] = h->d_out[b] + (size_t)i * frame_bytes;
            h->imgs[i].pitch[0] = (size_t)h->width * 3;
            return 0;
        }
        // header walk on the host (a few markers); the Huffman tables are built on the device from the picture's own DHT
        esdjpeg::JpegHeader jh;
        std::string why;
        const uint8_t* pic = h->map + p.offset;
        if (!esdjpeg::parse_jpeg(pic, p.size, &jh, &why, false)) { snprintf(buf, sizeof buf, "picture %lld: %s", (long long)(h->pos + i), why.c_str()); *msg = buf; return ESD_DEC_ERR_UNSUPPORTED; }
        if (h->flat) {
            // staged: [headers up to the scan][scan with the byte stuffing removed, on a 4-byte boundary, zero padded]
            memcpy(h->h_stage[b] + off, pic, jh.scan_offset);
            const size_t so = (jh.scan_offset + 3) & ~(size_t)3;
            bool clean = true;
            const size_t nb = esdjpeg::unstuff_scan(pic + jh.scan_offset, jh.scan_len, h->h_stage[b] + off + so, &clean);
            if (!clean) { snprintf(buf, sizeof buf, "picture %lld carries restart markers without a restart interval", (long long)(h->pos + i)); *msg = buf; return ESD_DEC_ERR_UNSUPPORTED; }
            const size_t padded = (nb + 3 + 8) & ~(size_t)3;  // two zero words behind the data
            memset(h->h_stage[b] + off + so + nb, 0, padded - nb);
            jh.scan_offset = so;
            jh.scan_len = padded / 4;  // words
        } else {
            memcpy(h->h_stage[b] + off, pic, p.size);
        }
        bool same = jh.width == h->width && jh.height == h->height && jh.restart_interval == h->geo.restart_interval;
        for (int c = 0; c < 3; ++c) same = same && jh.td[c] == h->td[c] && jh.ta[c] == h->ta[c];
        if (!same) { snprintf(buf, sizeof buf, "picture %lld changes the stream's geometry / table selectors", (long long)(h->pos + i)); *msg = buf; return ESD_DEC_ERR_UNSUPPORTED; }
        hdesc[i].off = (uint32_t)(off + jh.scan_offset);
        hdesc[i].len = (uint32_t)jh.scan_len;
        for (int t = 0; t < 4; ++t) {  // DC0, DC1, AC0, AC1: where the picture's own tables start (built on the device)
            const int tc = t >> 1, id = t & 1;
            const bool have = tc ? jh.have_ac[id] : jh.have_dc[id];
            hdesc[i].dht[t] = have ? (uint32_t)(off + jh.dht_pos[tc][id]) : 0u;
            hdesc[i].nvals[t] = (uint16_t)(have ? jh.dht_nvals[tc][id] : 0);
        }
        for (int c = 0; c < 3; ++c) memcpy(hquant + ((size_t)i * 3 + c) * 64, jh.quant[jh.tq[c]], 64 * sizeof(uint16_t));
        return 0;
    };
    {
        // The GPU decodes a 1080p picture in ~15 us; one host thread stages one in ~21 us (header walk + copy with the byte
        // stuffing removed).  Helper threads take every k-th picture so that ONE session keeps the device busy.
        // (one helper per 64 pictures: spawning threads for a 64-picture batch cost more than it saved -- staging 1.2 -> 5.9 ms)
        const int helpers = std::max(1, std::min<int>(h->stage_threads, (int)(n / 64)));
        std::vector<int> rcs((size_t)std::max(1, helpers), 0);
        std::vector<std::string> msgs((size_t)std::max(1, helpers));
        auto run = [&](int k, int stride) {
            for (int64_t i = k; i < n; i += stride) {
                const int rc = stage_one(i, &msgs[(size_t)k]);
                if (rc) { rcs[(size_t)k] = rc; return; }
            }
        };
        if (helpers <= 1) {
            run(0, 1);
        } else {
            std::vector<std::thread> th;
            for (int k = 1; k < helpers; ++k) th.emplace_back(run, k, helpers);
            run(0, helpers);
            for (auto& t : th) t.join();
        }
        for (size_t k = 0; k < rcs.size(); ++k)
            if (rcs[k]) return fail(h, rcs[k], "%s", msgs[k].c_str());
    }
    if (native) {
        if (total > h->d_comp_bytes[ln]) {  // device mirror of the staging block (grow-only; everything that used it ran on `ds` before)
            if (h->d_comp[ln]) { cudaStreamSynchronize(ds); esdguard::gfree(h->d_comp[ln]); h->d_comp[ln] = nullptr; h->d_comp_bytes[ln] = 0; }
            const size_t want = total + total / 4 + 4096;
            if (esdguard::gmalloc(&h->d_comp[ln], want) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "device staging of %zu bytes could not be allocated", want);
            h->d_comp_bytes[ln] = want;
        }
        // this slot's output was last read by the consumer of the batch two reads ago: the lane waits for that work, not for the
        // caller's whole stream (which also holds the wait for the other lane's batch)
        if (laned && h->have_consumed[b] && cudaStreamWaitEvent(ds, h->consumed[b], 0) != cudaSuccess)
            return fail(h, ESD_DEC_ERR_CUDA, "cudaStreamWaitEvent failed");
        NativeLayout L{};
        L.n = (int)n;
        L.mcus_x = h->geo.mcus_x; L.mcus_y = h->geo.mcus_y; L.restart_interval = h->geo.restart_interval;
        L.width = h->width; L.height = h->height;
        L.blocks_per_frame = h->blocks_per_frame;
        for (int c = 0; c < 3; ++c) { L.td[c] = h->td[c]; L.ta[c] = h->ta[c]; }
        L.plane_bytes = h->plane_bytes;
        if (h->timing) {
            timespec ts1;
            clock_gettime(CLOCK_MONOTONIC, &ts1);
            h->t_sum[3] += (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6;
            h->t_pictures += n;
            if (h->timeline >= 2) {
                h->host_t[b][0] = t_enter;
                h->host_t[b][1] = ts0.tv_sec * 1e3 + ts0.tv_nsec * 1e-6 - g_origin_ms;
                h->host_t[b][2] = host_ms() - g_origin_ms;
            }
            for (int k = 0; k < 4; ++k) if (!h->tev[b][k]) cudaEventCreate(&h->tev[b][k]);
            cudaEventRecord(h->tev[b][0], ds);
        }
        cudaError_t e = cudaMemcpyAsync(h->d_comp[ln], h->h_stage[b], total, cudaMemcpyHostToDevice, ds);
        // Sparse hand-off (entry list + block offsets in place of the cleared dense array) whenever every scan of the batch fits the
        // list: a symbol that carries a coefficient is at least one bit long, so a scan of at most 32 x blocks bits cannot overflow
        // 32 entries per block.  Larger pictures (> 196 KB of scan at 1080p) keep the dense hand-off.
        bool sparse = h->flat && h->parallel_entropy && h->sparse_handoff && h->d_bstart[ln] != nullptr;
        for (int64_t i = 0; i < n && sparse; ++i) sparse = (uint64_t)hdesc[i].len * 32u <= (uint64_t)(h->blocks_per_frame - 1) * 32u;
        // dense: the coefficient scratch is cleared once; after that the IDCT kernel leaves it clean (a batch that failed half-way, or
        // a sparse batch, which keeps its list there, clears again)
        if (e == cudaSuccess && !sparse && h->coef_dirty[ln]) e = cudaMemsetAsync(h->d_coef[ln], 0, (size_t)h->batch * h->blocks_per_frame * 64 * sizeof(int16_t), ds);
        h->coef_dirty[ln] = true;
        if (e != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "native decode: staging copy failed: %s", cudaGetErrorString(e));
        const esd_mjpeg::NativeDesc* ddesc = reinterpret_cast<const esd_mjpeg::NativeDesc*>(h->d_comp[ln]);
        const uint16_t* dquant = reinterpret_cast<const uint16_t*>(h->d_comp[ln] + (size_t)n * sizeof(esd_mjpeg::NativeDesc));
        const unsigned egrid = (unsigned)((n + kEntropyThreads - 1) / kEntropyThreads);
        const size_t esmem = kEntropyThreads * sizeof(esdjpeg::ScanTables);
        if (h->timing) cudaEventRecord(h->tev[b][1], ds);
        if (sparse) jpeg_entropy_parallel_kernel<true><<<(unsigned)n, kParThreads, 0, ds>>>(L, h->d_comp[ln], ddesc, h->d_coef[ln], h->d_bstart[ln]);
        else if (h->flat && h->parallel_entropy) jpeg_entropy_parallel_kernel<false><<<(unsigned)n, kParThreads, 0, ds>>>(L, h->d_comp[ln], ddesc, h->d_coef[ln], nullptr);
        else if (h->flat) jpeg_entropy_flat_kernel<<<egrid, kEntropyThreads, esmem, ds>>>(L, h->d_comp[ln], ddesc, h->d_coef[ln]);
        else jpeg_entropy_kernel<<<egrid, kEntropyThreads, esmem, ds>>>(L, h->d_comp[ln], ddesc, h->d_coef[ln]);
        if (h->timing) cudaEventRecord(h->tev[b][2], ds);
        if (sparse) jpeg_idct_sparse_kernel<<<dim3((unsigned)((h->blocks_per_frame + 127) / 128), (unsigned)n), 128, 0, ds>>>(L, h->d_coef[ln], h->d_bstart[ln], dquant, h->d_planes[ln]);
        else jpeg_idct_kernel<<<dim3((unsigned)((h->blocks_per_frame + 127) / 128), (unsigned)n), 128, 0, ds>>>(L, h->d_coef[ln], dquant, h->d_planes[ln]);
        jpeg_color_kernel<<<dim3((unsigned)(((h->width + 3) / 4 + 255) / 256), (unsigned)((h->height + 1) / 2), (unsigned)n), 256, 0, ds>>>(L, h->d_planes[ln], h->d_out[b]);
        if (h->timing) { cudaEventRecord(h->tev[b][3], ds); h->tev_armed[b] = true; }
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "native decode kernels: %s", cudaGetErrorString(e));
        h->coef_dirty[ln] = sparse;   // dense: entropy + IDCT are enqueued and the IDCT restores the zeros; sparse: the list stays
    } else {
        if (h->initialized_batch != (int)n) {  // the batched API wants exactly the initialised number of pictures (tail of the stream)
            nvjpegStatus_t js = nvjpegDecodeBatchedInitialize(h->nj, h->state, (int)n, 1, NVJPEG_OUTPUT_BGRI);
            if (js != NVJPEG_STATUS_SUCCESS) return fail(h, ESD_DEC_ERR_NVJPEG, "nvjpegDecodeBatchedInitialize(%lld) failed: status %d", (long long)n, (int)js);
            h->initialized_batch = (int)n;
        }
        nvjpegStatus_t js = nvjpegDecodeBatched(h->nj, h->state, h->ptrs.data(), h->lens.data(), h->imgs.data(), st);
        if (js != NVJPEG_STATUS_SUCCESS)
            return fail(h, ESD_DEC_ERR_NVJPEG, "nvjpegDecodeBatched failed at frame %lld: status %d (%s)", (long long)h->pos, (int)js,
                        js == NVJPEG_STATUS_JPEG_NOT_SUPPORTED ? "bitstream not supported by this back end" : js == NVJPEG_STATUS_BAD_JPEG ? "bad JPEG" : "see nvjpeg.h");
    }
    if (cudaEventRecord(h->done[b], ds) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "cudaEventRecord failed");
    if (laned && cudaStreamWaitEvent(st, h->done[b], 0) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "cudaStreamWaitEvent failed");
    h->in_flight[b] = true;
    h->reads++;
    h->pos += n;
    *d_bgr = h->d_out[b];
    *n_frames = n;
    return ESD_DEC_OK;
}

}  // extern "C"
