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
#include <string>
#include <vector>

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
    int64_t reads = 0;
    std::vector<const unsigned char*> ptrs;
    std::vector<size_t> lens;
    std::vector<nvjpegImage_t> imgs;
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

extern "C" {

int esd_decode_abi_version(void) { return ESD_DECODE_ABI_VERSION; }

const char* esd_mjpeg_last_error(const esd_mjpeg* h) { return h ? h->err.c_str() : g_open_error.c_str(); }

void esd_mjpeg_close(esd_mjpeg* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int b = 0; b < 2; ++b) {
        if (h->done[b]) { cudaEventSynchronize(h->done[b]); cudaEventDestroy(h->done[b]); }
        cudaFree(h->d_out[b]);
        if (h->h_stage[b]) cudaFreeHost(h->h_stage[b]);
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
    if (backend < ESD_JPEG_AUTO || backend > ESD_JPEG_HARDWARE) return fail(nullptr, ESD_DEC_ERR_INVALID, "esd_mjpeg_open: unknown backend %d", backend);
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
    // back end: what the caller asked for, or the best one the library grants
    const int order_auto[3] = {ESD_JPEG_HARDWARE, ESD_JPEG_GPU_HYBRID, ESD_JPEG_DEFAULT};
    const int order_one[1] = {backend};
    const int* order = backend == ESD_JPEG_AUTO ? order_auto : order_one;
    const int n_order = backend == ESD_JPEG_AUTO ? 3 : 1;
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
    if (!h->nj) { fail(h, ESD_DEC_ERR_NVJPEG, "nvJPEG: no usable back end (last status %d)", (int)js); return bail(ESD_DEC_ERR_NVJPEG); }
    if (h->backend == ESD_JPEG_HARDWARE) {
        unsigned cores = 0;
        if (nvjpegGetHardwareDecoderInfo(h->nj, &h->hw_engines, &cores) != NVJPEG_STATUS_SUCCESS) h->hw_engines = 0;
    }
    const size_t frame_bytes = (size_t)h->width * h->height * 3;
    for (int b = 0; b < 2; ++b) {
        if (cudaMalloc(&h->d_out[b], frame_bytes * h->batch) != cudaSuccess || cudaEventCreateWithFlags(&h->done[b], cudaEventDisableTiming) != cudaSuccess) {
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
    // the pinned staging of this slot was last read by the decode two reads ago
    if (h->in_flight[b]) {
        cudaError_t e = cudaEventSynchronize(h->done[b]);
        if (e != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "decode of an earlier batch failed: %s", cudaGetErrorString(e));
        h->in_flight[b] = false;
    }
    size_t total = 0;
    for (int64_t i = 0; i < n; ++i) total += ((size_t)h->pics[h->pos + i].size + 63) & ~(size_t)63;
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
    size_t off = 0;
    for (int64_t i = 0; i < n; ++i) {
        const Picture& p = h->pics[h->pos + i];
        if (p.offset + p.size > h->map_bytes) return fail(h, ESD_DEC_ERR_FORMAT, "picture %lld lies outside the file", (long long)(h->pos + i));
        memcpy(h->h_stage[b] + off, h->map + p.offset, p.size);
        h->ptrs[i] = h->h_stage[b] + off;
        h->lens[i] = p.size;
        off += ((size_t)p.size + 63) & ~(size_t)63;
        memset(&h->imgs[i], 0, sizeof(nvjpegImage_t));
        h->imgs[i].channel[0] = h->d_out[b] + (size_t)i * frame_bytes;
        h->imgs[i].pitch[0] = (size_t)h->width * 3;
    }
    if (h->initialized_batch != (int)n) {  // the batched API wants exactly the initialised number of pictures (tail of the stream)
        nvjpegStatus_t js = nvjpegDecodeBatchedInitialize(h->nj, h->state, (int)n, 1, NVJPEG_OUTPUT_BGRI);
        if (js != NVJPEG_STATUS_SUCCESS) return fail(h, ESD_DEC_ERR_NVJPEG, "nvjpegDecodeBatchedInitialize(%lld) failed: status %d", (long long)n, (int)js);
        h->initialized_batch = (int)n;
    }
    nvjpegStatus_t js = nvjpegDecodeBatched(h->nj, h->state, h->ptrs.data(), h->lens.data(), h->imgs.data(), st);
    if (js != NVJPEG_STATUS_SUCCESS)
        return fail(h, ESD_DEC_ERR_NVJPEG, "nvjpegDecodeBatched failed at frame %lld: status %d (%s)", (long long)h->pos, (int)js,
                    js == NVJPEG_STATUS_JPEG_NOT_SUPPORTED ? "bitstream not supported by this back end" : js == NVJPEG_STATUS_BAD_JPEG ? "bad JPEG" : "see nvjpeg.h");
    if (cudaEventRecord(h->done[b], st) != cudaSuccess) return fail(h, ESD_DEC_ERR_CUDA, "cudaEventRecord failed");
    h->in_flight[b] = true;
    h->reads++;
    h->pos += n;
    *d_bgr = h->d_out[b];
    *n_frames = n;
    return ESD_DEC_OK;
}

}  // extern "C"
