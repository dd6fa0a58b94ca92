// libesd_synth.so -- CUDA filler of the synthetic benchmark / test clips (synth_core.h).  Input infrastructure only:
// a separate library so that the product's libesd.so exports no generator and the reference arm never loads it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "synth_core.h"

namespace {
thread_local char g_err[256] = "";

__global__ void syn_fill_kernel(uint8_t* out, int W, int H, long long pitch, long long frame_stride, uint32_t seed,
                                const syn_frame_desc* __restrict__ descs, long long n) {
    const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)W * H;
    if (px >= per * n) return;
    const long long f = px / per;
    const int rem = (int)(px - f * per);
    const int y = rem / W, x = rem - y * W;
    const syn_frame_desc d = descs[f];
    uint8_t* o = out + f * frame_stride + (long long)y * pitch + 3LL * x;
    o[0] = (uint8_t)syn_pixel(seed, &d, x, y, 0, W, H);
    o[1] = (uint8_t)syn_pixel(seed, &d, x, y, 1, W, H);
    o[2] = (uint8_t)syn_pixel(seed, &d, x, y, 2, W, H);
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char* syn_last_error(void) { return g_err; }

// writes n frames of WxHx3 BGR described by `descs` (int32[n][8], host) into device memory; synchronises `stream`
__attribute__((visibility("default"))) int syn_fill(uint8_t* d_out, int32_t width, int32_t height, int64_t pitch,
                                                    int64_t frame_stride, uint32_t seed, const int32_t* descs, int64_t n,
                                                    int device, void* stream) {
    if (!d_out || !descs || n <= 0 || width < 1 || height < 1) { snprintf(g_err, sizeof g_err, "bad argument"); return -1; }
    cudaError_t e = cudaSetDevice(device);
    syn_frame_desc* d_desc = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (e == cudaSuccess) e = cudaMalloc(&d_desc, sizeof(syn_frame_desc) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, descs, sizeof(syn_frame_desc) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const long long total = (long long)width * height * n;
        syn_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_out, width, height, pitch, frame_stride, seed, d_desc, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_desc);
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // extern "C"
