// sector_probe.cu -- does skipping isolated 32-byte sectors of a row save HBM time on B200?  
// Three readers over the same 3.4 GB buffer, rows of 11 520 B (one 4K BGR row):
//   ldg_all      every sector read with one 16-byte load per sector
//   ldg_skip     4 of every 5 sectors (the 4K tap pattern: ~82 % of the sectors hold a tap, dead ones isolated)
//   ldg_skip64   the same fraction, but dead bytes in aligned 64-byte pairs (1 pair of every 5 pairs)
//   bulk_row     one cp.async.bulk per row into shared memory (what fused_score_kernel does)
//   bulk_runs    one cp.async.bulk per run of 4 live sectors (128 B of every 160 B): 72 copies per row
// Prints ms and "useful GB/s" (bytes a tap would read) for each.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int kRow = 11520;

template <int MODE>
__global__ void ldg_kernel(const uint4* __restrict__ src, long n_live, unsigned* sink) {
    // live index i -> sector index
    unsigned acc = 0;
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_live; i += stride) {
        long s;
        if (MODE == 0) s = i;
        else if (MODE == 1) s = i + i / 4;               // skip every 5th sector
        else s = (i / 8) * 10 + (i % 8);                 // skip sectors 8,9 of every 10 (an aligned 64-byte pair)
        uint4 v = __ldg(src + s * 2);                    // first 16 B of the 32-byte sector
        acc ^= v.x ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int n) { asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int RUNS, int LANES>
__global__ void bulk_kernel(const uint8_t* __restrict__ src, long n_rows) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kStages = 4;
    __shared__ __align__(8) uint64_t bars[kStages];
    uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(smem);
    int lane = threadIdx.x;
    if (lane == 0) for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    long it = 0;
    for (long r = blockIdx.x; r < n_rows; r += gridDim.x, ++it) {
        int s = (int)(it % kStages);
        uint32_t ph = (uint32_t)((it / kStages) & 1);
        if (it >= kStages) mbar_wait(bar0 + 8 * s, ph ^ 1);   // the previous use of this stage has landed
        __syncwarp();
        const uint8_t* row = src + r * (long)kRow;
        uint32_t dst = buf0 + s * kRow;
        if (RUNS == 0) {
            if (lane == 0) { mbar_expect(bar0 + 8 * s, kRow); bulk(dst, row, kRow, bar0 + 8 * s); }
        } else {
            if (lane == 0) mbar_expect(bar0 + 8 * s, 72 * 128);
            __syncwarp();
            for (int k = lane; k < 72; k += LANES) if (lane < LANES) bulk(dst + k * 128, row + k * 160, 128, bar0 + 8 * s);
        }
    }
    // drain
    for (long j = (it > kStages ? it - kStages : 0); j < it; ++j) mbar_wait(bar0 + 8 * (int)(j % kStages), (uint32_t)((j / kStages) & 1));
}

int main() {
    long n_rows = 2160L * 136;                    // 136 4K frames = 3.38 GB
    long bytes = n_rows * kRow;
    uint8_t* d; unsigned* sink;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    long sectors = bytes / 32;
    auto report = [&](const char* name, float ms, double useful) { printf("%-12s %8.3f ms  %8.1f GB/s useful  (%.1f GB/s if every byte of the buffer counted)\n", name, ms, useful / ms / 1e6, bytes / ms / 1e6); };
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        long live = sectors;
        cudaEventRecord(e0); ldg_kernel<0><<<148 * 8, 512>>>((const uint4*)d, live, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("ldg_all", ms, (double)live * 32);
        live = sectors / 5 * 4;
        cudaEventRecord(e0); ldg_kernel<1><<<148 * 8, 512>>>((const uint4*)d, live, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("ldg_skip", ms, (double)live * 32);
        live = sectors / 10 * 8;
        cudaEventRecord(e0); ldg_kernel<2><<<148 * 8, 512>>>((const uint4*)d, live, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("ldg_skip64", ms, (double)live * 32);
        CK(cudaFuncSetAttribute(bulk_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kRow));
        CK(cudaFuncSetAttribute(bulk_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kRow));
        CK(cudaFuncSetAttribute(bulk_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kRow));
        cudaEventRecord(e0); bulk_kernel<0, 1><<<148 * 4, 32, 4 * kRow>>>(d, n_rows); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("bulk_row", ms, (double)bytes * 0.8);
        cudaEventRecord(e0); bulk_kernel<1, 1><<<148 * 4, 32, 4 * kRow>>>(d, n_rows); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("bulk_runs_1", ms, (double)bytes * 0.8);
        cudaEventRecord(e0); bulk_kernel<1, 32><<<148 * 4, 32, 4 * kRow>>>(d, n_rows); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("bulk_runs_32", ms, (double)bytes * 0.8);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
