// guard_alloc.h -- device allocations with red zones, the library's own stand-in for compute-sanitizer's memcheck (which this
// pool does not allow).  With ESD_GUARD=1 in the environment every device buffer of the library is allocated with a 256-byte
// zone of 0xA5 before and after it; the zones are verified when the buffer is freed (and on demand), and a damaged zone aborts
// the process with the buffer's size and the first damaged byte -- an out-of-bounds WRITE next to any internal buffer cannot pass
// the test-suite silently.  Without the variable the wrappers are plain cudaMalloc / cudaFree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

namespace esdguard {

constexpr size_t kZone = 256;

inline bool enabled() {
    static const bool on = getenv("ESD_GUARD") && atoi(getenv("ESD_GUARD")) != 0;
    return on;
}
struct Registry {
    std::mutex m;
    std::map<void*, size_t> live;  // user pointer -> user bytes
    long long checked = 0;
    ~Registry() {
        if (enabled()) fprintf(stderr, "[esd guard] %lld device buffers verified, %zu still live at exit\n", checked, live.size());
    }
};
inline Registry& reg() { static Registry r; return r; }

inline void verify(void* user, size_t bytes, const char* when) {
    std::vector<uint8_t> z(2 * kZone);
    uint8_t* base = static_cast<uint8_t*>(user) - kZone;
    if (cudaMemcpy(z.data(), base, kZone, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(z.data() + kZone, static_cast<uint8_t*>(user) + bytes, kZone, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return;  // the context is gone (process teardown): nothing to read
    }
    for (size_t i = 0; i < 2 * kZone; ++i)
        if (z[i] != 0xA5) {
            fprintf(stderr, "[esd guard] RED ZONE DAMAGED (%s): buffer of %zu bytes, %s zone, byte %zu holds 0x%02x\n", when, bytes,
                    i < kZone ? "leading" : "trailing", i < kZone ? i : i - kZone, z[i]);
            abort();
        }
    reg().checked++;
}

template <typename T>
inline cudaError_t gmalloc(T** p, size_t bytes) {
    if (!enabled()) return cudaMalloc(p, bytes);
    uint8_t* base = nullptr;
    const size_t padded = (bytes + 255) & ~(size_t)255;  // keep the trailing zone aligned; the slack belongs to the zone check too
    cudaError_t e = cudaMalloc(&base, padded + 2 * kZone);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, 0xA5, padded + 2 * kZone);
    // the fill runs on the legacy stream, which non-blocking streams do not wait for: without this the library's first use of the
    // buffer on one of its own streams can be overtaken by the fill (seen: the decoder's staging mirror turned into 0xA5 descriptors)
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(base); return e; }
    *p = reinterpret_cast<T*>(base + kZone);
    std::lock_guard<std::mutex> lk(reg().m);
    reg().live[base + kZone] = bytes;
    return cudaSuccess;
}

inline cudaError_t gfree(void* user) {
    if (!enabled() || !user) return cudaFree(user);
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(reg().m);
        auto it = reg().live.find(user);
        if (it == reg().live.end()) return cudaFree(user);  // not ours (allocated before the switch was read)
        bytes = it->second;
        reg().live.erase(it);
    }
    cudaDeviceSynchronize();
    // bytes between the user size and its 256-byte round-up were filled with the pattern as well: check from the exact end
    verify(user, bytes, "free");
    return cudaFree(static_cast<uint8_t*>(user) - kZone);
}

}  // namespace esdguard
