// Internal helpers shared by the translation units of librr_b200.so.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstdio>

#include "../../include/rr_b200.h"

// thread-local error message + code passthrough
int rr_fail(int code, const char* fmt, ...);

extern std::atomic<int64_t> g_rr_launches;
inline void rr_count_launch(int n = 1) { g_rr_launches.fetch_add(n, std::memory_order_relaxed); }

#ifdef __CUDACC__
#include <cuda_runtime.h>

#define RR_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rr_fail(RR_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                                \
    } while (0)

#define RR_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        rr_count_launch();                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return rr_fail(RR_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                \
    } while (0)

#define RR_TRY(expr)                  \
    do {                              \
        int _rc = (expr);             \
        if (_rc != RR_OK) return _rc; \
    } while (0)

// Per-device high-water mark of the dynamic shared memory a kernel has been opted in for.  cudaFuncSetAttribute is
// per device, and launchers may be called from several threads (one per index handle), hence atomics.
struct RrSmemOptIn {
    std::atomic<size_t> cap[64];
    // true if `bytes` exceeds what the current device was configured for; the caller then sets the attribute and
    // calls done().  Racing callers may both set the attribute, which is harmless.
    bool needed(size_t bytes, int* device) {
        int d = 0;
        cudaGetDevice(&d);
        *device = d & 63;
        return bytes + 1 > cap[*device].load(std::memory_order_acquire);
    }
    void done(size_t bytes, int device) {
        size_t cur = cap[device].load(std::memory_order_relaxed);
        while (cur < bytes + 1 && !cap[device].compare_exchange_weak(cur, bytes + 1, std::memory_order_release)) {}
    }
};

// device-side timing of one kernel class (no-op unless rr_profile_enable(1))
struct RrProfScope {
    int cls; cudaStream_t stream; void* rec;
    RrProfScope(int cls_, cudaStream_t s);
    ~RrProfScope();
};

// order-preserving map float -> uint32 (larger float => larger key); NaN maps below -inf
__host__ __device__ inline uint32_t rr_float_key(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
    if (f != f) return 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float rr_key_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// composite 64-bit key: (score desc, index asc) <=> key desc
__host__ __device__ inline uint64_t rr_make_key(float score, uint32_t idx) {
    return ((uint64_t)rr_float_key(score) << 32) | (uint64_t)(~idx);
}
__host__ __device__ inline uint32_t rr_key_index(uint64_t k) { return ~(uint32_t)(k & 0xFFFFFFFFull); }
__host__ __device__ inline float rr_key_score(uint64_t k) { return rr_key_float((uint32_t)(k >> 32)); }
#endif
