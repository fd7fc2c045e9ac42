// Shared host-side plumbing for the C-ABI entry points (error text, device gate).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "mapstorch_b200.h"

namespace mt {

// thread-local text of the last failure, returned by mt_last_error()
char *error_buffer();
int fail(int code, const char *fmt, ...);

// MT_OK iff the current device is compute capability 10.x (cached per device).
int require_sm100();

#define MT_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::mt::fail(MT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define MT_REQUIRE(cond, ...)                                          \
    do {                                                               \
        if (!(cond)) return ::mt::fail(MT_ERR_INVALID, __VA_ARGS__);   \
    } while (0)

// stream-ordered scratch allocation that is returned on every exit path of an entry point
struct StreamScratch {
    char *ptr = nullptr;
    cudaStream_t stream = nullptr;
    StreamScratch() = default;
    StreamScratch(const StreamScratch &) = delete;
    StreamScratch &operator=(const StreamScratch &) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        stream = s;
        keep_pool_memory();
        return cudaMallocAsync(reinterpret_cast<void **>(&ptr), bytes, s);
    }
    // The default memory pool hands freed blocks back to the driver at the next synchronisation (release threshold
    // 0), so an entry point called in a host loop -- mt_model_terms once per trial point of fit_spec(tune_params=True)
    // -- paid a driver allocation per call: 0.27 s per fit in a fresh process, 0.5-1.7 s after other work had freed
    // gigabytes.  Keep what the pool has (these scratch blocks are kilobytes).
    static void keep_pool_memory() {
        static bool done[64] = {false};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        done[dev] = true;
    }
    ~StreamScratch() {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

#ifdef __CUDACC__
// High part of an exact two-operand split: the value rounded TO NEAREST to the 11 significant bits of a tf32
// operand (low 13 mantissa bits zero).  The remainder v - hi is exact in fp32 and at most 2^-12 |v|, so what the
// tensor core drops when it truncates the remainder is below 2^-23 |v| (masking the low bits instead would leave
// a remainder twice as large, and a one-sided error).
__device__ __forceinline__ float tf32_round(float v) {
    uint32_t b;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
    return __uint_as_float(b);
}
#endif

inline int num_sms() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

}  // namespace mt
