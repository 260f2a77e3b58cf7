// TEST INFRASTRUCTURE ONLY -- stands in for ginkgo_b200/csrc/elementwise.cuh (+ common.cuh) when
// tests/test_kernel_sources_cpu.py / tests/test_dist_assembly_cpu.py compile a COPY of an
// element-wise .cu file (krylov_steps.cu, dist_assembly.cu) with plain g++: the extended lambdas of the kernels become ordinary lambdas and launch_ew runs
// them in a host loop over (row, col).  This checks the arithmetic written in the kernel bodies
// against the oracle without a GPU; the launch mechanics themselves are exercised by the GPU
// tests.  Compiled with -ffp-contract=off, the counterpart of nvcc's -fmad=false.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#ifdef B200_SHIM_THREADS
#include <thread>
#include <vector>
#endif

#include "ginkgo_b200.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__

// the tests hand in a zero-filled 64-byte buffer as the context
struct b200_ctx {
    int64_t launches;
    void* scratch_buf;
    size_t scratch_cap;
    int stream;
    void* scratch(size_t bytes)
    {
        if (bytes > scratch_cap) {
            std::free(scratch_buf);
            scratch_buf = std::malloc(bytes);
            scratch_cap = bytes;
        }
        return scratch_buf;
    }
};

// the few CUDA runtime / device names an element-wise file may use, on host memory
enum cudaMemcpyKind { cudaMemcpyDeviceToHost = 2 };
inline int cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, int)
{
    std::memcpy(dst, src, bytes);
    return 0;
}
inline int cudaStreamSynchronize(int) { return 0; }
#define B200_CUDA_CHECK(expr) \
    do {                      \
        if ((expr) != 0) return B200_ERR_CUDA; \
    } while (0)
inline int __popc(unsigned int w) { return __builtin_popcount(w); }
inline int __ffs(unsigned int w) { return __builtin_ffs((int)w); }
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

namespace b200 {

using std::sqrt;

inline void set_error(const char*, ...) {}
inline int64_t ceildiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr uint8_t kFinalizedMask = 1u << 6;
constexpr uint8_t kIdMask = (1u << 6) - 1u;
inline bool has_stopped(uint8_t s) { return (s & kIdMask) != 0; }
inline bool is_finalized(uint8_t s) { return (s & kFinalizedMask) != 0; }

#define B200_REQUIRE(cond, msg)                 \
    do {                                        \
        if (!(cond)) return B200_ERR_INVALID;   \
    } while (0)

// The host loop may visit the (row, col) pairs in any order -- a GPU does.  -DB200_SHIM_ORDER=1 runs
// them backwards, =2 in a scrambled order (a stride coprime to the count), so that a kernel body
// which silently relies on ascending execution order fails on the CPU already.
#ifndef B200_SHIM_ORDER
#define B200_SHIM_ORDER 0
#endif
template <typename F>
inline b200_status launch_ew(b200_ctx* ctx, int64_t rows, int64_t cols, F f)
{
    const int64_t total = rows * cols;
    if (total <= 0) return B200_OK;
    int64_t stride = 1;
    if (B200_SHIM_ORDER == 2) {
        stride = total / 2 + 1;
        auto gcd = [](int64_t a, int64_t b) {
            while (b) {
                const int64_t t = a % b;
                a = b;
                b = t;
            }
            return a;
        };
        while (gcd(stride, total) != 1) ++stride;
    }
#ifdef B200_SHIM_THREADS
    // real concurrency (for ThreadSanitizer runs): the index space is dealt round-robin to threads
    {
        std::vector<std::thread> pool;
        for (int w = 0; w < B200_SHIM_THREADS; ++w)
            pool.emplace_back([=] {
                for (int64_t k = w; k < total; k += B200_SHIM_THREADS) f(k / cols, k % cols);
            });
        for (auto& th : pool) th.join();
        ctx->launches++;
        return B200_OK;
    }
#endif
    for (int64_t k = 0; k < total; ++k) {
        int64_t t = k;
        if (B200_SHIM_ORDER == 1) t = total - 1 - k;
        if (B200_SHIM_ORDER == 2) t = (int64_t)(((__int128)k * stride + 7) % total);
        f(t / cols, t % cols);
    }
    ctx->launches++;
    return B200_OK;
}

}  // namespace b200
