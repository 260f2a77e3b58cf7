// TEST INFRASTRUCTURE ONLY -- stands in for ginkgo_b200/csrc/elementwise.cuh (+ common.cuh) when
// tests/test_kernel_sources_cpu.py compiles a COPY of an element-wise .cu file (krylov_steps.cu)
// with plain g++: the extended lambdas of the kernels become ordinary lambdas and launch_ew runs
// them in a host loop over (row, col).  This checks the arithmetic written in the kernel bodies
// against the oracle without a GPU; the launch mechanics themselves are exercised by the GPU
// tests.  Compiled with -ffp-contract=off, the counterpart of nvcc's -fmad=false.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "ginkgo_b200.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__

struct b200_ctx {
    int64_t launches = 0;
};

namespace b200 {

using std::sqrt;

inline void set_error(const char*, ...) {}

constexpr uint8_t kFinalizedMask = 1u << 6;
constexpr uint8_t kIdMask = (1u << 6) - 1u;
inline bool has_stopped(uint8_t s) { return (s & kIdMask) != 0; }
inline bool is_finalized(uint8_t s) { return (s & kFinalizedMask) != 0; }

#define B200_REQUIRE(cond, msg)                 \
    do {                                        \
        if (!(cond)) return B200_ERR_INVALID;   \
    } while (0)

template <typename F>
inline b200_status launch_ew(b200_ctx* ctx, int64_t rows, int64_t cols, F f)
{
    if (rows * cols <= 0) return B200_OK;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) f(i, j);
    ctx->launches++;
    return B200_OK;
}

}  // namespace b200
