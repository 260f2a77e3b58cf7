// Element-wise launch helper: one thread per (row, col) of a row-major strided
// Dense operand, columns fastest (coalesced for any stride >= cols).  Single
// column operands (the Krylov fast path) skip the index division.
#pragma once
#include "common.cuh"

namespace b200 {

template <bool SINGLE_COL, typename F>
__global__ void __launch_bounds__(256) ew_kernel(int64_t rows, int64_t cols, F f)
{
    const int64_t total = rows * cols;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    if (SINGLE_COL) {
        f(t, int64_t(0));
    } else {
        const int64_t r = t / cols;
        f(r, t - r * cols);
    }
}

template <typename F>
inline b200_status launch_ew(b200_ctx* ctx, int64_t rows, int64_t cols, F f)
{
    const int64_t total = rows * cols;
    if (total <= 0) return B200_OK;
    const int64_t grid = ceildiv(total, 256);
    if (grid > 0x7fffffffLL) {
        set_error("element-wise launch too large");
        return B200_ERR_INVALID;
    }
    if (cols == 1)
        ew_kernel<true><<<(unsigned)grid, 256, 0, ctx->stream>>>(rows, cols, f);
    else
        ew_kernel<false><<<(unsigned)grid, 256, 0, ctx->stream>>>(rows, cols, f);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

}  // namespace b200
