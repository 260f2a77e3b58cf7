// Exclusive prefix sum of integers on the stream (the reference's
// components::prefix_sum_nonnegative, reference/components/prefix_sum_kernels.cpp):
// out[i] = sum_{j<i} load(j), i in [0, n).  Three launches: per-tile sums, a single-CTA
// scan of the tile sums, per-tile rescan with the tile offset.  Integer addition is
// associative, so the result is bit-exact whatever the tiling.
#pragma once
#include "common.cuh"

namespace b200 {
namespace scan {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

// exclusive scan of one value per thread across the CTA; *total = sum of all (all threads)
template <typename T>
__device__ __forceinline__ T block_exclusive(T v, T* total, T* warp_sums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (warp == 0) {
        T w = lane < nw ? warp_sums[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += up;
        }
        if (lane < nw) warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == nw - 1) warp_sums[32] = wi;   // total
    }
    __syncthreads();
    const T res = warp_sums[warp] + incl - v;
    *total = warp_sums[32];
    __syncthreads();
    return res;
}

template <typename T, typename Load>
__global__ void __launch_bounds__(kThreads) tile_sums_kernel(int64_t n, Load load, T* __restrict__ sums)
{
    __shared__ T ws[33];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    T s = T(0);
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) s += load(base + k);
    T total;
    block_exclusive(s, &total, ws);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(1024) scan_sums_kernel(int64_t m, T* __restrict__ sums)
{
    __shared__ T ws[33];
    T carry = T(0);
    for (int64_t base = 0; base < m; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const T v = i < m ? sums[i] : T(0);
        T total;
        const T ex = block_exclusive(v, &total, ws);
        if (i < m) sums[i] = carry + ex;
        carry += total;
    }
}

template <typename T, typename Load>
__global__ void __launch_bounds__(kThreads)
    tile_scan_kernel(int64_t n, Load load, const T* __restrict__ sums, T* __restrict__ out)
{
    __shared__ T ws[33];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    T v[kItems];
    T s = T(0);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = base + k < n ? load(base + k) : T(0);
        s += v[k];
    }
    T total;
    T run = block_exclusive(s, &total, ws) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// `tile_sums`: device scratch of at least ceildiv(n, kTile) elements of T
inline int64_t num_tiles(int64_t n) { return ceildiv(n, (int64_t)kTile); }

template <typename T, typename Load>
b200_status exclusive(b200_ctx* ctx, int64_t n, Load load, T* out, T* tile_sums)
{
    if (n <= 0) return B200_OK;
    const int64_t m = num_tiles(n);
    tile_sums_kernel<T, Load><<<(unsigned)m, kThreads, 0, ctx->stream>>>(n, load, tile_sums);
    B200_LAUNCH_CHECK(ctx);
    scan_sums_kernel<T><<<1, 1024, 0, ctx->stream>>>(m, tile_sums);
    B200_LAUNCH_CHECK(ctx);
    tile_scan_kernel<T, Load><<<(unsigned)m, kThreads, 0, ctx->stream>>>(n, load, tile_sums, out);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

}  // namespace scan
}  // namespace b200
