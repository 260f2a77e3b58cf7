// csr_kernels.cuh -- CSR SpMV kernels for sm_100a (shared by csr_spmv.cu and fused_cg.cu):  c = A b   and   c = alpha A b + beta c   (+ optional fused b.c)
//
// Replaces gko::kernels::cuda::csr::{spmv, advanced_spmv}
// (reference common/cuda_hip/matrix/csr_kernels.template.cpp:2353-2468); the
// arithmetic contract is the reference executor's
// (reference/matrix/csr_kernels.cpp:47-118): per row, products accumulated left
// to right, for advanced_spmv starting from beta*c (never reading c if beta==0)
// and adding (alpha*val)*b.
//
// Kernel design ("row-segmented slab" kernel, single right-hand side):
//   * Partition (b200_csr_plan, the analogue of the reference's `srow`): the
//     merge-path coordinate 2*row + row_ptrs[row] is cut into tiles of kTile
//     items, so every tile owns whole rows, at most kTile/2 of them, and fewer
//     than kTile nonzeros plus its last row.  The plan stores (first row, first
//     nonzero) per tile, so a CTA needs no dependent loads to find its slab.
//   * Persistent CTAs (3 per SM) walk the tiles round-robin with a 2-stage
//     pipeline: while tile i is processed, the col_idxs / values slabs of tile
//     i+1 are brought into shared memory by the bulk-copy engine
//     (cp.async.bulk, SASS UBLKCP, L2 evict-first, mbarrier completion) and its
//     row_ptrs slab by cp.async (LDGSTS) -- the HBM stream never waits for a
//     thread and is perfectly load balanced whatever the row lengths are.
//   * Gather phase: one nonzero per thread-slot, b[col] gathered with an L2
//     evict-last policy, val*b written back in place in shared memory.
//   * Row phase: LANES threads per row add the row's products from shared
//     memory.  LANES == 1 (average row <= 32) is strictly left to right, i.e.
//     bit-identical to the reference executor; LANES > 1 uses a fixed shuffle
//     tree; a last row that does not fit the staging buffer is summed by the
//     whole CTA.  No floating-point atomics anywhere: results are deterministic.
// Base pointers that are not 16-byte aligned take `slab_kernel`, the same
// algorithm with ordinary coalesced loads.  Multiple right-hand sides use a
// thread-per-(row,rhs) kernel with the reference's summation order.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace b200 {
namespace csr {

constexpr int kThreads = 256;
constexpr int kTile = 2048;           // merge items (2*rows + nnz) per tile
constexpr int kRowW = 2;              // weight of a row in the merge coordinate
constexpr int kMaxRows = kTile / kRowW;
constexpr int kCap = kTile + 512 + 8;  // staged nonzeros per tile (incl. alignment slack)
constexpr int kStages = 3;
constexpr int kCtasPerSm = 2;
constexpr int kGatherUnroll = (kCap + kThreads - 1) / kThreads;  // 11

template <typename I>
__global__ void plan_kernel(const I* __restrict__ row_ptrs, int64_t num_rows, int64_t num_tiles,
                            int64_t tile_items, int64_t* __restrict__ tiles)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    const int64_t d = t * tile_items;
    int64_t lo = 0, hi = num_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (kRowW * mid + (int64_t)row_ptrs[mid] >= d)
            hi = mid;
        else
            lo = mid + 1;
    }
    tiles[2 * t] = lo;
    tiles[2 * t + 1] = (int64_t)row_ptrs[lo];
}

// Optional fused dot product (Krylov: q = A p and p.q in one launch): every CTA
// adds b[row]*c[row] over its rows, publishes one partial, and the CTA that
// arrives last sums the partials in CTA order (deterministic).  `ctl`
// (optional) is the fused solvers' control block: ctl[0] != 0 (stopped) turns
// the launch into a no-op.
template <typename V>
struct DotArgs {
    V* partials;            // one per CTA
    unsigned int* counter;  // self-resetting ticket
    V* result;
    const int32_t* ctl;
};

template <typename V>
__device__ __forceinline__ void dot_epilogue(V dot_acc, const DotArgs<V>& dot, V* red, bool* is_last)
{
    const int tid = threadIdx.x;
    const V s = block_sum(dot_acc, red);
    if (tid == 0) {
        dot.partials[blockIdx.x] = s;
        __threadfence();
        const unsigned int ticket = atomicAdd(dot.counter, 1u);
        *is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (*is_last) {
        __threadfence();
        V t = V(0);
        for (int64_t k = tid; k < (int64_t)gridDim.x; k += kThreads) t += __ldcg(dot.partials + k);
        t = block_sum(t, red);
        if (tid == 0) {
            *dot.result = t;
            *dot.counter = 0u;
        }
    }
}

// Row phase shared by both kernels: rows [r0, rows_end) of the tile, products in
// prod[] indexed by (nonzero - a0), row pointers from rp(r).
template <typename V, int LANES, bool ADVANCED, bool DOT, typename RowPtr>
__device__ __forceinline__ void row_phase(int64_t r0, int64_t rows_end, int64_t a0, const V* prod,
                                          RowPtr rp, V beta, const V* __restrict__ b,
                                          int64_t b_stride, V* __restrict__ c, int64_t c_stride,
                                          V& dot_acc)
{
    constexpr int kRowsPerPass = kThreads / LANES;
    const int tid = threadIdx.x;
    const int sub = tid % LANES;
    const int64_t nrows = rows_end - r0;
    const int64_t passes = (nrows + kRowsPerPass - 1) / kRowsPerPass;
    for (int64_t ps = 0; ps < passes; ++ps) {
        const int64_t r = r0 + ps * kRowsPerPass + tid / LANES;
        const bool rv = r < rows_end;
        int64_t s = 0, e = 0;
        if (rv) {
            s = rp(r);
            e = rp(r + 1);
        }
        V acc = V(0);
        if (LANES == 1) {
            if (ADVANCED && rv && beta != V(0)) acc = c[r * c_stride] * beta;
            for (int64_t i = s; i < e; ++i) acc += prod[i - a0];
        } else {
            for (int64_t i = s + sub; i < e; i += LANES) acc += prod[i - a0];
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ADVANCED && rv && sub == 0 && beta != V(0)) acc = c[r * c_stride] * beta + acc;
        }
        if (rv && sub == 0) {
            c[r * c_stride] = acc;
            if (DOT) dot_acc += b[r * b_stride] * acc;
        }
    }
}

// last row of a tile that does not fit the staging buffer: whole-CTA sum
template <typename V, typename I, bool ADVANCED, bool DOT>
__device__ __forceinline__ void long_row(int64_t rl, int64_t sl, int64_t p1,
                                         const I* __restrict__ col_idxs,
                                         const V* __restrict__ values, V alpha, V beta,
                                         const V* __restrict__ b, int64_t b_stride,
                                         V* __restrict__ c, int64_t c_stride, V* red, V& dot_acc,
                                         uint64_t pol_first, uint64_t pol_last)
{
    const int tid = threadIdx.x;
    V acc = V(0);
    for (int64_t i = sl + tid; i < p1; i += kThreads) {
        const I col = ld_stream(col_idxs + i, pol_first);
        const V val = ld_stream(values + i, pol_first);
        const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
        acc += ADVANCED ? (alpha * val) * x : val * x;
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
        if (ADVANCED && beta != V(0)) acc = c[rl * c_stride] * beta + acc;
        c[rl * c_stride] = acc;
        if (DOT) dot_acc += b[rl * b_stride] * acc;
    }
}

template <typename V, typename I>
struct StageLayout {
    static constexpr size_t vals_off = 0;
    static constexpr size_t cols_off = vals_off + sizeof(V) * kCap;
    static constexpr size_t rp_off = (cols_off + sizeof(I) * kCap + 15) & ~size_t(15);
    static constexpr size_t bytes = (rp_off + sizeof(I) * (kMaxRows + 8) + 127) & ~size_t(127);
};

// --------------------------------------------------------------------------
// persistent, bulk-copy pipelined kernel
// --------------------------------------------------------------------------
// Row phase of the pipelined kernel: LANES threads per row read the row's column
// indices and values from the staged slab, gather b in batches of kBatch
// independent loads and accumulate in storage order (LANES == 1: exactly the
// reference's left-to-right sum).
constexpr int kBatch = 8;

template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__device__ __forceinline__ void stream_rows(int64_t r0, int64_t rows_end, int64_t a0,
                                            const V* vals_s, const I* cols_s, const I* rp_s,
                                            V alpha, V beta, const V* __restrict__ b,
                                            int64_t b_stride, V* __restrict__ c, int64_t c_stride,
                                            uint64_t pol_last, V& dot_acc)
{
    constexpr int kRowsPerPass = kThreads / LANES;
    const int tid = threadIdx.x;
    const int sub = tid % LANES;
    const int nrows = (int)(rows_end - r0);
    const int passes = (nrows + kRowsPerPass - 1) / kRowsPerPass;
    for (int ps = 0; ps < passes; ++ps) {
        const int rl = ps * kRowsPerPass + tid / LANES;
        const bool rv = rl < nrows;
        int s = 0, e = 0;
        if (rv) {
            s = (int)((int64_t)rp_s[rl] - a0);
            e = (int)((int64_t)rp_s[rl + 1] - a0);
        }
        V acc = V(0);
        if (LANES == 1 && ADVANCED && rv && beta != V(0)) acc = c[(r0 + rl) * c_stride] * beta;
        for (int i = s + sub; i < e; i += LANES * kBatch) {
            V xs[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int idx = i + k * LANES;
                xs[k] = V(0);
                if (idx < e) xs[k] = ld_gather(b + (int64_t)cols_s[idx] * b_stride, pol_last);
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const int idx = i + k * LANES;
                if (idx < e) acc += ADVANCED ? (alpha * vals_s[idx]) * xs[k] : vals_s[idx] * xs[k];
            }
        }
        if (LANES > 1) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ADVANCED && rv && sub == 0 && beta != V(0))
                acc = c[(r0 + rl) * c_stride] * beta + acc;
        }
        if (rv && sub == 0) {
            c[(r0 + rl) * c_stride] = acc;
            if (DOT) dot_acc += b[(r0 + rl) * b_stride] * acc;
        }
    }
}

template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
    slab_tma_kernel(const int64_t* __restrict__ tiles, int64_t num_tiles, int64_t nnz,
                    const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                    const V* __restrict__ values, const V* __restrict__ alpha_p,
                    const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                    V* __restrict__ c, int64_t c_stride, DotArgs<V> dot)
{
    using L = StageLayout<V, I>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ V red[32];
    __shared__ bool is_last;

    const int tid = threadIdx.x;
    if (DOT && dot.ctl && dot.ctl[0] != 0) return;

    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const uint64_t pol_last = policy_evict_last();
    const uint64_t pol_first = policy_evict_first();
    const int64_t floor4 = nnz & ~int64_t(3);
    const int64_t G = gridDim.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // tile extents (r0, p0, r1, p1); tiles past the end are empty
    auto load_ext = [&](int64_t t, int64_t (&e)[4]) {
        if (t < num_tiles) {
            const longlong2 a = *reinterpret_cast<const longlong2*>(tiles + 2 * t);
            const longlong2 bb = *reinterpret_cast<const longlong2*>(tiles + 2 * t + 2);
            e[0] = a.x;
            e[1] = a.y;
            e[2] = bb.x;
            e[3] = bb.y;
        } else {
            e[0] = e[1] = e[2] = e[3] = 0;
        }
    };

    // issue the asynchronous loads of one tile into a stage: row pointers by cp.async
    // (all threads), col_idxs / values slabs by two bulk copies (thread 0)
    auto issue = [&](const int64_t (&e)[4], int stage) {
        if (e[2] <= e[0]) return;  // empty tile
        unsigned char* sp = smem_raw + (size_t)stage * L::bytes;
        V* vals_s = reinterpret_cast<V*>(sp + L::vals_off);
        I* cols_s = reinterpret_cast<I*>(sp + L::cols_off);
        I* rp_s = reinterpret_cast<I*>(sp + L::rp_off);
        const int64_t r0 = e[0], r1 = e[2];
        for (int64_t i = tid; i <= r1 - r0; i += kThreads)
            cp_async<(int)sizeof(I)>(rp_s + i, row_ptrs + r0 + i);
        if (tid == 0) {
            const int64_t a0 = e[1] & ~int64_t(7);
            // everything before the last row always fits; the last row is staged only if it
            // fits as well (the consumer takes the same decision from the same numbers)
            int64_t pend = e[3];
            if (pend - a0 > kCap) pend = (int64_t)row_ptrs[r1 - 1];
            int64_t be = (pend + 3) & ~int64_t(3);
            if (be > floor4) be = floor4;
            const int64_t cnt = be - a0;
            if (cnt > 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(cnt * (sizeof(V) + sizeof(I))));
                tma_load_1d(vals_s, values + a0, (uint32_t)(cnt * sizeof(V)), &full_bar[stage],
                            pol_first);
                tma_load_1d(cols_s, col_idxs + a0, (uint32_t)(cnt * sizeof(I)), &full_bar[stage],
                            pol_first);
            }
        }
    };

    // ext[k]: extents of the tile k iterations ahead (k = 0 current .. kStages - 1)
    int64_t ext[kStages][4];
    int64_t t = blockIdx.x;
#pragma unroll
    for (int k = 0; k < kStages; ++k) load_ext(t + k * G, ext[k]);
#pragma unroll
    for (int k = 0; k < kStages - 1; ++k) {  // prologue: fill kStages - 1 stages
        issue(ext[k], k);
        cp_async_commit();
    }

    uint32_t uses = 0;  // bit s: parity of stage s's barrier
    V dot_acc = V(0);
    int stage = 0;
    for (; t < num_tiles; t += G) {
        // prefetch kStages - 1 tiles ahead into the stage freed by the previous iteration
        int pf = stage + kStages - 1;
        if (pf >= kStages) pf -= kStages;
        issue(ext[kStages - 1], pf);
        cp_async_commit();
        int64_t nn[4];
        load_ext(t + kStages * G, nn);

        const int64_t r0 = ext[0][0], p0 = ext[0][1], r1 = ext[0][2], p1 = ext[0][3];
        if (r1 > r0) {
            unsigned char* sp = smem_raw + (size_t)stage * L::bytes;
            V* vals_s = reinterpret_cast<V*>(sp + L::vals_off);
            I* cols_s = reinterpret_cast<I*>(sp + L::cols_off);
            const I* rp_s = reinterpret_cast<const I*>(sp + L::rp_off);

            cp_async_wait<kStages - 1>();  // this tile's row pointers have landed (own copies)
            const int64_t a0 = p0 & ~int64_t(7);
            const bool long_last = (p1 - a0) > kCap;
            const int64_t rl = r1 - 1;
            // the last row's start decides what was staged; read it from global (L2 hit, it
            // was just fetched for the cp.async) so no block barrier is needed here
            const int64_t sl = long_last ? (int64_t)row_ptrs[rl] : p1;
            const int64_t pend = long_last ? sl : p1;
            const int64_t rows_end = long_last ? rl : r1;
            int64_t be = (pend + 3) & ~int64_t(3);
            if (be > floor4) be = floor4;
            if (be - a0 > 0) {
                mbar_wait(&full_bar[stage], (uses >> stage) & 1u);
                uses ^= (1u << stage);
            }
            if (pend > be) {  // the <= 3 trailing nonzeros of the arrays (last tile only)
                const int64_t lo = be > a0 ? be : a0;
                if (lo + tid < pend) {
                    vals_s[lo + tid - a0] = values[lo + tid];
                    cols_s[lo + tid - a0] = col_idxs[lo + tid];
                }
            }
            __syncthreads();  // row pointers (cp.async of other threads) + tail visible

            stream_rows<V, I, LANES, ADVANCED, DOT>(r0, rows_end, a0, vals_s, cols_s, rp_s, alpha,
                                                    beta, b, b_stride, c, c_stride, pol_last,
                                                    dot_acc);
            if (long_last)
                long_row<V, I, ADVANCED, DOT>(rl, sl, p1, col_idxs, values, alpha, beta, b,
                                              b_stride, c, c_stride, red, dot_acc, pol_first,
                                              pol_last);
        }
        __syncthreads();  // stage is free for the bulk copy issued next iteration
#pragma unroll
        for (int k = 0; k < kStages - 1; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) ext[k][q] = ext[k + 1][q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) ext[kStages - 1][q] = nn[q];
        if (++stage == kStages) stage = 0;
    }
    cp_async_wait<0>();
    if (DOT) dot_epilogue(dot_acc, dot, red, &is_last);
}

// --------------------------------------------------------------------------
// warp-stream kernel: no block barriers at all
// --------------------------------------------------------------------------
// Every WARP walks its own sequence of small tiles (kWTile merge items, the same
// 2*row + row_ptrs[row] coordinate, so <= kWTile/2 rows and < kWTile nonzeros plus
// the last row).  Per tile the 32 lanes stream the slab with 256-bit loads
// (LDG.E.NA.EFL2.256), gather b for 8 nonzeros each, park the products in the
// warp's private shared-memory strip and then sum whole rows (LANES lanes per row,
// LANES == 1: left to right = reference order).  The slab and the extents of the
// NEXT tile are loaded into registers before the current tile is consumed, so the
// only exposed latency is the gather itself, and 24 independent warps per SM
// keep thousands of gathers in flight.  Nothing but __syncwarp is needed.
constexpr int kWTile = 256;
constexpr int kWCap = kWTile + 64 + 8;   // staged nonzeros per warp tile
constexpr int kWarpsPerCta = 8;
constexpr int kWCtasPerSm = 2;  // 128 registers: the register-resident prefetch must not spill

template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kWCtasPerSm)
    warp_stream_kernel(const int64_t* __restrict__ tiles, int64_t num_tiles, int64_t nnz,
                       const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                       const V* __restrict__ values, const V* __restrict__ alpha_p,
                       const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                       V* __restrict__ c, int64_t c_stride, DotArgs<V> dot)
{
    __shared__ __align__(16) V prod_all[kWarpsPerCta][kWCap];
    __shared__ V red[32];
    __shared__ bool is_last;
    if (DOT && dot.ctl && dot.ctl[0] != 0) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    V* prod = prod_all[warp];
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const uint64_t pol_last = policy_evict_last();
    const uint64_t pol_first = policy_evict_first();
    const int64_t W = (int64_t)gridDim.x * kWarpsPerCta;
    int64_t t = (int64_t)blockIdx.x * kWarpsPerCta + warp;

    // extents of a tile: r0, p0, r1, p1 (all lanes hold them; empty past the end)
    auto load_ext = [&](int64_t tt, int64_t (&e)[4]) {
        if (tt < num_tiles) {
            const longlong2 a = *reinterpret_cast<const longlong2*>(tiles + 2 * tt);
            const longlong2 bb = *reinterpret_cast<const longlong2*>(tiles + 2 * tt + 2);
            e[0] = a.x;
            e[1] = a.y;
            e[2] = bb.x;
            e[3] = bb.y;
        } else {
            e[0] = e[1] = e[2] = e[3] = 0;
        }
    };
    // Lane l owns the nonzeros p0 + l + 32 k: consecutive lanes read consecutive entries, so
    // the slab loads are coalesced AND the gathers of structured matrices (stencils, bands:
    // neighbouring nonzeros reference neighbouring columns) fall into few cache lines.
    auto load_slab = [&](const int64_t (&e)[4], I (&cols)[8], V (&vals)[8]) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t idx = e[1] + lane + 32 * k;
            const bool ok = idx < e[3];
            cols[k] = ok ? ld_stream(col_idxs + idx, pol_first) : I(0);
            vals[k] = ok ? ld_stream(values + idx, pol_first) : V(0);
        }
    };

    int64_t cur[4], nxt[4];
    I ncols[8];
    V nvals[8];
    I nrp = 0;  // row pointer of row r0 + lane (first pass of the row phase)
    load_ext(t, cur);
    load_ext(t + W, nxt);
    if (cur[2] > cur[0]) {
        load_slab(cur, ncols, nvals);
        if (lane <= cur[2] - cur[0]) nrp = row_ptrs[cur[0] + lane];
    }
    V dot_acc = V(0);
    for (; t < num_tiles; t += W) {
        const int64_t r0 = cur[0], p0 = cur[1], r1 = cur[2], p1 = cur[3];
        I cols[8];
        V vals[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cols[k] = ncols[k];
            vals[k] = nvals[k];
        }
        const I rp_first = nrp;
        // ---- prefetch the next tile's slab, row pointers and the extents after it
        int64_t nn[4];
        load_ext(t + 2 * W, nn);
        if (nxt[2] > nxt[0]) {
            load_slab(nxt, ncols, nvals);
            if (lane <= nxt[2] - nxt[0]) nrp = row_ptrs[nxt[0] + lane];
        }
        if (r1 > r0) {
            const bool long_last = (p1 - p0) > kWCap;
            const int64_t rl = r1 - 1;
            const int64_t sl = long_last ? (int64_t)row_ptrs[rl] : p1;
            const int64_t pend = long_last ? sl : p1;
            const int64_t rows_end = long_last ? rl : r1;
            // ---- gather + products: first 256 nonzeros from the prefetched registers
            {
                V xs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    xs[k] = V(0);
                    if (p0 + lane + 32 * k < pend)
                        xs[k] = ld_gather(b + (int64_t)cols[k] * b_stride, pol_last);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (p0 + lane + 32 * k < pend)
                        prod[lane + 32 * k] = ADVANCED ? (alpha * vals[k]) * xs[k] : vals[k] * xs[k];
            }
            // ---- (rare) the part of the last row beyond 256 that still fits the strip
            for (int64_t i = p0 + 256 + lane; i < pend; i += 32) {
                const I col = ld_stream(col_idxs + i, pol_first);
                const V val = ld_stream(values + i, pol_first);
                const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                prod[i - p0] = ADVANCED ? (alpha * val) * x : val * x;
            }
            __syncwarp();
            // ---- row phase: LANES lanes per row, rows_per_pass = 32 / LANES
            constexpr int kRpp = 32 / LANES;
            const int sub = lane % LANES;
            const int nrows = (int)(rows_end - r0);
            for (int ps = 0; ps * kRpp < nrows; ++ps) {
                const int rloc = ps * kRpp + lane / LANES;
                const bool rv = rloc < nrows;
                int64_t s = 0, e = 0;
                if (LANES == 1 && ps == 0) {
                    // row pointers of the first 33 rows came with the prefetch
                    const I nxt_rp = __shfl_down_sync(0xffffffffu, rp_first, 1);
                    s = rp_first;
                    e = (lane == 31) ? (rv ? (int64_t)row_ptrs[r0 + 32] : 0) : (int64_t)nxt_rp;
                } else if (rv) {
                    s = row_ptrs[r0 + rloc];
                    e = row_ptrs[r0 + rloc + 1];
                }
                if (!rv) s = e = 0;
                V acc = V(0);
                if (LANES == 1) {
                    if (ADVANCED && rv && beta != V(0)) acc = c[(r0 + rloc) * c_stride] * beta;
                    for (int64_t i = s; i < e; ++i) acc += prod[i - p0];
                } else {
                    for (int64_t i = s + sub; i < e; i += LANES) acc += prod[i - p0];
#pragma unroll
                    for (int o = LANES / 2; o > 0; o >>= 1)
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (ADVANCED && rv && sub == 0 && beta != V(0))
                        acc = c[(r0 + rloc) * c_stride] * beta + acc;
                }
                if (rv && sub == 0) {
                    c[(r0 + rloc) * c_stride] = acc;
                    if (DOT) dot_acc += b[(r0 + rloc) * b_stride] * acc;
                }
            }
            // ---- a last row that does not fit the strip: the whole warp sums it
            if (long_last) {
                V acc = V(0);
                for (int64_t i = sl + lane; i < p1; i += 32) {
                    const I col = ld_stream(col_idxs + i, pol_first);
                    const V val = ld_stream(values + i, pol_first);
                    const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                    acc += ADVANCED ? (alpha * val) * x : val * x;
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    if (ADVANCED && beta != V(0)) acc = c[rl * c_stride] * beta + acc;
                    c[rl * c_stride] = acc;
                    if (DOT) dot_acc += b[rl * b_stride] * acc;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cur[q] = nxt[q];
            nxt[q] = nn[q];
        }
    }
    if (DOT) dot_epilogue(dot_acc, dot, red, &is_last);
}

// --------------------------------------------------------------------------
// warp-stream kernel, software pipelined: the row sums of tile k-1 (shared memory only)
// run while the gathers of tile k and the slab loads of tile k+1 are in flight
// --------------------------------------------------------------------------
template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kWCtasPerSm)
    warp_pipe_kernel(const int64_t* __restrict__ tiles, int64_t num_tiles, int64_t nnz,
                     const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                     const V* __restrict__ values, const V* __restrict__ alpha_p,
                     const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                     V* __restrict__ c, int64_t c_stride, DotArgs<V> dot)
{
    __shared__ __align__(16) V prod_all[kWarpsPerCta][2][kWCap];
    __shared__ V red[32];
    __shared__ bool is_last;
    if (DOT && dot.ctl && dot.ctl[0] != 0) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const uint64_t pol_last = policy_evict_last();
    const uint64_t pol_first = policy_evict_first();
    const int64_t W = (int64_t)gridDim.x * kWarpsPerCta;
    int64_t t = (int64_t)blockIdx.x * kWarpsPerCta + warp;

    auto load_ext = [&](int64_t tt, int64_t (&e)[4]) {
        if (tt < num_tiles) {
            const longlong2 a = *reinterpret_cast<const longlong2*>(tiles + 2 * tt);
            const longlong2 bb = *reinterpret_cast<const longlong2*>(tiles + 2 * tt + 2);
            e[0] = a.x;
            e[1] = a.y;
            e[2] = bb.x;
            e[3] = bb.y;
        } else {
            e[0] = e[1] = e[2] = e[3] = 0;
        }
    };
    auto load_slab = [&](const int64_t (&e)[4], I (&cols)[8], V (&vals)[8]) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t idx = e[1] + lane + 32 * k;
            const bool ok = idx < e[3];
            cols[k] = ok ? ld_stream(col_idxs + idx, pol_first) : I(0);
            vals[k] = ok ? ld_stream(values + idx, pol_first) : V(0);
        }
    };
    // row sums of a finished tile from its strip (reads shared memory + a few row pointers)
    auto row_phase = [&](const V* prod, int64_t r0, int64_t p0, int nrows, I rp_first, V& dot_acc) {
        constexpr int kRpp = 32 / LANES;
        const int sub = lane % LANES;
        for (int ps = 0; ps * kRpp < nrows; ++ps) {
            const int rloc = ps * kRpp + lane / LANES;
            const bool rv = rloc < nrows;
            int64_t s = 0, e = 0;
            if (LANES == 1 && ps == 0) {
                const I nxt_rp = __shfl_down_sync(0xffffffffu, rp_first, 1);
                s = rp_first;
                e = (lane == 31) ? (rv ? (int64_t)row_ptrs[r0 + 32] : 0) : (int64_t)nxt_rp;
            } else if (rv) {
                s = row_ptrs[r0 + rloc];
                e = row_ptrs[r0 + rloc + 1];
            }
            if (!rv) s = e = 0;
            V acc = V(0);
            if (LANES == 1) {
                if (ADVANCED && rv && beta != V(0)) acc = c[(r0 + rloc) * c_stride] * beta;
                for (int64_t i = s; i < e; ++i) acc += prod[i - p0];
            } else {
                for (int64_t i = s + sub; i < e; i += LANES) acc += prod[i - p0];
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (ADVANCED && rv && sub == 0 && beta != V(0))
                    acc = c[(r0 + rloc) * c_stride] * beta + acc;
            }
            if (rv && sub == 0) {
                c[(r0 + rloc) * c_stride] = acc;
                if (DOT) dot_acc += b[(r0 + rloc) * b_stride] * acc;
            }
        }
    };

    int64_t cur[4], nxt[4];
    I cols[8];
    V vals[8];
    I rp_cur = 0;
    load_ext(t, cur);
    load_ext(t + W, nxt);
    if (cur[2] > cur[0]) {
        load_slab(cur, cols, vals);
        if (lane <= cur[2] - cur[0]) rp_cur = row_ptrs[cur[0] + lane];
    }
    // the tile whose products are parked and still have to be summed
    int64_t prev_r0 = 0, prev_p0 = 0;
    int prev_nrows = 0;
    I prev_rp = 0;
    int par = 0;
    V dot_acc = V(0);
    for (; t < num_tiles; t += W) {
        const int64_t r0 = cur[0], p0 = cur[1], r1 = cur[2], p1 = cur[3];
        const bool have = r1 > r0;
        bool long_last = false;
        int64_t rl = 0, sl = p1, pend = p1, rows_end = r1;
        V xs[8];
        if (have) {
            long_last = (p1 - p0) > kWCap;
            rl = r1 - 1;
            sl = long_last ? (int64_t)row_ptrs[rl] : p1;
            pend = long_last ? sl : p1;
            rows_end = long_last ? rl : r1;
            // 1. gathers of this tile go out first
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                xs[k] = V(0);
                if (p0 + lane + 32 * k < pend)
                    xs[k] = ld_gather(b + (int64_t)cols[k] * b_stride, pol_last);
            }
        }
        // 2. slab + row pointers of the next tile, extents of the one after it
        I ncols[8];
        V nvals[8];
        I nrp = 0;
        int64_t nn[4];
        load_ext(t + 2 * W, nn);
        if (nxt[2] > nxt[0]) {
            load_slab(nxt, ncols, nvals);
            if (lane <= nxt[2] - nxt[0]) nrp = row_ptrs[nxt[0] + lane];
        }
        // 3. row sums of the PREVIOUS tile while all of that is in flight
        if (prev_nrows > 0) row_phase(prod_all[warp][par ^ 1], prev_r0, prev_p0, prev_nrows, prev_rp, dot_acc);
        prev_nrows = 0;
        // 4. products of this tile into the other strip
        if (have) {
            V* prod = prod_all[warp][par];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (p0 + lane + 32 * k < pend)
                    prod[lane + 32 * k] = ADVANCED ? (alpha * vals[k]) * xs[k] : vals[k] * xs[k];
            for (int64_t i = p0 + 256 + lane; i < pend; i += 32) {
                const I col = ld_stream(col_idxs + i, pol_first);
                const V val = ld_stream(values + i, pol_first);
                const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                prod[i - p0] = ADVANCED ? (alpha * val) * x : val * x;
            }
            if (long_last) {
                V acc = V(0);
                for (int64_t i = sl + lane; i < p1; i += 32) {
                    const I col = ld_stream(col_idxs + i, pol_first);
                    const V val = ld_stream(values + i, pol_first);
                    const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                    acc += ADVANCED ? (alpha * val) * x : val * x;
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    if (ADVANCED && beta != V(0)) acc = c[rl * c_stride] * beta + acc;
                    c[rl * c_stride] = acc;
                    if (DOT) dot_acc += b[rl * b_stride] * acc;
                }
            }
            prev_r0 = r0;
            prev_p0 = p0;
            prev_nrows = (int)(rows_end - r0);
            prev_rp = rp_cur;
            par ^= 1;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cols[k] = ncols[k];
            vals[k] = nvals[k];
        }
        rp_cur = nrp;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cur[q] = nxt[q];
            nxt[q] = nn[q];
        }
    }
    if (prev_nrows > 0) row_phase(prod_all[warp][par ^ 1], prev_r0, prev_p0, prev_nrows, prev_rp, dot_acc);
    if (DOT) dot_epilogue(dot_acc, dot, red, &is_last);
}

// --------------------------------------------------------------------------
// warp-ring kernel: the warp-stream algorithm with the slab prefetched by the
// bulk-copy engine into a per-warp ring of shared-memory slots
// --------------------------------------------------------------------------
// Each warp owns kRing slots {values, col_idxs, row_ptrs, extents}.  While tile k is
// consumed, the slabs of tiles k+1 .. k+kRing-1 are already in flight: lane 0 issues two
// cp.async.bulk copies (SASS UBLKCP, L2 evict-first) per tile onto the slot's mbarrier,
// all lanes cp.async the <= 129 row pointers.  No slab data lives in registers, so the
// kernel needs ~half the registers of warp_stream_kernel and prefetches twice as deep.
constexpr int kRing = 3;
constexpr int kRingWarps = 8;
constexpr int kRingCtasPerSm = 2;
constexpr int kRingRp = kWTile / kRowW + 8;  // 136 row pointers per slot

template <typename V, typename I>
struct RingSlot {
    static constexpr size_t vals_off = 0;
    static constexpr size_t cols_off = sizeof(V) * kWCap;
    static constexpr size_t rp_off = (cols_off + sizeof(I) * kWCap + 15) & ~size_t(15);
    static constexpr size_t bytes = (rp_off + sizeof(I) * kRingRp + 63) & ~size_t(63);
};

template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__global__ void __launch_bounds__(kRingWarps * 32, kRingCtasPerSm)
    warp_ring_kernel(const int64_t* __restrict__ tiles, int64_t num_tiles, int64_t nnz,
                     const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                     const V* __restrict__ values, const V* __restrict__ alpha_p,
                     const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                     V* __restrict__ c, int64_t c_stride, DotArgs<V> dot)
{
    using S = RingSlot<V, I>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[kRingWarps][kRing];
    __shared__ V red[32];
    __shared__ bool is_last;
    if (DOT && dot.ctl && dot.ctl[0] != 0) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* my = smem_raw + (size_t)warp * kRing * S::bytes;
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const uint64_t pol_last = policy_evict_last();
    const uint64_t pol_first = policy_evict_first();
    const int64_t floor4 = nnz & ~int64_t(3);
    const int64_t W = (int64_t)gridDim.x * kRingWarps;
    const int64_t t0 = (int64_t)blockIdx.x * kRingWarps + warp;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) mbar_init(&bars[warp][s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    auto load_ext = [&](int64_t tt, int64_t (&e)[4]) {
        if (tt < num_tiles) {
            const longlong2 a = *reinterpret_cast<const longlong2*>(tiles + 2 * tt);
            const longlong2 bb = *reinterpret_cast<const longlong2*>(tiles + 2 * tt + 2);
            e[0] = a.x;
            e[1] = a.y;
            e[2] = bb.x;
            e[3] = bb.y;
        } else {
            e[0] = e[1] = e[2] = e[3] = 0;
        }
    };
    // staged range of a tile: [a0, pend), bulk-copied part [a0, be)
    auto staged = [&](const int64_t (&e)[4], int64_t& a0, int64_t& pend, int64_t& be) {
        a0 = e[1] & ~int64_t(3);
        pend = e[3];
        if (pend - a0 > kWCap) pend = (int64_t)row_ptrs[e[2] - 1];  // long last row: not staged
        be = (pend + 3) & ~int64_t(3);
        if (be > floor4) be = floor4;
    };
    auto issue = [&](const int64_t (&e)[4], int slot) {
        if (e[2] <= e[0]) return;
        unsigned char* sp = my + (size_t)slot * S::bytes;
        I* rp_s = reinterpret_cast<I*>(sp + S::rp_off);
        for (int64_t i = lane; i <= e[2] - e[0]; i += 32)
            cp_async<(int)sizeof(I)>(rp_s + i, row_ptrs + e[0] + i);
        if (lane == 0) {
            int64_t a0, pend, be;
            staged(e, a0, pend, be);
            const int64_t cnt = be - a0;
            if (cnt > 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[warp][slot], (uint32_t)(cnt * (sizeof(V) + sizeof(I))));
                tma_load_1d(sp + S::vals_off, values + a0, (uint32_t)(cnt * sizeof(V)),
                            &bars[warp][slot], pol_first);
                tma_load_1d(sp + S::cols_off, col_idxs + a0, (uint32_t)(cnt * sizeof(I)),
                            &bars[warp][slot], pol_first);
            }
        }
    };

    // ext[k]: extents of the tile k iterations ahead
    int64_t ext[kRing][4];
#pragma unroll
    for (int k = 0; k < kRing; ++k) load_ext(t0 + k * W, ext[k]);
#pragma unroll
    for (int k = 0; k < kRing - 1; ++k) {
        issue(ext[k], k);
        cp_async_commit();
    }
    uint32_t parity = 0;  // bit s: parity of slot s
    int slot = 0;
    V dot_acc = V(0);
    for (int64_t t = t0; t < num_tiles; t += W) {
        int pf = slot + kRing - 1;
        if (pf >= kRing) pf -= kRing;
        issue(ext[kRing - 1], pf);  // the slot freed by the previous iteration
        cp_async_commit();
        int64_t nn[4];
        load_ext(t + kRing * W, nn);

        const int64_t r0 = ext[0][0], p0 = ext[0][1], r1 = ext[0][2], p1 = ext[0][3];
        if (r1 > r0) {
            unsigned char* sp = my + (size_t)slot * S::bytes;
            V* vals_s = reinterpret_cast<V*>(sp + S::vals_off);
            I* cols_s = reinterpret_cast<I*>(sp + S::cols_off);
            const I* rp_s = reinterpret_cast<const I*>(sp + S::rp_off);
            int64_t a0, pend, be;
            staged(ext[0], a0, pend, be);
            const bool long_last = pend != p1;
            const int64_t rl = r1 - 1;
            const int64_t rows_end = long_last ? rl : r1;
            cp_async_wait<kRing - 1>();
            if (be - a0 > 0) {
                mbar_wait(&bars[warp][slot], (parity >> slot) & 1u);
                parity ^= (1u << slot);
            }
            if (pend > be) {  // the <= 3 trailing nonzeros of the arrays
                const int64_t lo = be > a0 ? be : a0;
                if (lo + lane < pend) {
                    vals_s[lo + lane - a0] = values[lo + lane];
                    cols_s[lo + lane - a0] = col_idxs[lo + lane];
                }
            }
            __syncwarp();
            // ---- gather phase: lane l owns nonzeros lead + l + 32 k; products in place
            const int cnt = (int)(pend - a0);
            const int lead = (int)(p0 - a0);
            for (int base = lead; base < cnt; base += 256) {
                V xs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + lane + 32 * k;
                    xs[k] = V(0);
                    if (i < cnt) xs[k] = ld_gather(b + (int64_t)cols_s[i] * b_stride, pol_last);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + lane + 32 * k;
                    if (i < cnt) vals_s[i] = ADVANCED ? (alpha * vals_s[i]) * xs[k] : vals_s[i] * xs[k];
                }
            }
            __syncwarp();
            // ---- row phase
            constexpr int kRpp = 32 / LANES;
            const int sub = lane % LANES;
            const int nrows = (int)(rows_end - r0);
            for (int ps = 0; ps * kRpp < nrows; ++ps) {
                const int rloc = ps * kRpp + lane / LANES;
                const bool rv = rloc < nrows;
                int s = 0, e = 0;
                if (rv) {
                    s = (int)((int64_t)rp_s[rloc] - a0);
                    e = (int)((int64_t)rp_s[rloc + 1] - a0);
                }
                V acc = V(0);
                if (LANES == 1) {
                    if (ADVANCED && rv && beta != V(0)) acc = c[(r0 + rloc) * c_stride] * beta;
                    for (int i = s; i < e; ++i) acc += vals_s[i];
                } else {
                    for (int i = s + sub; i < e; i += LANES) acc += vals_s[i];
#pragma unroll
                    for (int o = LANES / 2; o > 0; o >>= 1)
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (ADVANCED && rv && sub == 0 && beta != V(0))
                        acc = c[(r0 + rloc) * c_stride] * beta + acc;
                }
                if (rv && sub == 0) {
                    c[(r0 + rloc) * c_stride] = acc;
                    if (DOT) dot_acc += b[(r0 + rloc) * b_stride] * acc;
                }
            }
            if (long_last) {
                V acc = V(0);
                for (int64_t i = pend + lane; i < p1; i += 32) {
                    const I col = ld_stream(col_idxs + i, pol_first);
                    const V val = ld_stream(values + i, pol_first);
                    const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                    acc += ADVANCED ? (alpha * val) * x : val * x;
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    if (ADVANCED && beta != V(0)) acc = c[rl * c_stride] * beta + acc;
                    c[rl * c_stride] = acc;
                    if (DOT) dot_acc += b[rl * b_stride] * acc;
                }
            }
        }
        __syncwarp();  // the slot may be overwritten by the next iteration's bulk copy
#pragma unroll
        for (int k = 0; k < kRing - 1; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) ext[k][q] = ext[k + 1][q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) ext[kRing - 1][q] = nn[q];
        if (++slot == kRing) slot = 0;
    }
    cp_async_wait<0>();
    if (DOT) dot_epilogue(dot_acc, dot, red, &is_last);
}

// --------------------------------------------------------------------------
// fallback for unaligned base pointers: one tile per CTA, ordinary loads
// --------------------------------------------------------------------------
template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
__global__ void __launch_bounds__(kThreads, 3)
    slab_kernel(const int64_t* __restrict__ tiles, int64_t nnz, const I* __restrict__ row_ptrs,
                const I* __restrict__ col_idxs, const V* __restrict__ values,
                const V* __restrict__ alpha_p, const V* __restrict__ b, int64_t b_stride,
                const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride, DotArgs<V> dot)
{
    __shared__ __align__(16) V prod[kCap];
    __shared__ V red[32];
    __shared__ bool is_last;

    const int tid = threadIdx.x;
    if (DOT && dot.ctl && dot.ctl[0] != 0) return;
    const int64_t r0 = tiles[2 * blockIdx.x], p0 = tiles[2 * blockIdx.x + 1];
    const int64_t r1 = tiles[2 * blockIdx.x + 2], p1 = tiles[2 * blockIdx.x + 3];
    V dot_acc = V(0);
    if (r1 > r0) {
        const int64_t a0 = p0 & ~int64_t(7);
        V alpha = V(1), beta = V(0);
        if (ADVANCED) {
            alpha = *alpha_p;
            beta = *beta_p;
        }
        const uint64_t pol_last = policy_evict_last();
        const uint64_t pol_first = policy_evict_first();
        const bool long_last = (p1 - a0) > kCap;
        const int64_t rl = r1 - 1;
        const int64_t sl = long_last ? (int64_t)row_ptrs[rl] : p1;
        const int64_t pend = long_last ? sl : p1;
        const int64_t rows_end = long_last ? rl : r1;
        if ((((uintptr_t)col_idxs | (uintptr_t)values) & 31u) == 0) {
            // 8 consecutive nonzeros per thread, 256-bit streaming loads
#pragma unroll 2
            for (int64_t base = a0 + (int64_t)tid * 8; base < pend; base += (int64_t)kThreads * 8) {
                I cols[8];
                V vals[8];
                if (base + 8 <= nnz) {
                    ld_stream_x8(col_idxs + base, cols);
                    ld_stream_x8(values + base, vals);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool ok = base + k < nnz;
                        cols[k] = ok ? ld_stream(col_idxs + base + k, pol_first) : I(0);
                        vals[k] = ok ? ld_stream(values + base + k, pol_first) : V(0);
                    }
                }
                V xs[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int64_t idx = base + k;
                    xs[k] = (idx >= p0 && idx < pend)
                                ? ld_gather(b + (int64_t)cols[k] * b_stride, pol_last)
                                : V(0);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    prod[base - a0 + k] = ADVANCED ? (alpha * vals[k]) * xs[k] : vals[k] * xs[k];
            }
        } else {
            for (int64_t i = p0 + tid; i < pend; i += kThreads) {
                const I col = ld_stream(col_idxs + i, pol_first);
                const V val = ld_stream(values + i, pol_first);
                const V x = ld_gather(b + (int64_t)col * b_stride, pol_last);
                prod[i - a0] = ADVANCED ? (alpha * val) * x : val * x;
            }
        }
        __syncthreads();
        row_phase<V, LANES, ADVANCED, DOT>(
            r0, rows_end, a0, prod, [&](int64_t r) { return (int64_t)row_ptrs[r]; }, beta, b, b_stride,
            c, c_stride, dot_acc);
        if (long_last)
            long_row<V, I, ADVANCED, DOT>(rl, sl, p1, col_idxs, values, alpha, beta, b, b_stride, c,
                                          c_stride, red, dot_acc, pol_first, pol_last);
    }
    if (DOT) dot_epilogue(dot_acc, dot, red, &is_last);
}

// thread per (row, rhs): reference summation order, any strides
template <typename V, typename I, bool ADVANCED>
__global__ void __launch_bounds__(256)
    multi_rhs_kernel(int64_t num_rows, int64_t num_rhs, const I* __restrict__ row_ptrs,
                     const I* __restrict__ col_idxs, const V* __restrict__ values,
                     const V* __restrict__ alpha_p, const V* __restrict__ b, int64_t b_stride,
                     const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride)
{
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const int64_t total = num_rows * num_rhs;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / num_rhs;
        const int64_t j = t - row * num_rhs;
        const int64_t s = row_ptrs[row], e = row_ptrs[row + 1];
        V acc = V(0);
        if (ADVANCED && beta != V(0)) acc = c[row * c_stride + j] * beta;
        for (int64_t k = s; k < e; ++k) {
            const V val = values[k];
            const V x = b[(int64_t)col_idxs[k] * b_stride + j];
            acc += ADVANCED ? (alpha * val) * x : val * x;
        }
        c[row * c_stride + j] = acc;
    }
}

inline int pick_lanes(int64_t num_rows, int64_t nnz)
{
    const double avg = num_rows > 0 ? (double)nnz / (double)num_rows : 0.0;
    if (avg <= 32.0) return 1;
    if (avg <= 64.0) return 2;
    if (avg <= 128.0) return 4;
    if (avg <= 256.0) return 8;
    if (avg <= 512.0) return 16;
    return 32;
}

inline int64_t num_tiles_for(int64_t num_rows, int64_t nnz)
{
    return ceildiv(kRowW * num_rows + nnz, kTile);
}
inline int64_t num_wtiles_for(int64_t num_rows, int64_t nnz)
{
    return ceildiv(kRowW * num_rows + nnz, kWTile);
}

}  // namespace csr
}  // namespace b200

struct b200_csr_plan {
    int64_t num_rows = 0;
    int64_t nnz = 0;
    int64_t num_tiles = 0;
    int64_t* tiles = nullptr;  // device, 2 * (num_tiles + 1): (first row, first nonzero)
    int64_t num_wtiles = 0;
    int64_t* wtiles = nullptr;  // same for the warp-stream kernel's kWTile-item tiles
    int lanes = 1;
    int device = 0;
    int variant = -1;  // kernel variant chosen by b200_csr_plan_tune_*, -1 = not tuned
    // Column-blocked copy of the matrix (b200_csr_plan_tune_* builds it when it wins): part p
    // holds, row by row, the entries with column in [col_split[p], col_split[p+1]).  Rows are
    // column-sorted (checked), so applying the parts in order -- part 0 as c = A0 b, part p
    // as c = 1*Ap b + 1*c -- adds every row's products in exactly the original order: same
    // bits, but the gathers of one launch stay inside a slice of b that fits in L2.
    static constexpr int kMaxParts = 4;
    int parts = 0;
    const void* src_cols = nullptr;  // the arrays the copy was made from (identity check)
    const void* src_vals = nullptr;
    void* part_row_ptrs[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    void* part_cols[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    void* part_vals[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    int64_t part_nnz[kMaxParts] = {0, 0, 0, 0};
    b200_csr_plan* part_plan[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    void* ones = nullptr;  // device {1, 1} in the value type
};

namespace b200 {
namespace csr {

template <typename I>
b200_status fill_plan(b200_ctx* ctx, int64_t num_rows, int64_t nnz, const I* row_ptrs,
                      int64_t num_tiles, int64_t* tiles, int64_t tile_items = kTile)
{
    const int block = 256;
    const int grid = (int)ceildiv(num_tiles + 1, block);
    plan_kernel<I><<<grid, block, 0, ctx->stream>>>(row_ptrs, num_rows, num_tiles, tile_items,
                                                    tiles);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename K>
b200_status set_smem(K kernel, size_t bytes)
{
    static thread_local const void* done = nullptr;  // one attribute call per kernel / thread
    if (done != (const void*)kernel) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)bytes));
        B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared));
        done = (const void*)kernel;
    }
    return B200_OK;
}

// kernel variants
enum Variant { kSlab = 0, kTma = 1, kWarp = 2, kRingV = 3, kPipe = 4 };

inline Variant pick_variant(const void* col_idxs, const void* values,
                            const b200_csr_plan* plan = nullptr)
{
    const uintptr_t a = (uintptr_t)col_idxs | (uintptr_t)values;
    static const char* env = getenv("B200_CSR_KERNEL");
    Variant want = (plan && plan->variant >= 0) ? (Variant)plan->variant : kWarp;
    if (env && !strcmp(env, "tma")) want = kTma;
    if (env && !strcmp(env, "slab")) want = kSlab;
    if (env && !strcmp(env, "ring")) want = kRingV;
    if (env && !strcmp(env, "pipe")) want = kPipe;
    if (want == kRingV && (a & 15u)) want = kWarp;
    if (want == kTma && (a & 15u)) want = kSlab;   // bulk copies need 16-byte alignment
    return want;
}

// tile array / tile count a variant works on
inline int64_t variant_tiles(Variant v, int64_t num_rows, int64_t nnz)
{
    return (v == kWarp || v == kRingV || v == kPipe) ? num_wtiles_for(num_rows, nnz) : num_tiles_for(num_rows, nnz);
}

// number of CTAs a launch uses (the size of the fused-dot partials array)
inline int grid_size(const b200_ctx* ctx, Variant v, int64_t num_tiles)
{
    if (v == kSlab) return (int)num_tiles;
    if (v == kTma) {
        const int64_t cap = (int64_t)ctx->num_sms * kCtasPerSm;
        return (int)(num_tiles < cap ? num_tiles : cap);
    }
    if (v == kRingV) {
        const int64_t need = ceildiv(num_tiles, kRingWarps);
        const int64_t cap = (int64_t)ctx->num_sms * kRingCtasPerSm;
        return (int)(need < cap ? need : cap);
    }
    const int64_t need = ceildiv(num_tiles, kWarpsPerCta);
    const int64_t cap = (int64_t)ctx->num_sms * kWCtasPerSm;
    return (int)(need < cap ? need : cap);
}
inline int max_grid_size(const b200_ctx* ctx) { return ctx->num_sms * 4; }

template <typename V, typename I, int LANES, bool ADVANCED, bool DOT>
b200_status launch_one(b200_ctx* ctx, Variant v, int64_t num_tiles, const int64_t* tiles,
                       int64_t nnz, const I* row_ptrs, const I* col_idxs, const V* values,
                       const V* alpha, const V* b, int64_t b_stride, const V* beta, V* c,
                       int64_t c_stride, DotArgs<V> dot, int grid)
{
    if (v == kTma) {
        constexpr size_t smem = StageLayout<V, I>::bytes * kStages;
        auto k = slab_tma_kernel<V, I, LANES, ADVANCED, DOT>;
        b200_status st = set_smem(k, smem);
        if (st != B200_OK) return st;
        k<<<grid, kThreads, smem, ctx->stream>>>(tiles, num_tiles, nnz, row_ptrs, col_idxs, values,
                                                 alpha, b, b_stride, beta, c, c_stride, dot);
    } else if (v == kRingV) {
        constexpr size_t smem = RingSlot<V, I>::bytes * kRing * kRingWarps;
        auto k = warp_ring_kernel<V, I, LANES, ADVANCED, DOT>;
        b200_status st = set_smem(k, smem);
        if (st != B200_OK) return st;
        static bool dbg = getenv("B200_DEBUG") != nullptr;
        if (dbg) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, kRingWarps * 32, smem);
            fprintf(stderr, "[b200] warp_ring_kernel: %d CTAs/SM, smem %zu, grid %d, tiles %lld\n", nb,
                    smem, grid, (long long)num_tiles);
            dbg = false;
        }
        k<<<grid, kRingWarps * 32, smem, ctx->stream>>>(tiles, num_tiles, nnz, row_ptrs, col_idxs,
                                                        values, alpha, b, b_stride, beta, c,
                                                        c_stride, dot);
    } else if (v == kPipe) {
        warp_pipe_kernel<V, I, LANES, ADVANCED, DOT><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(
            tiles, num_tiles, nnz, row_ptrs, col_idxs, values, alpha, b, b_stride, beta, c, c_stride,
            dot);
    } else if (v == kWarp) {
        auto k = warp_stream_kernel<V, I, LANES, ADVANCED, DOT>;
        static bool dbg = getenv("B200_DEBUG") != nullptr;
        if (dbg) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, kWarpsPerCta * 32, 0);
            fprintf(stderr, "[b200] warp_stream_kernel: %d CTAs/SM, grid %d, tiles %lld\n", nb, grid,
                    (long long)num_tiles);
            dbg = false;
        }
        k<<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(tiles, num_tiles, nnz, row_ptrs, col_idxs,
                                                       values, alpha, b, b_stride, beta, c, c_stride,
                                                       dot);
    } else {
        slab_kernel<V, I, LANES, ADVANCED, DOT><<<(unsigned)num_tiles, kThreads, 0, ctx->stream>>>(
            tiles, nnz, row_ptrs, col_idxs, values, alpha, b, b_stride, beta, c, c_stride, dot);
    }
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename V, typename I, bool ADVANCED, bool DOT>
b200_status launch_slab(b200_ctx* ctx, int lanes, Variant v, int64_t num_tiles,
                        const int64_t* tiles, int64_t nnz, const I* row_ptrs, const I* col_idxs,
                        const V* values, const V* alpha, const V* b, int64_t b_stride,
                        const V* beta, V* c, int64_t c_stride, DotArgs<V> dot = DotArgs<V>{})
{
    if (num_tiles <= 0) return B200_OK;
    const int grid = grid_size(ctx, v, num_tiles);
#define B200_SLAB(L)                                                                           \
    return launch_one<V, I, L, ADVANCED, DOT>(ctx, v, num_tiles, tiles, nnz, row_ptrs,         \
                                              col_idxs, values, alpha, b, b_stride, beta, c,   \
                                              c_stride, dot, grid)
    switch (lanes) {
    case 1: B200_SLAB(1);
    case 2: B200_SLAB(2);
    case 4: B200_SLAB(4);
    case 8: B200_SLAB(8);
    case 16: B200_SLAB(16);
    default: B200_SLAB(32);
    }
#undef B200_SLAB
}

}  // namespace csr
}  // namespace b200
