// ELL and SELL-P SpMV for sm_100a.
//
// Replaces gko::kernels::cuda::ell::{spmv,advanced_spmv}
// (reference common/cuda_hip/matrix/ell_kernels.cpp:81-440) and
// gko::kernels::cuda::sellp::{spmv,advanced_spmv}
// (reference common/cuda_hip/matrix/sellp_kernels.cpp:36-135).  Arithmetic
// contract: reference/matrix/ell_kernels.cpp:29-120 and
// reference/matrix/sellp_kernels.cpp:27-100 -- per row the stored entries are
// accumulated in storage order (i = 0 .. width-1), padding (col == -1) skipped,
// advanced form starting from beta*c and adding (alpha*val)*b.
//
// Both formats are column-major inside a (slice of) rows, so consecutive
// threads own consecutive rows and every load of values/col_idxs is a fully
// coalesced streaming load (no L1 allocation, L2 evict-first); b is gathered
// with L2 evict-last.  ELL matrices with few rows and wide rows additionally
// split each row over LANES threads (block = rows x LANES) and combine the
// partial sums through shared memory in a fixed order -- no atomics, unlike the
// reference kernel (ell_kernels.cpp:118-160).
#include "common.cuh"

namespace b200 {

// acc += sum_i val_i * b[col_i] over `len` stored entries spaced `step` apart, in storage
// order, skipping padding (col == -1) exactly like the reference loops.  The loads of kRowBatch
// entries are issued together (predicated off for padding and for positions past `len`): a row of
// 7 entries costs two rounds of (column, value) loads and gathers instead of seven.  Batches of 8
// (one round for the 7-pt stencil) were measured and REJECTED: 80 instead of 46 registers per
// thread take a third of the resident warps away and with them the bytes in flight -- ELL 77 % ->
// 63 %, SELL-P 53 % -> 41 % of the HBM roofline (profiles/r01h_ vs r02l_kernels_roofline.json).
template <typename V, typename I, bool ADVANCED>
__device__ __forceinline__ V strided_row_sum(V acc, const I* __restrict__ cols,
                                             const V* __restrict__ vals, int64_t step, int64_t len,
                                             int64_t lane_first, int64_t lane_step, V alpha,
                                             const V* __restrict__ b, int64_t b_stride,
                                             uint64_t pol_first, uint64_t pol_last)
{
    constexpr int kRowBatch = 4;
    for (int64_t i = lane_first; i < len; i += kRowBatch * lane_step) {
        I c[kRowBatch];
        V v[kRowBatch], x[kRowBatch];
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
            const int64_t ik = i + k * lane_step;
            c[k] = ik < len ? ld_stream(cols + ik * step, pol_first) : I(-1);
        }
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
            const int64_t ik = i + k * lane_step;
            v[k] = ik < len ? ld_stream(vals + ik * step, pol_first) : V(0);
        }
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k)
            x[k] = c[k] >= I(0) ? ld_gather(b + (int64_t)c[k] * b_stride, pol_last) : V(0);
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k)
            if (c[k] != I(-1)) acc += ADVANCED ? (alpha * v[k]) * x[k] : v[k] * x[k];
    }
    return acc;
}

namespace ell {

constexpr int kThreads = 256;

template <typename V, typename I, int LANES, bool ADVANCED>
__global__ void __launch_bounds__(kThreads)
    spmv_kernel(int64_t num_rows, int64_t width, int64_t stride, const I* __restrict__ col_idxs,
                const V* __restrict__ values, const V* __restrict__ alpha_p,
                const V* __restrict__ b, int64_t b_stride, int64_t num_rhs,
                const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride)
{
    constexpr int kRows = kThreads / LANES;
    __shared__ V part[LANES > 1 ? kThreads : 1];
    const int rl = threadIdx.x % kRows;  // consecutive threads -> consecutive rows
    const int lane = threadIdx.x / kRows;
    const int64_t row = blockIdx.x * (int64_t)kRows + rl;
    const int64_t j = blockIdx.y;  // right-hand side
    const uint64_t pol_first = policy_evict_first();
    const uint64_t pol_last = policy_evict_last();
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    V acc = V(0);
    const bool rv = row < num_rows;
    if (rv) {
        if (LANES == 1 && ADVANCED && beta != V(0)) acc = beta * c[row * c_stride + j];
        acc = strided_row_sum<V, I, ADVANCED>(acc, col_idxs + row, values + row, stride, width,
                                              lane, LANES, alpha, b + j, b_stride, pol_first,
                                              pol_last);
    }
    if (LANES == 1) {
        if (rv) c[row * c_stride + j] = acc;
    } else {
        part[lane * kRows + rl] = acc;
        __syncthreads();
        if (lane == 0 && rv) {
            V s = part[rl];
#pragma unroll
            for (int l = 1; l < LANES; ++l) s += part[l * kRows + rl];
            if (ADVANCED && beta != V(0)) s = beta * c[row * c_stride + j] + s;
            c[row * c_stride + j] = s;
        }
    }
}

template <typename V, typename I, bool ADVANCED>
b200_status spmv(b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t width, int64_t stride,
                 const I* col_idxs, const V* values, const V* alpha, const V* b, int64_t b_stride,
                 int64_t num_rhs, const V* beta, V* c, int64_t c_stride)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(num_rows >= 0 && width >= 0 && num_rhs >= 0, "negative size");
    B200_REQUIRE(stride >= num_rows, "ell stride smaller than num_rows");
    if (num_rows == 0 || num_rhs == 0) return B200_OK;
    B200_REQUIRE(num_rhs <= 65535, "too many right-hand sides");
    // enough rows to fill the machine -> one thread per row (reference order)
    const int64_t fill = (int64_t)ctx->num_sms * 1024;
    int lanes = 1;
    if (num_rows < fill && width >= 16) lanes = (num_rows * 8 < fill && width >= 64) ? 32 : 8;
#define B200_ELL(L)                                                                             \
    do {                                                                                        \
        dim3 grid((unsigned)ceildiv(num_rows, kThreads / L), (unsigned)num_rhs);                \
        spmv_kernel<V, I, L, ADVANCED><<<grid, kThreads, 0, ctx->stream>>>(                     \
            num_rows, width, stride, col_idxs, values, alpha, b, b_stride, num_rhs, beta, c,    \
            c_stride);                                                                          \
    } while (0)
    if (lanes == 1)
        B200_ELL(1);
    else if (lanes == 8)
        B200_ELL(8);
    else
        B200_ELL(32);
#undef B200_ELL
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

}  // namespace ell

namespace sellp {

// one thread per row; rows of a slice are consecutive threads, so with the
// default slice_size 64 every warp reads two 128/256-byte contiguous runs per
// stored column.
template <typename V, typename I, bool ADVANCED>
__global__ void __launch_bounds__(256)
    spmv_kernel(int64_t num_rows, int64_t slice_size, const uint64_t* __restrict__ slice_sets,
                const uint64_t* __restrict__ slice_lengths, const I* __restrict__ col_idxs,
                const V* __restrict__ values, const V* __restrict__ alpha_p,
                const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                V* __restrict__ c, int64_t c_stride)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t j = blockIdx.y;
    if (row >= num_rows) return;
    const uint64_t pol_first = policy_evict_first();
    const uint64_t pol_last = policy_evict_last();
    // 32-bit division when it fits (a 64-bit one costs as much as a 7-entry row)
    const int64_t slice = (num_rows >> 31) == 0 && (slice_size >> 31) == 0
                              ? (int64_t)((uint32_t)row / (uint32_t)slice_size)
                              : row / slice_size;
    const int64_t rin = row - slice * slice_size;
    const int64_t base = (int64_t)slice_sets[slice];
    const int64_t len = (int64_t)slice_lengths[slice];
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    V acc = V(0);
    if (ADVANCED && beta != V(0)) acc = c[row * c_stride + j] * beta;
    const int64_t first = base * slice_size + rin;
    acc = strided_row_sum<V, I, ADVANCED>(acc, col_idxs + first, values + first, slice_size, len, 0, 1,
                                          alpha, b + j, b_stride, pol_first, pol_last);
    c[row * c_stride + j] = acc;
}

template <typename V, typename I, bool ADVANCED>
b200_status spmv(b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t slice_size,
                 const uint64_t* slice_sets, const uint64_t* slice_lengths, const I* col_idxs,
                 const V* values, const V* alpha, const V* b, int64_t b_stride, int64_t num_rhs,
                 const V* beta, V* c, int64_t c_stride)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(num_rows >= 0 && num_rhs >= 0 && slice_size > 0, "bad size");
    if (num_rows == 0 || num_rhs == 0) return B200_OK;
    B200_REQUIRE(num_rhs <= 65535, "too many right-hand sides");
    dim3 grid((unsigned)ceildiv(num_rows, 256), (unsigned)num_rhs);
    spmv_kernel<V, I, ADVANCED><<<grid, 256, 0, ctx->stream>>>(num_rows, slice_size, slice_sets,
                                                               slice_lengths, col_idxs, values,
                                                               alpha, b, b_stride, beta, c, c_stride);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

}  // namespace sellp
}  // namespace b200

extern "C" {

#define B200_DEF_ELL_SELLP(V, VT, I, IT)                                                       \
    b200_status b200_ell_spmv_##V##_##I(                                                       \
        b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t width, int64_t stride,      \
        const IT* col_idxs, const VT* values, const VT* b, int64_t b_stride, int64_t num_rhs,  \
        VT* c, int64_t c_stride)                                                               \
    {                                                                                          \
        return b200::ell::spmv<VT, IT, false>(ctx, num_rows, num_cols, width, stride,          \
                                              col_idxs, values, nullptr, b, b_stride, num_rhs, \
                                              nullptr, c, c_stride);                           \
    }                                                                                          \
    b200_status b200_ell_advanced_spmv_##V##_##I(                                              \
        b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t width, int64_t stride,      \
        const IT* col_idxs, const VT* values, const VT* alpha, const VT* b, int64_t b_stride,  \
        int64_t num_rhs, const VT* beta, VT* c, int64_t c_stride)                              \
    {                                                                                          \
        return b200::ell::spmv<VT, IT, true>(ctx, num_rows, num_cols, width, stride, col_idxs, \
                                             values, alpha, b, b_stride, num_rhs, beta, c,     \
                                             c_stride);                                        \
    }                                                                                          \
    b200_status b200_sellp_spmv_##V##_##I(                                                     \
        b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t slice_size,                 \
        const uint64_t* slice_sets, const uint64_t* slice_lengths, const IT* col_idxs,         \
        const VT* values, const VT* b, int64_t b_stride, int64_t num_rhs, VT* c,               \
        int64_t c_stride)                                                                      \
    {                                                                                          \
        return b200::sellp::spmv<VT, IT, false>(ctx, num_rows, num_cols, slice_size,           \
                                                slice_sets, slice_lengths, col_idxs, values,   \
                                                nullptr, b, b_stride, num_rhs, nullptr, c,     \
                                                c_stride);                                     \
    }                                                                                          \
    b200_status b200_sellp_advanced_spmv_##V##_##I(                                            \
        b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t slice_size,                 \
        const uint64_t* slice_sets, const uint64_t* slice_lengths, const IT* col_idxs,         \
        const VT* values, const VT* alpha, const VT* b, int64_t b_stride, int64_t num_rhs,     \
        const VT* beta, VT* c, int64_t c_stride)                                               \
    {                                                                                          \
        return b200::sellp::spmv<VT, IT, true>(ctx, num_rows, num_cols, slice_size,            \
                                               slice_sets, slice_lengths, col_idxs, values,    \
                                               alpha, b, b_stride, num_rhs, beta, c,           \
                                               c_stride);                                      \
    }

B200_DEF_ELL_SELLP(f64, double, i32, int32_t)
B200_DEF_ELL_SELLP(f64, double, i64, int64_t)
B200_DEF_ELL_SELLP(f32, float, i32, int32_t)
B200_DEF_ELL_SELLP(f32, float, i64, int64_t)

}  // extern "C"
