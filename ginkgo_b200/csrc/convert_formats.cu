// Format conversions that create the other SpMV formats from CSR on the device
// (SURVEY.md 8f rank 1).  Every output array is bit-exact against the reference kernels:
//   ell::compute_max_row_nnz          reference/matrix/ell_kernels.cpp:130-140
//   csr::convert_to_ell               reference/matrix/csr_kernels.cpp:573-598
//   sellp::compute_slice_sets         reference/matrix/sellp_kernels.cpp:107-130
//   csr::convert_to_sellp             reference/matrix/csr_kernels.cpp:528-567
//   csr::compute_hybrid_coo_row_ptrs  reference/matrix/csr_kernels.cpp (prefix sum of
//                                     max(row_nnz - ell_lim, 0))
//   csr::convert_to_hybrid            reference/matrix/csr_kernels.cpp:910-953
//   Hybrid::imbalance_limit           include/ginkgo/core/matrix/hybrid.hpp:222-243 (the
//                                     order statistic of the row lengths, here from a device
//                                     histogram instead of a host std::sort)
//   csr::sort_by_column_index         reference/matrix/csr_kernels.cpp:1272-1290
// Layout notes: ELL element (row, i) lives at row + i * stride, SELL-P element (row r of
// slice s, i) at (slice_sets[s] + i) * slice_size + r; padding is column -1 / value 0.
// All kernels map consecutive threads to consecutive rows, so the column-major writes are
// coalesced; the CSR reads of a warp walk 32 neighbouring rows and are served from L1.
#include "scan.cuh"

namespace b200 {
namespace convert {

template <typename I>
__global__ void max_row_nnz_kernel(const I* __restrict__ rp, int64_t num_rows,
                                   unsigned long long* __restrict__ out)
{
    unsigned long long m = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < num_rows; r += stride) {
        const unsigned long long len = (unsigned long long)((int64_t)rp[r + 1] - (int64_t)rp[r]);
        m = len > m ? len : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

template <typename V, typename I>
__global__ void to_ell_kernel(int64_t num_rows, const I* __restrict__ rp, const I* __restrict__ ci,
                              const V* __restrict__ va, int64_t width, int64_t stride,
                              I* __restrict__ ecols, V* __restrict__ evals)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const int64_t s = rp[row];
    const int64_t len = (int64_t)rp[row + 1] - s;
    for (int64_t i = 0; i < width; ++i) {
        const bool in = i < len;
        ecols[row + i * stride] = in ? ci[s + i] : I(-1);
        evals[row + i * stride] = in ? va[s + i] : V(0);
    }
}

// one warp per slice: maximum row length of the slice, rounded up to the stride factor
template <typename I>
__global__ void slice_lengths_kernel(const I* __restrict__ rp, int64_t num_rows, int64_t slice_size,
                                     int64_t stride_factor, int64_t num_slices,
                                     uint64_t* __restrict__ slice_lengths)
{
    const int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slice >= num_slices) return;
    unsigned long long m = 0;
    for (int64_t lr = lane; lr < slice_size; lr += 32) {
        const int64_t row = slice * slice_size + lr;
        if (row < num_rows) {
            const unsigned long long len = (unsigned long long)((int64_t)rp[row + 1] - (int64_t)rp[row]);
            const unsigned long long padded =
                (len + stride_factor - 1) / (unsigned long long)stride_factor * stride_factor;
            m = padded > m ? padded : m;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if (lane == 0) slice_lengths[slice] = m;
}

template <typename V, typename I>
__global__ void to_sellp_kernel(int64_t num_rows, int64_t slice_size,
                                const uint64_t* __restrict__ slice_sets,
                                const uint64_t* __restrict__ slice_lengths,
                                const I* __restrict__ rp, const I* __restrict__ ci,
                                const V* __restrict__ va, I* __restrict__ scols,
                                V* __restrict__ svals)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const int64_t slice = row / slice_size, lr = row - slice * slice_size;
    const int64_t s = rp[row];
    const int64_t len = (int64_t)rp[row + 1] - s;
    const int64_t width = (int64_t)slice_lengths[slice];
    int64_t out = (int64_t)slice_sets[slice] * slice_size + lr;
    for (int64_t i = 0; i < width; ++i, out += slice_size) {
        const bool in = i < len;
        scols[out] = in ? ci[s + i] : I(-1);
        svals[out] = in ? va[s + i] : V(0);
    }
}

template <typename V, typename I>
__global__ void to_hybrid_kernel(int64_t num_rows, const I* __restrict__ rp,
                                 const I* __restrict__ ci, const V* __restrict__ va,
                                 int64_t ell_lim, int64_t ell_stride, I* __restrict__ ecols,
                                 V* __restrict__ evals, const int64_t* __restrict__ coo_row_ptrs,
                                 I* __restrict__ crows, I* __restrict__ ccols,
                                 V* __restrict__ cvals)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= ell_stride) return;
    if (row >= num_rows) {  // padding rows of the ELL part are initialised too
        for (int64_t i = 0; i < ell_lim; ++i) {
            ecols[row + i * ell_stride] = I(-1);
            evals[row + i * ell_stride] = V(0);
        }
        return;
    }
    const int64_t s = rp[row];
    const int64_t len = (int64_t)rp[row + 1] - s;
    for (int64_t i = 0; i < ell_lim; ++i) {
        const bool in = i < len;
        ecols[row + i * ell_stride] = in ? ci[s + i] : I(-1);
        evals[row + i * ell_stride] = in ? va[s + i] : V(0);
    }
    int64_t out = coo_row_ptrs[row];
    for (int64_t i = ell_lim; i < len; ++i, ++out) {
        crows[out] = (I)row;
        ccols[out] = ci[s + i];
        cvals[out] = va[s + i];
    }
}

template <typename I>
__global__ void row_nnz_histogram_kernel(const I* __restrict__ rp, int64_t num_rows,
                                         unsigned long long* __restrict__ bins)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < num_rows; r += stride)
        atomicAdd(bins + ((int64_t)rp[r + 1] - (int64_t)rp[r]), 1ull);
}

// position k of the sorted row lengths: the bin b with scan[b] <= k < scan[b + 1]
__global__ void select_bin_kernel(const unsigned long long* __restrict__ scanned, int64_t num_bins,
                                  unsigned long long k, unsigned long long* __restrict__ out)
{
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= num_bins) return;
    if (scanned[b] <= k && k < scanned[b + 1]) *out = (unsigned long long)b;
}

// ---- sort_by_column_index ------------------------------------------------------------
// Rank-by-counting (stable): the new position of an entry is the number of entries of its
// row with a smaller column, or an equal column and a smaller original position.  For rows
// with distinct columns -- the reference's precondition for a well-defined result, its
// std::sort is not stable -- this is exactly the reference order.
template <typename V, typename I>
__global__ void sort_short_rows_kernel(int64_t num_rows, const I* __restrict__ rp,
                                       I* __restrict__ ci, V* __restrict__ va)
{
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= num_rows) return;
    const int64_t s = rp[row];
    const int len = (int)((int64_t)rp[row + 1] - s);
    if (len < 2 || len > 32) return;
    const bool has = lane < len;
    const I col = has ? ci[s + lane] : I(0);
    const V val = has ? va[s + lane] : V(0);
    int rank = 0;
    for (int j = 0; j < len; ++j) {
        const I cj = __shfl_sync(0xffffffffu, col, j);
        rank += (cj < col || (cj == col && j < lane)) ? 1 : 0;
    }
    __syncwarp();
    if (has) {
        ci[s + rank] = col;
        va[s + rank] = val;
    }
}

constexpr int kSortThreads = 256;
constexpr int kSortSmem = 2048;  // entries of a long row held in shared memory

// one CTA per row longer than 32 entries (rows are visited grid-stride; short rows skipped)
template <typename V, typename I>
__global__ void __launch_bounds__(kSortThreads)
    sort_long_rows_kernel(int64_t num_rows, const I* __restrict__ rp, I* __restrict__ ci,
                          V* __restrict__ va, I* __restrict__ tmp_cols, V* __restrict__ tmp_vals,
                          int64_t tmp_per_cta)
{
    __shared__ I scol[kSortSmem];
    __shared__ V sval[kSortSmem];
    for (int64_t row = blockIdx.x; row < num_rows; row += gridDim.x) {
        const int64_t s = rp[row];
        const int64_t len = (int64_t)rp[row + 1] - s;
        if (len <= 32) continue;
        if (len <= kSortSmem) {
            for (int64_t i = threadIdx.x; i < len; i += kSortThreads) {
                scol[i] = ci[s + i];
                sval[i] = va[s + i];
            }
            __syncthreads();
            for (int64_t i = threadIdx.x; i < len; i += kSortThreads) {
                const I c = scol[i];
                int64_t rank = 0;
                for (int64_t j = 0; j < len; ++j) {
                    const I cj = scol[j];
                    rank += (cj < c || (cj == c && j < i)) ? 1 : 0;
                }
                ci[s + rank] = c;
                va[s + rank] = sval[i];
            }
            __syncthreads();
        } else {
            I* tc = tmp_cols + (int64_t)blockIdx.x * tmp_per_cta;
            V* tv = tmp_vals + (int64_t)blockIdx.x * tmp_per_cta;
            for (int64_t i = threadIdx.x; i < len; i += kSortThreads) {
                tc[i] = ci[s + i];
                tv[i] = va[s + i];
            }
            __syncthreads();
            for (int64_t i = threadIdx.x; i < len; i += kSortThreads) {
                const I c = tc[i];
                int64_t rank = 0;
                for (int64_t j = 0; j < len; ++j) {
                    const I cj = tc[j];
                    rank += (cj < c || (cj == c && j < i)) ? 1 : 0;
                }
                ci[s + rank] = c;
                va[s + rank] = tv[i];
            }
            __syncthreads();
        }
    }
}

template <typename I>
b200_status max_row_nnz(b200_ctx* ctx, const I* row_ptrs, int64_t num_rows, int64_t* out_host)
{
    B200_REQUIRE(ctx && out_host, "null argument");
    *out_host = 0;
    if (num_rows <= 0) return B200_OK;
    B200_REQUIRE(row_ptrs, "null pointer");
    unsigned long long* d = (unsigned long long*)ctx->scratch(sizeof(unsigned long long));
    if (!d) return B200_ERR_ALLOC;
    B200_CUDA_CHECK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
    max_row_nnz_kernel<I><<<grid_for(num_rows, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(
        row_ptrs, num_rows, d);
    B200_LAUNCH_CHECK(ctx);
    unsigned long long h = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out_host = (int64_t)h;
    return B200_OK;
}

template <typename I>
b200_status slice_sets(b200_ctx* ctx, const I* row_ptrs, int64_t num_rows, int64_t slice_size,
                       int64_t stride_factor, uint64_t* slice_sets_out, uint64_t* slice_lengths)
{
    B200_REQUIRE(ctx && slice_sets_out, "null argument");
    B200_REQUIRE(slice_size > 0 && stride_factor > 0 && num_rows >= 0, "bad slice shape");
    const int64_t ns = ceildiv(num_rows, slice_size);
    if (ns > 0) {
        B200_REQUIRE(row_ptrs && slice_lengths, "null pointer");
        slice_lengths_kernel<I><<<(unsigned)ceildiv(ns * 32, 256), 256, 0, ctx->stream>>>(
            row_ptrs, num_rows, slice_size, stride_factor, ns, slice_lengths);
        B200_LAUNCH_CHECK(ctx);
    }
    uint64_t* sums = (uint64_t*)ctx->scratch(sizeof(uint64_t) * (size_t)scan::num_tiles(ns + 1));
    if (!sums) return B200_ERR_ALLOC;
    const uint64_t* sl = slice_lengths;
    return scan::exclusive<uint64_t>(
        ctx, ns + 1, [=] __device__(int64_t i) -> uint64_t { return i < ns ? sl[i] : 0ull; },
        slice_sets_out, sums);
}

template <typename I>
b200_status hybrid_coo_row_ptrs(b200_ctx* ctx, const I* row_ptrs, int64_t num_rows,
                                int64_t ell_lim, int64_t* coo_row_ptrs)
{
    B200_REQUIRE(ctx && coo_row_ptrs, "null argument");
    B200_REQUIRE(num_rows >= 0 && ell_lim >= 0, "negative size");
    B200_REQUIRE(num_rows == 0 || row_ptrs, "null pointer");
    int64_t* sums = (int64_t*)ctx->scratch(sizeof(int64_t) * (size_t)scan::num_tiles(num_rows + 1));
    if (!sums) return B200_ERR_ALLOC;
    return scan::exclusive<int64_t>(
        ctx, num_rows + 1,
        [=] __device__(int64_t r) -> int64_t {
            if (r >= num_rows) return 0;
            const int64_t len = (int64_t)row_ptrs[r + 1] - (int64_t)row_ptrs[r];
            return len > ell_lim ? len - ell_lim : 0;
        },
        coo_row_ptrs, sums);
}

template <typename I>
b200_status order_statistic(b200_ctx* ctx, const I* row_ptrs, int64_t num_rows, int64_t k,
                            int64_t* value_host)
{
    B200_REQUIRE(ctx && value_host, "null argument");
    B200_REQUIRE(num_rows > 0 && k >= 0 && k < num_rows, "k outside [0, num_rows)");
    int64_t max_nnz = 0;
    b200_status st = max_row_nnz<I>(ctx, row_ptrs, num_rows, &max_nnz);
    if (st != B200_OK) return st;
    // scratch: bins[nb + 1] (histogram, scanned in place; entry nb = total) | result | tile sums
    const int64_t nb = max_nnz + 1;
    const size_t tiles = (size_t)scan::num_tiles(nb + 1);
    unsigned long long* bins =
        (unsigned long long*)ctx->scratch(sizeof(unsigned long long) * ((size_t)nb + 2 + tiles));
    if (!bins) return B200_ERR_ALLOC;
    unsigned long long* res = bins + nb + 1;
    unsigned long long* sums = bins + nb + 2;
    B200_CUDA_CHECK(cudaMemsetAsync(bins, 0, sizeof(unsigned long long) * (nb + 2), ctx->stream));
    row_nnz_histogram_kernel<I><<<grid_for(num_rows, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(
        row_ptrs, num_rows, bins);
    B200_LAUNCH_CHECK(ctx);
    const unsigned long long* cb = bins;
    st = scan::exclusive<unsigned long long>(
        ctx, nb + 1,
        [=] __device__(int64_t i) -> unsigned long long { return i < nb ? cb[i] : 0ull; }, bins,
        sums);
    if (st != B200_OK) return st;
    select_bin_kernel<<<(unsigned)ceildiv(nb, 256), 256, 0, ctx->stream>>>(
        bins, nb, (unsigned long long)k, res);
    B200_LAUNCH_CHECK(ctx);
    unsigned long long h = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&h, res, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *value_host = (int64_t)h;
    return B200_OK;
}

}  // namespace convert
}  // namespace b200

extern "C" {

#define B200_DEF_CONVERT_FMT_I(I, IT)                                                          \
    b200_status b200_ell_compute_max_row_nnz_##I(b200_ctx* ctx, const IT* row_ptrs,            \
                                                 int64_t num_rows, int64_t* max_nnz_host)      \
    {                                                                                          \
        return b200::convert::max_row_nnz<IT>(ctx, row_ptrs, num_rows, max_nnz_host);          \
    }                                                                                          \
    b200_status b200_sellp_compute_slice_sets_##I(                                             \
        b200_ctx* ctx, const IT* row_ptrs, int64_t num_rows, int64_t slice_size,               \
        int64_t stride_factor, uint64_t* slice_sets, uint64_t* slice_lengths)                  \
    {                                                                                          \
        return b200::convert::slice_sets<IT>(ctx, row_ptrs, num_rows, slice_size,              \
                                             stride_factor, slice_sets, slice_lengths);        \
    }                                                                                          \
    /* coo_row_ptrs[r] = sum_{q<r} max(len_q - ell_lim, 0), num_rows + 1 entries */            \
    b200_status b200_csr_compute_hybrid_coo_row_ptrs_##I(b200_ctx* ctx, const IT* row_ptrs,    \
                                                         int64_t num_rows, int64_t ell_lim,    \
                                                         int64_t* coo_row_ptrs)                \
    {                                                                                          \
        return b200::convert::hybrid_coo_row_ptrs<IT>(ctx, row_ptrs, num_rows, ell_lim,        \
                                                      coo_row_ptrs);                           \
    }                                                                                          \
    /* value at position k (0-based) of the ascending row lengths */                           \
    b200_status b200_csr_row_nnz_order_statistic_##I(b200_ctx* ctx, const IT* row_ptrs,        \
                                                     int64_t num_rows, int64_t k,              \
                                                     int64_t* value_host)                      \
    {                                                                                          \
        return b200::convert::order_statistic<IT>(ctx, row_ptrs, num_rows, k, value_host);     \
    }
B200_DEF_CONVERT_FMT_I(i32, int32_t)
B200_DEF_CONVERT_FMT_I(i64, int64_t)

#define B200_DEF_CONVERT_FMT(V, VT, I, IT)                                                     \
    b200_status b200_csr_convert_to_ell_##V##_##I(                                             \
        b200_ctx* ctx, int64_t num_rows, const IT* row_ptrs, const IT* col_idxs,               \
        const VT* values, int64_t num_stored_per_row, int64_t ell_stride, IT* ell_col_idxs,    \
        VT* ell_values)                                                                        \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        B200_REQUIRE(num_rows >= 0 && num_stored_per_row >= 0 && ell_stride >= num_rows,       \
                     "bad ELL shape");                                                         \
        if (num_rows == 0 || num_stored_per_row == 0) return B200_OK;                          \
        B200_REQUIRE(row_ptrs && ell_col_idxs && ell_values, "null pointer");                  \
        b200::convert::to_ell_kernel<VT, IT>                                                   \
            <<<(unsigned)b200::ceildiv(num_rows, 128), 128, 0, ctx->stream>>>(                 \
                num_rows, row_ptrs, col_idxs, values, num_stored_per_row, ell_stride,          \
                ell_col_idxs, ell_values);                                                     \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }                                                                                          \
    b200_status b200_csr_convert_to_sellp_##V##_##I(                                           \
        b200_ctx* ctx, int64_t num_rows, int64_t slice_size, const uint64_t* slice_sets,       \
        const uint64_t* slice_lengths, const IT* row_ptrs, const IT* col_idxs,                 \
        const VT* values, IT* sellp_col_idxs, VT* sellp_values)                                \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        B200_REQUIRE(num_rows >= 0 && slice_size > 0, "bad SELL-P shape");                     \
        if (num_rows == 0) return B200_OK;                                                     \
        B200_REQUIRE(row_ptrs && slice_sets && slice_lengths, "null pointer");                 \
        b200::convert::to_sellp_kernel<VT, IT>                                                 \
            <<<(unsigned)b200::ceildiv(num_rows, 128), 128, 0, ctx->stream>>>(                 \
                num_rows, slice_size, slice_sets, slice_lengths, row_ptrs, col_idxs, values,   \
                sellp_col_idxs, sellp_values);                                                 \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }                                                                                          \
    b200_status b200_csr_convert_to_hybrid_##V##_##I(                                          \
        b200_ctx* ctx, int64_t num_rows, const IT* row_ptrs, const IT* col_idxs,               \
        const VT* values, int64_t ell_lim, int64_t ell_stride, IT* ell_col_idxs,               \
        VT* ell_values, const int64_t* coo_row_ptrs, IT* coo_row_idxs, IT* coo_col_idxs,       \
        VT* coo_values)                                                                        \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        B200_REQUIRE(num_rows >= 0 && ell_lim >= 0 && ell_stride >= num_rows,                  \
                     "bad hybrid shape");                                                      \
        if (ell_stride == 0) return B200_OK;                                                   \
        B200_REQUIRE(num_rows == 0 || (row_ptrs && coo_row_ptrs), "null pointer");             \
        b200::convert::to_hybrid_kernel<VT, IT>                                                \
            <<<(unsigned)b200::ceildiv(ell_stride, 128), 128, 0, ctx->stream>>>(               \
                num_rows, row_ptrs, col_idxs, values, ell_lim, ell_stride, ell_col_idxs,       \
                ell_values, coo_row_ptrs, coo_row_idxs, coo_col_idxs, coo_values);             \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }                                                                                          \
    b200_status b200_csr_sort_by_column_index_##V##_##I(b200_ctx* ctx, int64_t num_rows,       \
                                                        const IT* row_ptrs, IT* col_idxs,      \
                                                        VT* values)                            \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        if (num_rows <= 0) return B200_OK;                                                     \
        B200_REQUIRE(row_ptrs, "null pointer");                                                \
        int64_t max_nnz = 0;                                                                   \
        b200_status st = b200::convert::max_row_nnz<IT>(ctx, row_ptrs, num_rows, &max_nnz);    \
        if (st != B200_OK) return st;                                                          \
        if (max_nnz < 2) return B200_OK;                                                       \
        b200::convert::sort_short_rows_kernel<VT, IT>                                          \
            <<<(unsigned)b200::ceildiv(num_rows * 32, 256), 256, 0, ctx->stream>>>(            \
                num_rows, row_ptrs, col_idxs, values);                                         \
        B200_LAUNCH_CHECK(ctx);                                                                \
        if (max_nnz <= 32) return B200_OK;                                                     \
        int64_t grid = (int64_t)ctx->num_sms * 4;                                              \
        if (grid > num_rows) grid = num_rows;                                                  \
        int64_t per_cta = 0;                                                                   \
        IT* tc = nullptr;                                                                      \
        VT* tv = nullptr;                                                                      \
        if (max_nnz > b200::convert::kSortSmem) {                                              \
            per_cta = (max_nnz + 15) & ~int64_t(15);                                           \
            const int64_t cap = (int64_t(1) << 30) / (per_cta * (int64_t)(sizeof(IT) + sizeof(VT))); \
            if (grid > cap) grid = cap < 1 ? 1 : cap;                                          \
            char* buf = (char*)ctx->scratch((size_t)grid * per_cta * (sizeof(IT) + sizeof(VT))); \
            if (!buf) return B200_ERR_ALLOC;                                                   \
            tv = (VT*)buf;                                                                     \
            tc = (IT*)(buf + (size_t)grid * per_cta * sizeof(VT));                             \
        }                                                                                      \
        b200::convert::sort_long_rows_kernel<VT, IT>                                           \
            <<<(unsigned)grid, b200::convert::kSortThreads, 0, ctx->stream>>>(                 \
                num_rows, row_ptrs, col_idxs, values, tc, tv, per_cta);                        \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }
B200_DEF_CONVERT_FMT(f64, double, i32, int32_t)
B200_DEF_CONVERT_FMT(f64, double, i64, int64_t)
B200_DEF_CONVERT_FMT(f32, float, i32, int32_t)
B200_DEF_CONVERT_FMT(f32, float, i64, int64_t)

}  // extern "C"
