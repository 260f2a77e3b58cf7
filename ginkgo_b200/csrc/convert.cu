// Integer-exact helpers on either side of the SpMV path ("next" rows of SURVEY.md 8f):
//   convert_ptrs_to_idxs / convert_idxs_to_ptrs  (CSR <-> COO row arrays;
//       reference/components/format_conversion_kernels.cpp)
//   csr::extract_diagonal                         (reference/matrix/csr_kernels.cpp,
//       `extract_diagonal`: first stored entry with col == row, 0 if none)
// Outputs of the index kernels are bit-exact by construction.
#include "elementwise.cuh"

namespace b200 {
namespace convert {

template <typename I>
__global__ void ptrs_to_idxs_kernel(const I* __restrict__ ptrs, int64_t num_rows, I* __restrict__ idxs)
{
    // one warp per row: rows are short on this path, and the writes of a row are contiguous
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= num_rows) return;
    const int64_t s = ptrs[row], e = ptrs[row + 1];
    for (int64_t k = s + lane; k < e; k += 32) idxs[k] = (I)row;
}

template <typename I>
__global__ void idxs_to_ptrs_kernel(const I* __restrict__ idxs, int64_t nnz, int64_t num_rows,
                                    I* __restrict__ ptrs)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nnz == 0) {
        if (i <= num_rows) ptrs[i] = 0;
        return;
    }
    if (i >= nnz) return;
    const int64_t cur = idxs[i];
    const int64_t prev = i == 0 ? -1 : (int64_t)idxs[i - 1];
    for (int64_t r = prev + 1; r <= cur; ++r) ptrs[r] = (I)i;
    if (i == nnz - 1)
        for (int64_t r = cur + 1; r <= num_rows; ++r) ptrs[r] = (I)nnz;
}

template <typename V, typename I>
__global__ void extract_diagonal_kernel(int64_t n, const I* __restrict__ rp,
                                        const I* __restrict__ ci, const V* __restrict__ va,
                                        V* __restrict__ diag)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    V d = V(0);
    for (int64_t k = rp[row]; k < (int64_t)rp[row + 1]; ++k) {
        if ((int64_t)ci[k] == row) {
            d = va[k];
            break;
        }
    }
    diag[row] = d;
}

// components::fill_array for any trivially copyable element of 1, 2, 4, 8 or 16 bytes
template <typename T>
__global__ void fill_array_kernel(T* __restrict__ data, int64_t n, T value)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        data[i] = value;
}

}  // namespace convert
}  // namespace b200

extern "C" {

/* components::fill_array (core/components/fill_array_kernels.hpp:22-25;
 * reference/components/fill_array_kernels.cpp:17-24): data[i] = value, element size 1..16 */
b200_status b200_fill_array(b200_ctx* ctx, void* data, int64_t n, const void* value_host,
                            int32_t elem_bytes)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(n >= 0, "negative size");
    if (n == 0) return B200_OK;
    B200_REQUIRE(data && value_host, "null pointer");
    const int grid = b200::grid_for(n, 256, ctx->num_sms, 16);
#define B200_FILL(T)                                                                          \
    {                                                                                         \
        T v;                                                                                  \
        memcpy(&v, value_host, sizeof(T));                                                    \
        b200::convert::fill_array_kernel<T><<<grid, 256, 0, ctx->stream>>>((T*)data, n, v);   \
    }
    switch (elem_bytes) {
    case 1: B200_FILL(uint8_t) break;
    case 2: B200_FILL(uint16_t) break;
    case 4: B200_FILL(uint32_t) break;
    case 8: B200_FILL(uint64_t) break;
    case 16: B200_FILL(ulonglong2) break;
    default: B200_REQUIRE(false, "fill_array: element size must be 1, 2, 4, 8 or 16 bytes");
    }
#undef B200_FILL
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}


#define B200_DEF_CONVERT_I(I, IT)                                                              \
    b200_status b200_convert_ptrs_to_idxs_##I(b200_ctx* ctx, const IT* ptrs, int64_t num_rows, \
                                              IT* idxs)                                        \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        if (num_rows <= 0) return B200_OK;                                                     \
        b200::convert::ptrs_to_idxs_kernel<IT>                                                 \
            <<<(unsigned)b200::ceildiv(num_rows * 32, 256), 256, 0, ctx->stream>>>(            \
                ptrs, num_rows, idxs);                                                         \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }                                                                                          \
    b200_status b200_convert_idxs_to_ptrs_##I(b200_ctx* ctx, const IT* idxs, int64_t nnz,      \
                                              int64_t num_rows, IT* ptrs)                      \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        const int64_t work = nnz == 0 ? num_rows + 1 : nnz;                                    \
        b200::convert::idxs_to_ptrs_kernel<IT>                                                 \
            <<<(unsigned)b200::ceildiv(work, 256), 256, 0, ctx->stream>>>(idxs, nnz,           \
                                                                          num_rows, ptrs);     \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }
B200_DEF_CONVERT_I(i32, int32_t)
B200_DEF_CONVERT_I(i64, int64_t)

#define B200_DEF_EXTRACT_DIAG(V, VT, I, IT)                                                    \
    b200_status b200_csr_extract_diagonal_##V##_##I(b200_ctx* ctx, int64_t n,                  \
                                                    const IT* row_ptrs, const IT* col_idxs,    \
                                                    const VT* values, VT* diag)                \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        if (n <= 0) return B200_OK;                                                            \
        b200::convert::extract_diagonal_kernel<VT, IT>                                         \
            <<<(unsigned)b200::ceildiv(n, 256), 256, 0, ctx->stream>>>(n, row_ptrs, col_idxs,  \
                                                                       values, diag);          \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }
B200_DEF_EXTRACT_DIAG(f64, double, i32, int32_t)
B200_DEF_EXTRACT_DIAG(f64, double, i64, int64_t)
B200_DEF_EXTRACT_DIAG(f32, float, i32, int32_t)
B200_DEF_EXTRACT_DIAG(f32, float, i64, int64_t)

}  // extern "C"
