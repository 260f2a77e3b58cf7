// Dense BLAS-1 for the Krylov loops: dot / conj_dot / norm2 / squared_norm2 and
// the axpy/scal family.  Replaces gko::kernels::cuda::dense::* (reference
// common/unified/matrix/dense_kernels.template.cpp:29-365 and the cuBLAS
// dispatch common/cuda_hip/matrix/dense_kernels.cpp:667-733); arithmetic
// contract reference/matrix/dense_kernels.cpp:95-437.
//
// Reductions: ONE launch per reduction.  Every CTA reduces a strided slab with
// a fixed shuffle/shared-memory tree, publishes its partial, and the CTA that
// arrives last (self-resetting atomic ticket) adds the partials in index order
// and writes the 1 x cols result (sqrt fused for norm2).  No floating-point
// atomics: bit-reproducible run to run.
#include "elementwise.cuh"

namespace b200 {
namespace dense {

enum class Red { dot, sqnorm };

template <typename V, Red OP, bool SQRT>
__global__ void __launch_bounds__(512)
    reduce_col1_kernel(int64_t rows, const V* __restrict__ x, int64_t xs, const V* __restrict__ y,
                       int64_t ys, V* __restrict__ partials, unsigned int* __restrict__ counter,
                       V* __restrict__ result)
{
    __shared__ V red[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    V acc0 = V(0), acc1 = V(0), acc2 = V(0), acc3 = V(0);
    int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
    for (; i + 3 * step < rows; i += 4 * step) {
        const V a0 = x[i * xs], a1 = x[(i + step) * xs], a2 = x[(i + 2 * step) * xs],
                a3 = x[(i + 3 * step) * xs];
        V b0 = a0, b1 = a1, b2 = a2, b3 = a3;
        if (OP == Red::dot) {
            b0 = y[i * ys];
            b1 = y[(i + step) * ys];
            b2 = y[(i + 2 * step) * ys];
            b3 = y[(i + 3 * step) * ys];
        }
        acc0 += a0 * b0;
        acc1 += a1 * b1;
        acc2 += a2 * b2;
        acc3 += a3 * b3;
    }
    for (; i < rows; i += step) {
        const V a = x[i * xs];
        const V b = OP == Red::dot ? y[i * ys] : a;
        acc0 += a * b;
    }
    V acc = block_sum((acc0 + acc1) + (acc2 + acc3), red);
    if (tid == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        V s = V(0);
        for (int k = tid; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(partials + k);
        s = block_sum(s, red);
        if (tid == 0) {
            result[0] = SQRT ? sqrt(s) : s;
            *counter = 0u;
        }
    }
}

// general 1 x cols column reduction, block = (32 columns) x (8 row lanes)
template <typename V, Red OP, bool SQRT>
__global__ void __launch_bounds__(256)
    reduce_cols_kernel(int64_t rows, int64_t cols, const V* __restrict__ x, int64_t xs,
                       const V* __restrict__ y, int64_t ys, V* __restrict__ partials,
                       unsigned int* __restrict__ counters, V* __restrict__ result)
{
    __shared__ V tile[8][33];
    __shared__ bool is_last;
    const int cx = threadIdx.x & 31;
    const int ry = threadIdx.x >> 5;
    const int64_t col = blockIdx.y * 32 + cx;
    V acc = V(0);
    if (col < cols) {
        for (int64_t r = blockIdx.x * 8 + ry; r < rows; r += (int64_t)gridDim.x * 8) {
            const V a = x[r * xs + col];
            const V b = OP == Red::dot ? y[r * ys + col] : a;
            acc += a * b;
        }
    }
    tile[ry][cx] = acc;
    __syncthreads();
    if (ry == 0) {
        V s = tile[0][cx];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += tile[k][cx];
        if (col < cols) partials[blockIdx.x * cols + col] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(counters + blockIdx.y, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (ry == 0 && col < cols) {
            V s = V(0);
            for (int k = 0; k < (int)gridDim.x; ++k) s += __ldcg(partials + (int64_t)k * cols + col);
            result[col] = SQRT ? sqrt(s) : s;
        }
        if (threadIdx.x == 0) counters[blockIdx.y] = 0u;
    }
}

template <typename V, Red OP, bool SQRT>
b200_status reduce(b200_ctx* ctx, int64_t rows, int64_t cols, const V* x, int64_t xs, const V* y,
                   int64_t ys, V* result)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(rows >= 0 && cols >= 0, "negative size");
    if (cols == 0) return B200_OK;
    B200_REQUIRE(result != nullptr, "result is null");
    B200_REQUIRE(xs >= cols && (OP != Red::dot || ys >= cols), "stride smaller than cols");
    if (cols == 1) {
        int grid = (int)ceildiv(rows, 512 * 8);
        if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
        if (grid < 1) grid = 1;
        V* partials = (V*)ctx->scratch(sizeof(V) * grid);
        if (!partials) return B200_ERR_ALLOC;
        reduce_col1_kernel<V, OP, SQRT><<<grid, 512, 0, ctx->stream>>>(rows, x, xs, y, ys, partials,
                                                                      ctx->counters, result);
        B200_LAUNCH_CHECK(ctx);
        return B200_OK;
    }
    // column tiles are processed in batches of <= 256 tiles (ticket counters)
    for (int64_t c0 = 0; c0 < cols; c0 += 256 * 32) {
        const int64_t cb = (cols - c0) < 256 * 32 ? (cols - c0) : 256 * 32;
        int gx = (int)ceildiv(rows, 8 * 8);
        const int cap = (int)ceildiv((int64_t)ctx->num_sms * 8, ceildiv(cb, 32));
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        V* partials = (V*)ctx->scratch(sizeof(V) * gx * cb);
        if (!partials) return B200_ERR_ALLOC;
        dim3 grid(gx, (unsigned)ceildiv(cb, 32));
        reduce_cols_kernel<V, OP, SQRT><<<grid, 256, 0, ctx->stream>>>(
            rows, cb, x + c0, xs, OP == Red::dot ? y + c0 : nullptr, ys, partials, ctx->counters,
            result + c0);
        B200_LAUNCH_CHECK(ctx);
    }
    return B200_OK;
}

template <typename V>
b200_status add_scaled(b200_ctx* ctx, int64_t rows, int64_t cols, const V* alpha,
                       int64_t alpha_cols, const V* x, int64_t xs, V* y, int64_t ys, bool sub)
{
    B200_REQUIRE(alpha_cols == 1 || alpha_cols == cols, "alpha must be 1x1 or 1xcols");
    const bool scalar = alpha_cols == 1;
    if (sub) {
        return launch_ew(ctx, rows, cols, [=] __device__(int64_t r, int64_t c) {
            const V a = alpha[scalar ? 0 : c];
            if (!scalar || a != V(0)) y[r * ys + c] -= a * x[r * xs + c];
        });
    }
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t r, int64_t c) {
        const V a = alpha[scalar ? 0 : c];
        if (!scalar || a != V(0)) y[r * ys + c] += a * x[r * xs + c];
    });
}

template <typename V>
b200_status scale(b200_ctx* ctx, int64_t rows, int64_t cols, const V* alpha, int64_t alpha_cols,
                  V* x, int64_t xs, bool inverse)
{
    B200_REQUIRE(alpha_cols == 1 || alpha_cols == cols, "alpha must be 1x1 or 1xcols");
    const bool scalar = alpha_cols == 1;
    if (inverse) {
        return launch_ew(ctx, rows, cols, [=] __device__(int64_t r, int64_t c) {
            x[r * xs + c] /= alpha[scalar ? 0 : c];
        });
    }
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t r, int64_t c) {
        const V a = alpha[scalar ? 0 : c];
        // reference/matrix/dense_kernels.cpp:131-137: a zero 1x1 alpha overwrites (NaN-safe)
        x[r * xs + c] = (scalar && a == V(0)) ? V(0) : x[r * xs + c] * a;
    });
}

}  // namespace dense
}  // namespace b200

extern "C" {

#define B200_DEF_DENSE(V, VT)                                                                  \
    b200_status b200_dense_compute_dot_##V(b200_ctx* ctx, int64_t rows, int64_t cols,          \
                                           const VT* x, int64_t xs, const VT* y, int64_t ys,   \
                                           VT* result)                                         \
    {                                                                                          \
        return b200::dense::reduce<VT, b200::dense::Red::dot, false>(ctx, rows, cols, x, xs,   \
                                                                     y, ys, result);           \
    }                                                                                          \
    b200_status b200_dense_compute_conj_dot_##V(b200_ctx* ctx, int64_t rows, int64_t cols,     \
                                                const VT* x, int64_t xs, const VT* y,          \
                                                int64_t ys, VT* result)                        \
    {                                                                                          \
        return b200::dense::reduce<VT, b200::dense::Red::dot, false>(ctx, rows, cols, x, xs,   \
                                                                     y, ys, result);           \
    }                                                                                          \
    b200_status b200_dense_compute_norm2_##V(b200_ctx* ctx, int64_t rows, int64_t cols,        \
                                             const VT* x, int64_t xs, VT* result)              \
    {                                                                                          \
        return b200::dense::reduce<VT, b200::dense::Red::sqnorm, true>(ctx, rows, cols, x,     \
                                                                       xs, nullptr, 0, result);\
    }                                                                                          \
    b200_status b200_dense_compute_squared_norm2_##V(b200_ctx* ctx, int64_t rows,              \
                                                     int64_t cols, const VT* x, int64_t xs,    \
                                                     VT* result)                               \
    {                                                                                          \
        return b200::dense::reduce<VT, b200::dense::Red::sqnorm, false>(                       \
            ctx, rows, cols, x, xs, nullptr, 0, result);                                       \
    }                                                                                          \
    b200_status b200_dense_add_scaled_##V(b200_ctx* ctx, int64_t rows, int64_t cols,           \
                                          const VT* alpha, int64_t alpha_cols, const VT* x,    \
                                          int64_t xs, VT* y, int64_t ys)                       \
    {                                                                                          \
        return b200::dense::add_scaled<VT>(ctx, rows, cols, alpha, alpha_cols, x, xs, y, ys,   \
                                           false);                                             \
    }                                                                                          \
    b200_status b200_dense_sub_scaled_##V(b200_ctx* ctx, int64_t rows, int64_t cols,           \
                                          const VT* alpha, int64_t alpha_cols, const VT* x,    \
                                          int64_t xs, VT* y, int64_t ys)                       \
    {                                                                                          \
        return b200::dense::add_scaled<VT>(ctx, rows, cols, alpha, alpha_cols, x, xs, y, ys,   \
                                           true);                                              \
    }                                                                                          \
    b200_status b200_dense_scale_##V(b200_ctx* ctx, int64_t rows, int64_t cols,                \
                                     const VT* alpha, int64_t alpha_cols, VT* x, int64_t xs)   \
    {                                                                                          \
        return b200::dense::scale<VT>(ctx, rows, cols, alpha, alpha_cols, x, xs, false);       \
    }                                                                                          \
    b200_status b200_dense_inv_scale_##V(b200_ctx* ctx, int64_t rows, int64_t cols,            \
                                         const VT* alpha, int64_t alpha_cols, VT* x,           \
                                         int64_t xs)                                           \
    {                                                                                          \
        return b200::dense::scale<VT>(ctx, rows, cols, alpha, alpha_cols, x, xs, true);        \
    }                                                                                          \
    b200_status b200_dense_copy_##V(b200_ctx* ctx, int64_t rows, int64_t cols, const VT* in,   \
                                    int64_t is, VT* out, int64_t os)                           \
    {                                                                                          \
        return b200::launch_ew(ctx, rows, cols, [=] __device__(int64_t r, int64_t c) {         \
            out[r * os + c] = in[r * is + c];                                                  \
        });                                                                                    \
    }                                                                                          \
    b200_status b200_dense_fill_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x,          \
                                    int64_t xs, VT value)                                      \
    {                                                                                          \
        return b200::launch_ew(ctx, rows, cols,                                                \
                               [=] __device__(int64_t r, int64_t c) { x[r * xs + c] = value; });\
    }

B200_DEF_DENSE(f64, double)
B200_DEF_DENSE(f32, float)

}  // extern "C"
