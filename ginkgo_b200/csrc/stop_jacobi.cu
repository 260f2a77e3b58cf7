// Stopping-criterion checks and Jacobi preconditioner application.
//
// residual_norm / implicit_residual_norm / set_all_statuses replace
// gko::kernels::cuda::{residual_norm,implicit_residual_norm,set_all_statuses}
// (reference common/cuda_hip/stop/residual_norm_kernels.cpp:33-170,
//  contract reference/stop/residual_norm_kernels.cpp:27-92).  The reference
// launches an init kernel + the check kernel and then does TWO blocking 1-byte
// device->host copies; here one kernel writes both flags into a pinned host
// mailbox and a single stream synchronise hands them back.
//
// Jacobi: scalar apply / invert_diagonal replace
// common/unified/preconditioner/jacobi_kernels.cpp:39-105; block apply replaces
// common/cuda_hip/preconditioner/jacobi_{simple,advanced}_apply_kernels
// (contract reference/preconditioner/jacobi_kernels.cpp:419-592).
#include "elementwise.cuh"
#include "jacobi_precision.cuh"

namespace b200 {
namespace stop {

template <typename V, bool IMPLICIT>
__global__ void residual_norm_kernel(int64_t cols, const V* tau, const V* orig_tau, V goal,
                                     uint8_t id, bool set_finalized, uint8_t* stop,
                                     uint8_t* device_storage, uint8_t* mailbox)
{
    // single CTA; cols is the number of right-hand sides (small)
    __shared__ int not_conv, changed;
    if (threadIdx.x == 0) {
        not_conv = 0;
        changed = 0;
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
        const V t = IMPLICIT ? sqrt(fabs(tau[i])) : tau[i];
        uint8_t s = stop[i];
        if (t <= goal * orig_tau[i]) {
            if (!has_stopped(s)) {
                s |= kConvergedMask | (id & kIdMask);
                if (set_finalized) s |= kFinalizedMask;
                stop[i] = s;
            }
            changed = 1;  // reference sets one_changed even if already stopped
        }
        if (!has_stopped(s)) not_conv = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t all_conv = not_conv ? 0 : 1;
        const uint8_t one_ch = changed ? 1 : 0;
        if (device_storage) {
            device_storage[0] = all_conv;
            device_storage[1] = one_ch;
        }
        mailbox[0] = all_conv;
        mailbox[1] = one_ch;
    }
}

template <typename V, bool IMPLICIT>
b200_status check(b200_ctx* ctx, int64_t cols, const V* tau, const V* orig_tau, V goal, uint8_t id,
                  int32_t set_finalized, uint8_t* stop, uint8_t* device_storage,
                  int32_t* all_converged, int32_t* one_changed)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(cols >= 0, "negative size");
    B200_REQUIRE(all_converged && one_changed, "null output");
    uint8_t* dev_view = nullptr;
    B200_CUDA_CHECK(cudaHostGetDevicePointer((void**)&dev_view, ctx->pinned, 0));
    residual_norm_kernel<V, IMPLICIT><<<1, 256, 0, ctx->stream>>>(
        cols, tau, orig_tau, goal, id, set_finalized != 0, stop, device_storage, dev_view);
    B200_LAUNCH_CHECK(ctx);
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *all_converged = ctx->pinned[0];
    *one_changed = ctx->pinned[1];
    return B200_OK;
}

__global__ void set_all_statuses_kernel(int64_t cols, uint8_t id, bool set_finalized, uint8_t* stop)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= cols) return;
    uint8_t s = stop[i];
    if (!has_stopped(s)) {
        s |= (id & kIdMask);
        if (set_finalized) s |= kFinalizedMask;
        stop[i] = s;
    }
}

}  // namespace stop

namespace jacobi {

// One warp per diagonal block (block sizes <= 32): lane r owns row r of the
// inverted block, b_blk is broadcast by shuffles; columns of the block are
// contiguous across lanes in the interleaved storage (coalesced).  With a
// single right-hand side this is a batch of tiny GEMVs at ~0.5 flop/byte:
// HBM-bound, so no tensor cores (SURVEY.md section 7 "Block-Jacobi layout").
// Summation order per row: inner = 0..bs-1, as the reference's apply_block.
// One warp per storage GROUP (32 / pow2(max_block_size) blocks interleaved so that column c
// of all of them is one contiguous run): lane l holds row l % block_offset of sub-block
// l / block_offset, so every load of a block column is a single coalesced run for the whole
// warp.  All MBS columns are fetched before the first multiply (no chain of dependent
// latencies) and stay in registers for every right-hand side.
template <typename V, typename I, int MBS, bool ADVANCED>
__global__ void __launch_bounds__(256)
    block_apply_kernel(int64_t num_blocks, int64_t block_offset, int64_t group_offset,
                       int32_t group_power, const uint8_t* __restrict__ block_precisions,
                       const I* __restrict__ block_ptrs, const V* __restrict__ blocks,
                       const V* __restrict__ alpha_p,
                       const V* __restrict__ b, int64_t bs_, int64_t num_rhs,
                       const V* __restrict__ beta_p, V* __restrict__ x, int64_t xs)
{
    const int lane = threadIdx.x & 31;
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int group_size = 1 << group_power;
    if (group * group_size >= num_blocks) return;
    const int bo = (int)block_offset;
    const int sub = lane / bo;
    const int r = lane - sub * bo;
    const int64_t k = group * group_size + sub;
    int64_t first = 0;
    int bsz = 0;
    if (sub < group_size && k < num_blocks) {
        first = block_ptrs[k];
        bsz = (int)((int64_t)block_ptrs[k + 1] - first);
    }
    const bool valid = r < bsz;
    const int64_t stride = block_offset << group_power;
    // adaptive precision (jacobi_precision.cuh): the block is stored as the type its precision_reduction
    // byte names, addressed in elements of THAT type from the group's base, and widened on load
    const int kind =
        storage_kind<V>((block_precisions && sub < group_size && k < num_blocks) ? block_precisions[k] : uint8_t(0));
    const void* gbase = blocks + group_offset * group;
    V a[MBS];
#pragma unroll
    for (int inner = 0; inner < MBS; ++inner)
        a[inner] = (valid && inner < bsz) ? load_elem(gbase, lane + inner * stride, kind, V(0)) : V(0);
    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const int src0 = sub * bo;
    for (int64_t j = 0; j < num_rhs; ++j) {
        const V bv = valid ? b[(first + r) * bs_ + j] : V(0);
        V acc = V(0);
        if (ADVANCED && valid && beta != V(0)) acc = x[(first + r) * xs + j] * beta;
#pragma unroll
        for (int inner = 0; inner < MBS; ++inner) {
            const V bi = __shfl_sync(0xffffffffu, bv, (src0 + inner) & 31);
            if (valid && inner < bsz) acc += ADVANCED ? (alpha * a[inner]) * bi : a[inner] * bi;
        }
        if (valid) x[(first + r) * xs + j] = acc;
    }
}

template <typename V, typename I, bool ADVANCED>
b200_status block_apply(b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size,
                        int64_t block_offset, int64_t group_offset, int32_t group_power,
                        const uint8_t* block_precisions, const I* block_ptrs, const V* blocks,
                        const V* alpha, const V* b, int64_t bs, int64_t num_rhs, const V* beta, V* x,
                        int64_t xs)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(max_block_size >= 1 && max_block_size <= 32, "max_block_size must be in [1,32]");
    B200_REQUIRE(block_offset >= 1 && (block_offset << group_power) <= 32 && group_power >= 0,
                 "storage scheme does not fit a warp");
    if (num_blocks <= 0 || num_rhs <= 0) return B200_OK;
    const int64_t groups = ceildiv(num_blocks, int64_t(1) << group_power);
    const unsigned grid = (unsigned)ceildiv(groups * 32, (int64_t)256);
#define B200_BJ(M)                                                                              \
    block_apply_kernel<V, I, M, ADVANCED><<<grid, 256, 0, ctx->stream>>>(                       \
        num_blocks, block_offset, group_offset, group_power, block_precisions, block_ptrs,      \
        blocks, alpha, b, bs, num_rhs, beta, x, xs)
    if (block_offset <= 4)
        B200_BJ(4);
    else if (block_offset <= 8)
        B200_BJ(8);
    else if (block_offset <= 16)
        B200_BJ(16);
    else
        B200_BJ(32);
#undef B200_BJ
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

}  // namespace jacobi
}  // namespace b200

extern "C" {

b200_status b200_set_all_statuses(b200_ctx* ctx, int64_t cols, uint8_t stopping_id,
                                  int32_t set_finalized, uint8_t* stop_status)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    if (cols <= 0) return B200_OK;
    b200::stop::set_all_statuses_kernel<<<(unsigned)b200::ceildiv(cols, 256), 256, 0, ctx->stream>>>(
        cols, stopping_id, set_finalized != 0, stop_status);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

#define B200_DEF_STOP_JACOBI(V, VT)                                                            \
    b200_status b200_residual_norm_##V(b200_ctx* ctx, int64_t cols, const VT* tau,             \
                                       const VT* orig_tau, VT goal, uint8_t id,                \
                                       int32_t set_finalized, uint8_t* stop,                   \
                                       uint8_t* device_storage, int32_t* all_converged,        \
                                       int32_t* one_changed)                                   \
    {                                                                                          \
        return b200::stop::check<VT, false>(ctx, cols, tau, orig_tau, goal, id, set_finalized, \
                                            stop, device_storage, all_converged,               \
                                            one_changed);                                      \
    }                                                                                          \
    b200_status b200_implicit_residual_norm_##V(                                               \
        b200_ctx* ctx, int64_t cols, const VT* tau, const VT* orig_tau, VT goal, uint8_t id,   \
        int32_t set_finalized, uint8_t* stop, uint8_t* device_storage,                         \
        int32_t* all_converged, int32_t* one_changed)                                          \
    {                                                                                          \
        return b200::stop::check<VT, true>(ctx, cols, tau, orig_tau, goal, id, set_finalized,  \
                                           stop, device_storage, all_converged, one_changed);  \
    }                                                                                          \
    b200_status b200_jacobi_invert_diagonal_##V(b200_ctx* ctx, int64_t n, const VT* diag,      \
                                                VT* inv_diag)                                  \
    {                                                                                          \
        return b200::launch_ew(ctx, n, 1, [=] __device__(int64_t i, int64_t) {                 \
            const VT d = diag[i] == VT(0) ? VT(1) : diag[i];                                   \
            inv_diag[i] = VT(1) / d;                                                           \
        });                                                                                    \
    }                                                                                          \
    b200_status b200_jacobi_simple_scalar_apply_##V(b200_ctx* ctx, int64_t rows, int64_t cols, \
                                                    const VT* inv_diag, const VT* b,           \
                                                    int64_t bs, VT* x, int64_t xs)             \
    {                                                                                          \
        return b200::launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {         \
            x[i * xs + j] = b[i * bs + j] * inv_diag[i];                                       \
        });                                                                                    \
    }                                                                                          \
    b200_status b200_jacobi_scalar_apply_##V(b200_ctx* ctx, int64_t rows, int64_t cols,        \
                                             const VT* inv_diag, const VT* alpha,              \
                                             const VT* b, int64_t bs, const VT* beta, VT* x,   \
                                             int64_t xs)                                       \
    {                                                                                          \
        return b200::launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {         \
            x[i * xs + j] = beta[0] * x[i * xs + j] + alpha[0] * b[i * bs + j] * inv_diag[i];  \
        });                                                                                    \
    }

B200_DEF_STOP_JACOBI(f64, double)
B200_DEF_STOP_JACOBI(f32, float)

#define B200_DEF_JACOBI_BLOCK(V, VT, I, IT)                                                    \
    b200_status b200_jacobi_simple_apply_##V##_##I(                                            \
        b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,       \
        int64_t group_offset, int32_t group_power, const IT* block_pointers,                   \
        const VT* blocks, const VT* b, int64_t bs, int64_t num_rhs, VT* x, int64_t xs)         \
    {                                                                                          \
        return b200::jacobi::block_apply<VT, IT, false>(                                       \
            ctx, num_blocks, max_block_size, block_offset, group_offset, group_power, nullptr, \
            block_pointers, blocks, nullptr, b, bs, num_rhs, nullptr, x, xs);                  \
    }                                                                                          \
    b200_status b200_jacobi_simple_apply_adaptive_##V##_##I(                                   \
        b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,       \
        int64_t group_offset, int32_t group_power, const uint8_t* block_precisions,            \
        const IT* block_pointers, const VT* blocks, const VT* b, int64_t bs, int64_t num_rhs,  \
        VT* x, int64_t xs)                                                                     \
    {                                                                                          \
        return b200::jacobi::block_apply<VT, IT, false>(                                       \
            ctx, num_blocks, max_block_size, block_offset, group_offset, group_power,          \
            block_precisions, block_pointers, blocks, nullptr, b, bs, num_rhs, nullptr, x,     \
            xs);                                                                               \
    }                                                                                          \
    b200_status b200_jacobi_apply_adaptive_##V##_##I(                                          \
        b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,       \
        int64_t group_offset, int32_t group_power, const uint8_t* block_precisions,            \
        const IT* block_pointers, const VT* blocks, const VT* alpha, const VT* b, int64_t bs,  \
        int64_t num_rhs, const VT* beta, VT* x, int64_t xs)                                    \
    {                                                                                          \
        return b200::jacobi::block_apply<VT, IT, true>(                                        \
            ctx, num_blocks, max_block_size, block_offset, group_offset, group_power,          \
            block_precisions, block_pointers, blocks, alpha, b, bs, num_rhs, beta, x, xs);     \
    }                                                                                          \
    b200_status b200_jacobi_apply_##V##_##I(                                                   \
        b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,       \
        int64_t group_offset, int32_t group_power, const IT* block_pointers,                   \
        const VT* blocks, const VT* alpha, const VT* b, int64_t bs, int64_t num_rhs,           \
        const VT* beta, VT* x, int64_t xs)                                                     \
    {                                                                                          \
        return b200::jacobi::block_apply<VT, IT, true>(                                        \
            ctx, num_blocks, max_block_size, block_offset, group_offset, group_power, nullptr, \
            block_pointers, blocks, alpha, b, bs, num_rhs, beta, x, xs);                       \
    }

B200_DEF_JACOBI_BLOCK(f64, double, i32, int32_t)
B200_DEF_JACOBI_BLOCK(f64, double, i64, int64_t)
B200_DEF_JACOBI_BLOCK(f32, float, i32, int32_t)
B200_DEF_JACOBI_BLOCK(f32, float, i64, int64_t)

}  // extern "C"
