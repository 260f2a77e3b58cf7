// CSR SpMV for sm_100a:  c = A b   and   c = alpha A b + beta c.
//
// Replaces gko::kernels::cuda::csr::{spmv, advanced_spmv}
// (reference common/cuda_hip/matrix/csr_kernels.template.cpp:2353-2468); the
// arithmetic contract is the reference executor's
// (reference/matrix/csr_kernels.cpp:47-118): per row, products accumulated left
// to right, for advanced_spmv starting from beta*c (never reading c if beta==0)
// and adding (alpha*val)*b.
//
// Kernel design ("row-segmented slab" kernel, single right-hand side):
//   * The merge-path coordinate row + row_ptrs[row] is cut into equal tiles of
//     TILE items by a partition (the cached b200_csr_plan, the analogue of the
//     reference's `srow`), so every CTA owns whole rows and <= TILE nonzeros
//     (plus at most one over-long last row).
//   * Phase 1: the CTA streams its slab of col_idxs/values with 256-bit
//     coalesced loads (LDG.E.NA.EFL2.256: no L1 allocation, L2 evict-first),
//     one nonzero per lane-slot, so the HBM stream is perfectly load balanced
//     regardless of row lengths; every nonzero gathers b[col] (L2 evict-last)
//     and parks val*b in shared memory.
//   * Phase 2: LANES threads per row add the row's products from shared
//     memory.  With LANES == 1 the sum is strictly left to right, i.e.
//     bit-identical to the reference executor; LANES > 1 (long rows) uses a
//     fixed shuffle tree.  No floating-point atomics anywhere: the result is
//     deterministic.
// Multiple right-hand sides use a simple thread-per-(row,rhs) kernel with the
// reference's summation order.
#include "common.cuh"

namespace b200 {
namespace csr {

constexpr int kThreads = 256;
constexpr int kItemsPerThread = 8;
constexpr int kTile = kThreads * kItemsPerThread;  // merge items per CTA
// products buffer: all rows of a tile but the last hold < kTile nonzeros, so a
// last row of up to kTile nonzeros still fits and keeps its left-to-right sum
constexpr int kChunk = 2 * kTile + 8;

template <typename I>
__global__ void plan_kernel(const I* __restrict__ row_ptrs, int64_t num_rows, int64_t num_tiles,
                            int64_t* __restrict__ tile_rows)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    const int64_t d = t * (int64_t)kTile;
    int64_t lo = 0, hi = num_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (mid + (int64_t)row_ptrs[mid] >= d)
            hi = mid;
        else
            lo = mid + 1;
    }
    tile_rows[t] = lo;
}

template <typename V, typename I, int LANES, bool ADVANCED, bool VEC>
__global__ void __launch_bounds__(kThreads, 3)
    slab_kernel(const int64_t* __restrict__ tile_rows, int64_t nnz,
                const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                const V* __restrict__ values, const V* __restrict__ alpha_p,
                const V* __restrict__ b, int64_t b_stride, const V* __restrict__ beta_p,
                V* __restrict__ c, int64_t c_stride)
{
    __shared__ __align__(16) V prod[kChunk];
    __shared__ V red[32];

    const int tid = threadIdx.x;
    const int64_t r0 = tile_rows[blockIdx.x];
    const int64_t r1 = tile_rows[blockIdx.x + 1];
    if (r0 >= r1) return;
    const int64_t p0 = row_ptrs[r0];
    const int64_t p1 = row_ptrs[r1];
    const int64_t a0 = p0 & ~int64_t(7);

    V alpha = V(1), beta = V(0);
    if (ADVANCED) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const uint64_t pol_last = policy_evict_last();
    const uint64_t pol_first = policy_evict_first();

    // an over-long last row (does not fit the products buffer) is summed by the
    // whole CTA straight from global memory
    const bool long_last = (p1 - a0) > kChunk;
    const int64_t rl = r1 - 1;
    const int64_t sl = long_last ? (int64_t)row_ptrs[rl] : p1;
    const int64_t pend = long_last ? sl : p1;  // products are staged for [p0, pend)
    const int64_t rows_end = long_last ? rl : r1;

    // ---- phase 1: stream slab, gather b, stage products ----------------------
#pragma unroll 2
    for (int64_t base = a0 + (int64_t)tid * 8; base < pend; base += (int64_t)kThreads * 8) {
        I cols[8];
        V vals[8];
        if (VEC && base + 8 <= nnz) {
            ld_stream_x8(col_idxs + base, cols);
            ld_stream_x8(values + base, vals);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t idx = base + k;
                const bool ok = idx < nnz;
                cols[k] = ok ? ld_stream(col_idxs + idx, pol_first) : I(0);
                vals[k] = ok ? ld_stream(values + idx, pol_first) : V(0);
            }
        }
        V xs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t idx = base + k;
            const bool ok = (idx >= p0) && (idx < pend);
            xs[k] = ok ? ld_gather(b + (int64_t)cols[k] * b_stride, pol_last) : V(0);
        }
        V* dst = prod + (base - a0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            dst[k] = ADVANCED ? (alpha * vals[k]) * xs[k] : vals[k] * xs[k];
        }
    }
    __syncthreads();

    // ---- phase 2: per-row sums from shared memory ----------------------------
    constexpr int kRowsPerPass = kThreads / LANES;
    const int sub = tid % LANES;
    const int64_t nrows = rows_end - r0;
    const int64_t passes = (nrows + kRowsPerPass - 1) / kRowsPerPass;
    for (int64_t ps = 0; ps < passes; ++ps) {
        const int64_t r = r0 + ps * kRowsPerPass + tid / LANES;
        const bool rv = r < rows_end;
        int64_t s = 0, e = 0;
        if (rv) {
            s = row_ptrs[r];
            e = row_ptrs[r + 1];
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
        if (rv && sub == 0) c[r * c_stride] = acc;
    }

    // ---- over-long last row --------------------------------------------------
    if (long_last) {
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
        }
    }
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

}  // namespace csr
}  // namespace b200

struct b200_csr_plan {
    int64_t num_rows = 0;
    int64_t nnz = 0;
    int64_t num_tiles = 0;
    int64_t* tile_rows = nullptr;  // device, num_tiles + 1
    int lanes = 1;
    int device = 0;
};

namespace b200 {
namespace csr {

template <typename I>
b200_status fill_plan(b200_ctx* ctx, int64_t num_rows, int64_t nnz, const I* row_ptrs,
                      int64_t num_tiles, int64_t* tile_rows)
{
    const int block = 256;
    const int grid = (int)ceildiv(num_tiles + 1, block);
    plan_kernel<I><<<grid, block, 0, ctx->stream>>>(row_ptrs, num_rows, num_tiles, tile_rows);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename V, typename I, bool ADVANCED>
b200_status launch_slab(b200_ctx* ctx, int lanes, bool vec, int64_t num_tiles,
                        const int64_t* tile_rows, int64_t nnz, const I* row_ptrs,
                        const I* col_idxs, const V* values, const V* alpha, const V* b,
                        int64_t b_stride, const V* beta, V* c, int64_t c_stride)
{
#define B200_SLAB(L, VECF)                                                                    \
    slab_kernel<V, I, L, ADVANCED, VECF><<<(unsigned)num_tiles, kThreads, 0, ctx->stream>>>(  \
        tile_rows, nnz, row_ptrs, col_idxs, values, alpha, b, b_stride, beta, c, c_stride)
#define B200_SLAB_L(L)       \
    do {                     \
        if (vec)             \
            B200_SLAB(L, true);  \
        else                 \
            B200_SLAB(L, false); \
    } while (0)
    switch (lanes) {
    case 1: B200_SLAB_L(1); break;
    case 2: B200_SLAB_L(2); break;
    case 4: B200_SLAB_L(4); break;
    case 8: B200_SLAB_L(8); break;
    case 16: B200_SLAB_L(16); break;
    default: B200_SLAB_L(32); break;
    }
#undef B200_SLAB_L
#undef B200_SLAB
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename V, typename I, bool ADVANCED>
b200_status spmv_impl(b200_ctx* ctx, const b200_csr_plan* plan, int64_t num_rows,
                      int64_t num_cols, int64_t nnz, const I* row_ptrs, const I* col_idxs,
                      const V* values, const V* alpha, const V* b, int64_t b_stride,
                      int64_t num_rhs, const V* beta, V* c, int64_t c_stride)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(num_rows >= 0 && num_cols >= 0 && nnz >= 0 && num_rhs >= 0, "negative size");
    B200_REQUIRE(b_stride >= num_rhs && c_stride >= num_rhs, "stride smaller than num_rhs");
    if (num_rows == 0 || num_rhs == 0) return B200_OK;
    B200_REQUIRE(row_ptrs && c, "null pointer");
    B200_REQUIRE(nnz == 0 || (col_idxs && values && b), "null pointer");
    if (ADVANCED) B200_REQUIRE(alpha && beta, "alpha/beta must be device pointers");

    if (num_rhs > 1) {
        const int grid = grid_for(num_rows * num_rhs, 256, ctx->num_sms, 16);
        multi_rhs_kernel<V, I, ADVANCED><<<grid, 256, 0, ctx->stream>>>(
            num_rows, num_rhs, row_ptrs, col_idxs, values, alpha, b, b_stride, beta, c, c_stride);
        B200_LAUNCH_CHECK(ctx);
        return B200_OK;
    }

    const int64_t num_tiles = ceildiv(num_rows + nnz, kTile);
    const int64_t* tile_rows = nullptr;
    int lanes;
    if (plan) {
        B200_REQUIRE(plan->num_rows == num_rows && plan->nnz == nnz,
                     "plan does not match the matrix");
        tile_rows = plan->tile_rows;
        lanes = plan->lanes;
    } else {
        int64_t* tr = (int64_t*)ctx->scratch((num_tiles + 1) * sizeof(int64_t));
        if (!tr) {
            set_error("scratch allocation failed");
            return B200_ERR_ALLOC;
        }
        b200_status st = fill_plan<I>(ctx, num_rows, nnz, row_ptrs, num_tiles, tr);
        if (st != B200_OK) return st;
        tile_rows = tr;
        lanes = pick_lanes(num_rows, nnz);
    }
    const bool vec = (((uintptr_t)col_idxs | (uintptr_t)values) & 31u) == 0;
    return launch_slab<V, I, ADVANCED>(ctx, lanes, vec, num_tiles, tile_rows, nnz, row_ptrs,
                                       col_idxs, values, alpha, b, b_stride, beta, c, c_stride);
}

template <typename I>
b200_status plan_create(b200_ctx* ctx, int64_t num_rows, int64_t nnz, const I* row_ptrs,
                        b200_csr_plan** out)
{
    B200_REQUIRE(ctx && out, "null argument");
    B200_REQUIRE(num_rows >= 0 && nnz >= 0, "negative size");
    b200_csr_plan* p = new b200_csr_plan();
    p->num_rows = num_rows;
    p->nnz = nnz;
    p->num_tiles = ceildiv(num_rows + nnz, kTile);
    p->lanes = pick_lanes(num_rows, nnz);
    p->device = ctx->device;
    cudaError_t e = cudaMalloc((void**)&p->tile_rows, (p->num_tiles + 1) * sizeof(int64_t));
    if (e != cudaSuccess) {
        delete p;
        set_error("cudaMalloc failed for csr plan: %s", cudaGetErrorString(e));
        return B200_ERR_ALLOC;
    }
    if (num_rows > 0) {
        b200_status st = fill_plan<I>(ctx, num_rows, nnz, row_ptrs, p->num_tiles, p->tile_rows);
        if (st != B200_OK) {
            cudaFree(p->tile_rows);
            delete p;
            return st;
        }
    }
    *out = p;
    return B200_OK;
}

}  // namespace csr
}  // namespace b200

extern "C" {

void b200_csr_plan_destroy(b200_csr_plan* plan)
{
    if (!plan) return;
    cudaSetDevice(plan->device);
    cudaDeviceSynchronize();
    cudaFree(plan->tile_rows);
    delete plan;
}

#define B200_DEF_CSR(V, VT, I, IT)                                                             \
    b200_status b200_csr_plan_create_##V##_##I(b200_ctx* ctx, int64_t num_rows, int64_t nnz,   \
                                               const IT* row_ptrs, b200_csr_plan** out)        \
    {                                                                                          \
        return b200::csr::plan_create<IT>(ctx, num_rows, nnz, row_ptrs, out);                  \
    }                                                                                          \
    b200_status b200_csr_spmv_##V##_##I(                                                       \
        b200_ctx* ctx, const b200_csr_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_ptrs, const IT* col_idxs, const VT* values, const VT* b,    \
        int64_t b_stride, int64_t num_rhs, VT* c, int64_t c_stride)                            \
    {                                                                                          \
        return b200::csr::spmv_impl<VT, IT, false>(ctx, plan, num_rows, num_cols, nnz,         \
                                                   row_ptrs, col_idxs, values, nullptr, b,     \
                                                   b_stride, num_rhs, nullptr, c, c_stride);   \
    }                                                                                          \
    b200_status b200_csr_advanced_spmv_##V##_##I(                                              \
        b200_ctx* ctx, const b200_csr_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_ptrs, const IT* col_idxs, const VT* values,                 \
        const VT* alpha, const VT* b, int64_t b_stride, int64_t num_rhs, const VT* beta,       \
        VT* c, int64_t c_stride)                                                               \
    {                                                                                          \
        return b200::csr::spmv_impl<VT, IT, true>(ctx, plan, num_rows, num_cols, nnz,          \
                                                  row_ptrs, col_idxs, values, alpha, b,        \
                                                  b_stride, num_rhs, beta, c, c_stride);       \
    }

B200_DEF_CSR(f64, double, i32, int32_t)
B200_DEF_CSR(f64, double, i64, int64_t)
B200_DEF_CSR(f32, float, i32, int32_t)
B200_DEF_CSR(f32, float, i64, int64_t)

}  // extern "C"

// ===========================================================================
// COO on top of the CSR kernels.  A row-sorted COO is a CSR whose row pointers
// are implicit; convert_idxs_to_ptrs (reference/components/
// format_conversion_kernels.cpp, `convert_idxs_to_ptrs`) writes them down once
// (cached in b200_coo_plan) and every apply then runs the slab kernel:
//   spmv2:           c += A b        == advanced(alpha=1, beta=1)
//   advanced_spmv2:  c += alpha A b  == advanced(alpha,   beta=1)
// The slab kernel starts each row from c*beta and adds the products left to
// right, which is exactly the order of the reference's `c(row) += val*b(col)`
// loop (reference/matrix/coo_kernels.cpp:59-97), so no atomics are needed and
// the result is deterministic -- the reference CUDA path uses atomic_add
// (common/cuda_hip/matrix/coo_kernels.cpp:62-231).
// ===========================================================================
struct b200_coo_plan {
    int64_t num_rows = 0;
    int64_t nnz = 0;
    void* row_ptrs = nullptr;  // device, (num_rows + 1) index_type
    b200_csr_plan* csr = nullptr;
    void* ones = nullptr;  // device: {1.0, 1.0} in the plan's value type
    int device = 0;
};

namespace b200 {
namespace coo {

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
    if (i == nnz - 1) {
        for (int64_t r = cur + 1; r <= num_rows; ++r) ptrs[r] = (I)nnz;
    }
}

template <typename I>
b200_status idxs_to_ptrs(b200_ctx* ctx, const I* idxs, int64_t nnz, int64_t num_rows, I* ptrs)
{
    const int64_t work = nnz == 0 ? num_rows + 1 : nnz;
    idxs_to_ptrs_kernel<I><<<(unsigned)ceildiv(work, 256), 256, 0, ctx->stream>>>(idxs, nnz, num_rows,
                                                                                  ptrs);
    B200_LAUNCH_CHECK(ctx);
    return B200_OK;
}

template <typename V>
__global__ void set_ones_kernel(V* p)
{
    p[0] = V(1);
    p[1] = V(1);
}

template <typename V, typename I>
b200_status plan_create(b200_ctx* ctx, int64_t num_rows, int64_t nnz, const I* row_idxs,
                        b200_coo_plan** out)
{
    B200_REQUIRE(ctx && out, "null argument");
    B200_REQUIRE(num_rows >= 0 && nnz >= 0, "negative size");
    b200_coo_plan* p = new b200_coo_plan();
    p->num_rows = num_rows;
    p->nnz = nnz;
    p->device = ctx->device;
    if (cudaMalloc(&p->row_ptrs, (num_rows + 1) * sizeof(I)) != cudaSuccess ||
        cudaMalloc(&p->ones, 2 * sizeof(V)) != cudaSuccess) {
        set_error("cudaMalloc failed for coo plan");
        cudaFree(p->row_ptrs);
        delete p;
        return B200_ERR_ALLOC;
    }
    b200_status st = idxs_to_ptrs<I>(ctx, row_idxs, nnz, num_rows, (I*)p->row_ptrs);
    if (st == B200_OK) {
        set_ones_kernel<V><<<1, 1, 0, ctx->stream>>>((V*)p->ones);
        ctx->launches++;
        st = csr::plan_create<I>(ctx, num_rows, nnz, (const I*)p->row_ptrs, &p->csr);
    }
    if (st != B200_OK) {
        cudaFree(p->row_ptrs);
        cudaFree(p->ones);
        delete p;
        return st;
    }
    *out = p;
    return B200_OK;
}

// mode: 0 spmv, 1 advanced_spmv, 2 spmv2, 3 advanced_spmv2
template <typename V, typename I>
b200_status apply(b200_ctx* ctx, const b200_coo_plan* plan, int mode, int64_t num_rows,
                  int64_t num_cols, int64_t nnz, const I* row_idxs, const I* col_idxs,
                  const V* values, const V* alpha, const V* b, int64_t b_stride, int64_t num_rhs,
                  const V* beta, V* c, int64_t c_stride)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(num_rows >= 0 && nnz >= 0, "negative size");
    if (num_rows == 0 || num_rhs == 0) return B200_OK;
    const I* row_ptrs;
    const b200_csr_plan* cplan = nullptr;
    const V* ones;
    if (plan) {
        B200_REQUIRE(plan->num_rows == num_rows && plan->nnz == nnz, "plan does not match");
        row_ptrs = (const I*)plan->row_ptrs;
        cplan = plan->csr;
        ones = (const V*)plan->ones;
    } else {
        // one scratch block: [ones (16 B) | row_ptrs | csr tiles]; the CSR call below
        // must not re-use ctx->scratch, so the tile array is carved out here too
        const int64_t num_tiles = ceildiv(num_rows + nnz, csr::kTile);
        const size_t off_ptrs = 16;
        const size_t off_tiles = (off_ptrs + (num_rows + 1) * sizeof(I) + 15) & ~size_t(15);
        const size_t total = off_tiles + (num_tiles + 1) * sizeof(int64_t);
        char* base = (char*)ctx->scratch(total);
        if (!base) return B200_ERR_ALLOC;
        V* o = (V*)base;
        I* rp = (I*)(base + off_ptrs);
        int64_t* tiles = (int64_t*)(base + off_tiles);
        set_ones_kernel<V><<<1, 1, 0, ctx->stream>>>(o);
        ctx->launches++;
        b200_status st = idxs_to_ptrs<I>(ctx, row_idxs, nnz, num_rows, rp);
        if (st != B200_OK) return st;
        if (num_rhs == 1) {
            st = csr::fill_plan<I>(ctx, num_rows, nnz, rp, num_tiles, tiles);
            if (st != B200_OK) return st;
            const bool vec = (((uintptr_t)col_idxs | (uintptr_t)values) & 31u) == 0;
            const int lanes = csr::pick_lanes(num_rows, nnz);
            if (mode == 0)
                return csr::launch_slab<V, I, false>(ctx, lanes, vec, num_tiles, tiles, nnz, rp,
                                                     col_idxs, values, nullptr, b, b_stride,
                                                     nullptr, c, c_stride);
            return csr::launch_slab<V, I, true>(ctx, lanes, vec, num_tiles, tiles, nnz, rp,
                                                col_idxs, values, (mode == 2) ? o : alpha, b,
                                                b_stride, (mode == 1) ? beta : o + 1, c, c_stride);
        }
        row_ptrs = rp;
        ones = o;
        // multi-rhs CSR path does not touch ctx->scratch
    }
    switch (mode) {
    case 0:
        return csr::spmv_impl<V, I, false>(ctx, cplan, num_rows, num_cols, nnz, row_ptrs, col_idxs,
                                           values, nullptr, b, b_stride, num_rhs, nullptr, c,
                                           c_stride);
    case 1:
        return csr::spmv_impl<V, I, true>(ctx, cplan, num_rows, num_cols, nnz, row_ptrs, col_idxs,
                                          values, alpha, b, b_stride, num_rhs, beta, c, c_stride);
    case 2:
        return csr::spmv_impl<V, I, true>(ctx, cplan, num_rows, num_cols, nnz, row_ptrs, col_idxs,
                                          values, ones, b, b_stride, num_rhs, ones + 1, c,
                                          c_stride);
    default:
        return csr::spmv_impl<V, I, true>(ctx, cplan, num_rows, num_cols, nnz, row_ptrs, col_idxs,
                                          values, alpha, b, b_stride, num_rhs, ones + 1, c,
                                          c_stride);
    }
}

}  // namespace coo
}  // namespace b200

extern "C" {

void b200_coo_plan_destroy(b200_coo_plan* plan)
{
    if (!plan) return;
    b200_csr_plan_destroy(plan->csr);
    cudaSetDevice(plan->device);
    cudaFree(plan->row_ptrs);
    cudaFree(plan->ones);
    delete plan;
}

#define B200_DEF_COO(V, VT, I, IT)                                                             \
    b200_status b200_coo_plan_create_##V##_##I(b200_ctx* ctx, int64_t num_rows, int64_t nnz,   \
                                               const IT* row_idxs, b200_coo_plan** out)        \
    {                                                                                          \
        return b200::coo::plan_create<VT, IT>(ctx, num_rows, nnz, row_idxs, out);              \
    }                                                                                          \
    b200_status b200_coo_spmv_##V##_##I(                                                       \
        b200_ctx* ctx, const b200_coo_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_idxs, const IT* col_idxs, const VT* values, const VT* b,    \
        int64_t b_stride, int64_t num_rhs, VT* c, int64_t c_stride)                            \
    {                                                                                          \
        return b200::coo::apply<VT, IT>(ctx, plan, 0, num_rows, num_cols, nnz, row_idxs,       \
                                        col_idxs, values, nullptr, b, b_stride, num_rhs,       \
                                        nullptr, c, c_stride);                                 \
    }                                                                                          \
    b200_status b200_coo_advanced_spmv_##V##_##I(                                              \
        b200_ctx* ctx, const b200_coo_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_idxs, const IT* col_idxs, const VT* values,                 \
        const VT* alpha, const VT* b, int64_t b_stride, int64_t num_rhs, const VT* beta,       \
        VT* c, int64_t c_stride)                                                               \
    {                                                                                          \
        return b200::coo::apply<VT, IT>(ctx, plan, 1, num_rows, num_cols, nnz, row_idxs,       \
                                        col_idxs, values, alpha, b, b_stride, num_rhs, beta,   \
                                        c, c_stride);                                          \
    }                                                                                          \
    b200_status b200_coo_spmv2_##V##_##I(                                                      \
        b200_ctx* ctx, const b200_coo_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_idxs, const IT* col_idxs, const VT* values, const VT* b,    \
        int64_t b_stride, int64_t num_rhs, VT* c, int64_t c_stride)                            \
    {                                                                                          \
        return b200::coo::apply<VT, IT>(ctx, plan, 2, num_rows, num_cols, nnz, row_idxs,       \
                                        col_idxs, values, nullptr, b, b_stride, num_rhs,       \
                                        nullptr, c, c_stride);                                 \
    }                                                                                          \
    b200_status b200_coo_advanced_spmv2_##V##_##I(                                             \
        b200_ctx* ctx, const b200_coo_plan* plan, int64_t num_rows, int64_t num_cols,          \
        int64_t nnz, const IT* row_idxs, const IT* col_idxs, const VT* values,                 \
        const VT* alpha, const VT* b, int64_t b_stride, int64_t num_rhs, VT* c,                \
        int64_t c_stride)                                                                      \
    {                                                                                          \
        return b200::coo::apply<VT, IT>(ctx, plan, 3, num_rows, num_cols, nnz, row_idxs,       \
                                        col_idxs, values, alpha, b, b_stride, num_rhs,         \
                                        nullptr, c, c_stride);                                 \
    }

B200_DEF_COO(f64, double, i32, int32_t)
B200_DEF_COO(f64, double, i64, int64_t)
B200_DEF_COO(f32, float, i32, int32_t)
B200_DEF_COO(f32, float, i64, int64_t)

}  // extern "C"
