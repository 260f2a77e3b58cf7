// BiCG and what it needs besides the kernels the other solvers already
// have: the transposed system matrix and the transposed block-Jacobi preconditioner.
//   bicg::initialize / step_1 / step_2   reference/solver/bicg_kernels.cpp:25-118
//   csr::transpose                        reference/matrix/csr_kernels.cpp:694-731
//   jacobi::transpose_jacobi              reference/preconditioner/jacobi_kernels.cpp:208-218, :597-627
// all bit-exact.
//
// csr::transpose: the reference counts the columns, prefix-sums them and then walks the rows in
// order, so inside a row of the transpose the entries are ordered by (original row, position) --
// a STABLE sort of the entries by column.  Here: least-significant-digit radix sort of
// (column, entry index) with 8-bit digits.  Each pass is three steps over chunks of 2048
// consecutive entries: per-chunk digit histogram (one thread per chunk), one exclusive scan over
// the digit-major / chunk-minor counts, per-chunk stable scatter.  ceil(bits(num_cols) / 8)
// passes; values and row indices are gathered once at the end through the sorted permutation.
// (Set-up work for BiCG, not a hot kernel: the chunk walks are sequential per thread.)
// Element-wise lambdas + scans only, so a copy of this file compiles for the host and is checked
// there without a GPU (tests/test_transpose_bicg_cpu.py).
#include "elementwise.cuh"
#include "scan.cuh"

namespace b200 {
namespace steps {

template <typename V>
b200_status bicg_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs, V* r,
                            int64_t rs, V* z, int64_t zs, V* p, int64_t ps, V* q, int64_t qs, V* prev_rho,
                            V* rho, V* r2, int64_t r2s, V* z2, int64_t z2s, V* p2, int64_t p2s, V* q2,
                            int64_t q2s, uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            rho[j] = V(0);
            prev_rho[j] = V(1);
            stop[j] = 0;
        } else {
            const V v = b[i * bs + j];
            r[i * rs + j] = v;
            r2[i * r2s + j] = v;
            z[i * zs + j] = V(0);
            p[i * ps + j] = V(0);
            q[i * qs + j] = V(0);
            z2[i * z2s + j] = V(0);
            p2[i * p2s + j] = V(0);
            q2[i * q2s + j] = V(0);
        }
    });
}

template <typename V>
b200_status bicg_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, V* p, int64_t ps, const V* z, int64_t zs,
                        V* p2, int64_t p2s, const V* z2, int64_t z2s, const V* rho, const V* prev_rho,
                        const uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        if (prev_rho[j] == V(0)) {
            p[i * ps + j] = z[i * zs + j];
            p2[i * p2s + j] = z2[i * z2s + j];
        } else {
            const V tmp = rho[j] / prev_rho[j];
            p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
            p2[i * p2s + j] = z2[i * z2s + j] + tmp * p2[i * p2s + j];
        }
    });
}

template <typename V>
b200_status bicg_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs, V* r2,
                        int64_t r2s, const V* p, int64_t ps, const V* q, int64_t qs, const V* q2,
                        int64_t q2s, const V* beta, const V* rho, const uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        if (beta[j] != V(0)) {
            const V tmp = rho[j] / beta[j];
            x[i * xs + j] += tmp * p[i * ps + j];
            r[i * rs + j] -= tmp * q[i * qs + j];
            r2[i * r2s + j] -= tmp * q2[i * q2s + j];
        }
    });
}

}  // namespace steps

namespace transpose {

constexpr int kChunk = 2048;
constexpr int kDigits = 256;

inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

template <typename V, typename I>
b200_status csr_transpose(b200_ctx* ctx, int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_ptrs,
                          const I* col_idxs, const V* values, I* t_row_ptrs, I* t_col_idxs, V* t_values)
{
    B200_REQUIRE(ctx && t_row_ptrs, "null argument");
    B200_REQUIRE(num_rows >= 0 && num_cols >= 0 && nnz >= 0, "negative size");
    if (nnz == 0)
        return launch_ew(ctx, num_cols + 1, 1, [=] __device__(int64_t c, int64_t) { t_row_ptrs[c] = I(0); });
    B200_REQUIRE(row_ptrs && col_idxs && values && t_col_idxs && t_values, "null argument");
    B200_REQUIRE(num_cols > 0 && num_rows > 0, "entries in an empty matrix");
    int bits = 1;
    while (bits < 63 && ((int64_t)1 << bits) < num_cols) ++bits;
    const int passes = (bits + 7) / 8;
    const int64_t chunks = ceildiv(nnz, (int64_t)kChunk);
    const int64_t ncount = chunks * kDigits;
    const size_t o_key = al256(sizeof(I) * nnz), o_cnt = al256(sizeof(int64_t) * (ncount + 1));
    char* base = (char*)ctx->scratch(4 * o_key + 2 * o_cnt + al256(sizeof(int64_t) * scan::num_tiles(ncount + 1)));
    if (!base) return B200_ERR_ALLOC;
    I* key_a = (I*)base;
    I* key_b = (I*)(base + o_key);
    I* idx_a = (I*)(base + 2 * o_key);
    I* idx_b = (I*)(base + 3 * o_key);
    int64_t* counts = (int64_t*)(base + 4 * o_key);
    int64_t* offs = (int64_t*)(base + 4 * o_key + o_cnt);
    int64_t* sums = (int64_t*)(base + 4 * o_key + 2 * o_cnt);
    b200_status st = launch_ew(ctx, nnz, 1, [=] __device__(int64_t k, int64_t) {
        key_a[k] = col_idxs[k];
        idx_a[k] = (I)k;
    });
    if (st != B200_OK) return st;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 8 * pass;
        const I* key = key_a;
        const I* idx = idx_a;
        I* key_o = key_b;
        I* idx_o = idx_b;
        st = launch_ew(ctx, chunks, 1, [=] __device__(int64_t c, int64_t) {
            int cnt[kDigits];
            for (int d = 0; d < kDigits; ++d) cnt[d] = 0;
            const int64_t lo = c * kChunk;
            const int64_t hi = lo + kChunk < nnz ? lo + kChunk : nnz;
            for (int64_t k = lo; k < hi; ++k) cnt[(int)(((int64_t)key[k] >> shift) & (kDigits - 1))]++;
            for (int d = 0; d < kDigits; ++d) counts[(int64_t)d * chunks + c] = cnt[d];
        });
        if (st != B200_OK) return st;
        const int64_t* cn = counts;
        st = scan::exclusive<int64_t>(
            ctx, ncount + 1, [=] __device__(int64_t i) -> int64_t { return i < ncount ? cn[i] : 0; }, offs, sums);
        if (st != B200_OK) return st;
        const int64_t* of = offs;
        st = launch_ew(ctx, chunks, 1, [=] __device__(int64_t c, int64_t) {
            int64_t pos[kDigits];
            for (int d = 0; d < kDigits; ++d) pos[d] = of[(int64_t)d * chunks + c];
            const int64_t lo = c * kChunk;
            const int64_t hi = lo + kChunk < nnz ? lo + kChunk : nnz;
            for (int64_t k = lo; k < hi; ++k) {
                const I kk = key[k];
                const int64_t dst = pos[(int)(((int64_t)kk >> shift) & (kDigits - 1))]++;
                key_o[dst] = kk;
                idx_o[dst] = idx[k];
            }
        });
        if (st != B200_OK) return st;
        I* t = key_a;
        key_a = key_b;
        key_b = t;
        t = idx_a;
        idx_a = idx_b;
        idx_b = t;
    }
    const I* key = key_a;
    const I* idx = idx_a;
    // t_row_ptrs[c] = first position whose column is >= c (empty columns repeat the value)
    st = launch_ew(ctx, nnz + 1, 1, [=] __device__(int64_t k, int64_t) {
        const int64_t prev = k == 0 ? -1 : (int64_t)key[k - 1];
        const int64_t cur = k == nnz ? num_cols : (int64_t)key[k];
        for (int64_t c = prev + 1; c <= cur; ++c) t_row_ptrs[c] = (I)k;
    });
    if (st != B200_OK) return st;
    return launch_ew(ctx, nnz, 1, [=] __device__(int64_t k, int64_t) {
        const I e = idx[k];
        // the row of entry e: last row pointer <= e
        int64_t lo = 0, hi = num_rows;  // invariant: row_ptrs[lo] <= e < row_ptrs[hi]
        while (hi - lo > 1) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (row_ptrs[mid] <= e)
                lo = mid;
            else
                hi = mid;
        }
        t_col_idxs[k] = (I)lo;
        t_values[k] = values[e];
    });
}

// out(j, i) = in(i, j) inside every block; element (row, col) of block k sits at
// group_offset * (k >> group_power) + block_offset * (k & mask) + row + col * stride
template <typename V, typename I>
b200_status jacobi_transpose(b200_ctx* ctx, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                             int64_t group_offset, int32_t group_power, const I* block_ptrs, const V* blocks,
                             V* out_blocks)
{
    B200_REQUIRE(ctx, "null argument");
    B200_REQUIRE(num_blocks >= 0, "negative size");
    B200_REQUIRE(max_block_size >= 1 && max_block_size <= 32, "max_block_size in [1, 32]");
    if (num_blocks == 0) return B200_OK;
    B200_REQUIRE(block_ptrs && blocks && out_blocks, "null argument");
    const int64_t stride = block_offset << group_power;
    const int64_t mask = ((int64_t)1 << group_power) - 1;
    const int64_t mbs = max_block_size;
    return launch_ew(ctx, num_blocks, mbs * mbs, [=] __device__(int64_t k, int64_t e) {
        const int64_t n = (int64_t)block_ptrs[k + 1] - (int64_t)block_ptrs[k];
        const int64_t i = e / mbs, j = e - i * mbs;
        if (i >= n || j >= n) return;
        const int64_t ofs = group_offset * (k >> group_power) + block_offset * (k & mask);
        out_blocks[ofs + i * stride + j] = blocks[ofs + i + j * stride];
    });
}

}  // namespace transpose
}  // namespace b200

extern "C" {

#define B200_DEF_BICG(V, VT)                                                                             \
    b200_status b200_bicg_initialize_##V(                                                                \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* b, int64_t bs, VT* r, int64_t rs, VT* z,    \
        int64_t zs, VT* p, int64_t ps, VT* q, int64_t qs, VT* prev_rho, VT* rho, VT* r2, int64_t r2s,    \
        VT* z2, int64_t z2s, VT* p2, int64_t p2s, VT* q2, int64_t q2s, uint8_t* stop)                    \
    {                                                                                                    \
        return b200::steps::bicg_initialize<VT>(ctx, rows, cols, b, bs, r, rs, z, zs, p, ps, q, qs,      \
                                                prev_rho, rho, r2, r2s, z2, z2s, p2, p2s, q2, q2s,       \
                                                stop);                                                   \
    }                                                                                                    \
    b200_status b200_bicg_step_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* p, int64_t ps,       \
                                     const VT* z, int64_t zs, VT* p2, int64_t p2s, const VT* z2,         \
                                     int64_t z2s, const VT* rho, const VT* prev_rho,                     \
                                     const uint8_t* stop)                                                \
    {                                                                                                    \
        return b200::steps::bicg_step_1<VT>(ctx, rows, cols, p, ps, z, zs, p2, p2s, z2, z2s, rho,        \
                                            prev_rho, stop);                                             \
    }                                                                                                    \
    b200_status b200_bicg_step_2_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x, int64_t xs,       \
                                     VT* r, int64_t rs, VT* r2, int64_t r2s, const VT* p, int64_t ps,    \
                                     const VT* q, int64_t qs, const VT* q2, int64_t q2s,                 \
                                     const VT* beta, const VT* rho, const uint8_t* stop)                 \
    {                                                                                                    \
        return b200::steps::bicg_step_2<VT>(ctx, rows, cols, x, xs, r, rs, r2, r2s, p, ps, q, qs, q2,    \
                                            q2s, beta, rho, stop);                                       \
    }
B200_DEF_BICG(f64, double)
B200_DEF_BICG(f32, float)

#define B200_DEF_TRANSPOSE(V, VT, I, IT)                                                                 \
    b200_status b200_csr_transpose_##V##_##I(b200_ctx* ctx, int64_t num_rows, int64_t num_cols,          \
                                             int64_t nnz, const IT* row_ptrs, const IT* col_idxs,        \
                                             const VT* values, IT* t_row_ptrs, IT* t_col_idxs,           \
                                             VT* t_values)                                               \
    {                                                                                                    \
        return b200::transpose::csr_transpose<VT, IT>(ctx, num_rows, num_cols, nnz, row_ptrs, col_idxs,  \
                                                      values, t_row_ptrs, t_col_idxs, t_values);         \
    }                                                                                                    \
    b200_status b200_jacobi_transpose_##V##_##I(b200_ctx* ctx, int64_t num_blocks,                       \
                                                int32_t max_block_size, int64_t block_offset,            \
                                                int64_t group_offset, int32_t group_power,               \
                                                const IT* block_ptrs, const VT* blocks,                  \
                                                VT* out_blocks)                                          \
    {                                                                                                    \
        return b200::transpose::jacobi_transpose<VT, IT>(ctx, num_blocks, max_block_size, block_offset,  \
                                                         group_offset, group_power, block_ptrs, blocks,  \
                                                         out_blocks);                                    \
    }
B200_DEF_TRANSPOSE(f64, double, i32, int32_t)
B200_DEF_TRANSPOSE(f64, double, i64, int64_t)
B200_DEF_TRANSPOSE(f32, float, i32, int32_t)
B200_DEF_TRANSPOSE(f32, float, i64, int64_t)

}  // extern "C"
