// Block-Jacobi set-up on the device: extract every diagonal block of
// the CSR matrix, invert it with the reference's pivoted Gauss-Jordan and store it
// transposed and column-permuted in block_interleaved_storage_scheme -- bit-identical to
// reference/preconditioner/jacobi_kernels.cpp:113-147 (extract_block), :150-205
// (choose_pivot, swap_rows, apply_gauss_jordan_transform), :262-278 (invert_block),
// :243-258 (permute_and_transpose_block) and the full-precision branch of generate
// (:320-410).  Adaptive precision / conditioning are out of scope (block_precisions == NULL).
//
// One warp per block, the block lives in shared memory ([32][33], padded).  The elimination
// keeps the reference's operation order element by element:
//   pivot search (first maximum of |column k| below the diagonal), row swap, d = b[k][k],
//   column k /= -d, b[k][k] = 0, b[i][j] += b[i][k] * b[k][j] for ALL i, j (mul and add
//   rounded separately: the file is compiled with -fmad=false), row k /= d, b[k][k] = 1 / d.
// Every entry's arithmetic is the same sequence of IEEE operations as in the sequential
// code (for finite data the i == k / j == k terms of the update add an exact zero), so the
// inverse agrees bit for bit.  A zero pivot stops the block where the reference stops.
#include "common.cuh"

namespace b200 {
namespace jacobi {

constexpr int kGenWarps = 4;

template <typename V>
__device__ __forceinline__ V vabs(V x)
{
    return x < V(0) ? -x : x;
}

template <typename V, typename I>
__global__ void __launch_bounds__(kGenWarps * 32)
    generate_kernel(int64_t num_blocks, const I* __restrict__ rp, const I* __restrict__ ci,
                    const V* __restrict__ va, const I* __restrict__ block_ptrs,
                    int64_t block_offset, int64_t group_offset, int32_t group_power,
                    V* __restrict__ blocks)
{
    __shared__ V sm[kGenWarps][32][33];
    __shared__ int sperm[kGenWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x * kGenWarps + warp;
    if (blk >= num_blocks) return;
    V(*b)[33] = sm[warp];
    int* perm = sperm[warp];
    const int64_t start = block_ptrs[blk];
    const int bs = (int)((int64_t)block_ptrs[blk + 1] - start);
    // ---- extract_block
    for (int i = 0; i < bs; ++i)
        if (lane < bs) b[i][lane] = V(0);
    perm[lane] = lane;
    __syncwarp();
    for (int row = 0; row < bs; ++row) {
        const int64_t s = rp[start + row], e = rp[start + row + 1];
        for (int64_t p = s + lane; p < e; p += 32) {
            const int64_t col = (int64_t)ci[p] - start;
            if (col >= 0 && col < bs) b[row][col] = va[p];
        }
    }
    __syncwarp();
    // ---- invert_block
    bool ok = true;
    for (int k = 0; k < bs && ok; ++k) {
        // choose_pivot: first row i >= k with the largest |b[i][k]| (strict <)
        V best = (lane >= k && lane < bs) ? vabs(b[lane][k]) : V(-1);
        int arg = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const V ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
        }
        const int cp = arg;
        // swap_rows(k, cp) and the permutation
        if (cp != k) {
            if (lane < bs) {
                const V t = b[k][lane];
                b[k][lane] = b[cp][lane];
                b[cp][lane] = t;
            }
            if (lane == 0) {
                const int t = perm[k];
                perm[k] = perm[cp];
                perm[cp] = t;
            }
        }
        __syncwarp();
        // apply_gauss_jordan_transform(k, k)
        const V d = b[k][k];
        if (d == V(0)) {
            ok = false;
            break;
        }
        __syncwarp();
        if (lane < bs) b[lane][k] = b[lane][k] / (-d);
        __syncwarp();
        if (lane == 0) b[k][k] = V(0);
        __syncwarp();
        if (lane < bs) {
            const V bkj = b[k][lane];  // row k is not changed by the update (adds 0 * x)
            for (int i = 0; i < bs; ++i) {
                const V prod = b[i][k] * bkj;
                b[i][lane] = b[i][lane] + prod;
            }
        }
        __syncwarp();
        if (lane < bs) b[k][lane] = b[k][lane] / d;
        __syncwarp();
        if (lane == 0) b[k][k] = V(1) / d;
        __syncwarp();
    }
    // ---- permute_and_transpose_block into the interleaved storage
    const int64_t stride = block_offset << group_power;
    V* dst = blocks + group_offset * (blk >> group_power) +
             block_offset * (blk & ((int64_t(1) << group_power) - 1));
    if (lane < bs)
        for (int j = 0; j < bs; ++j) dst[lane + (int64_t)perm[j] * stride] = b[lane][j];
}

}  // namespace jacobi
}  // namespace b200

extern "C" {

#define B200_DEF_JACOBI_GENERATE(V, VT, I, IT)                                                 \
    b200_status b200_jacobi_generate_##V##_##I(                                                \
        b200_ctx* ctx, int64_t num_rows, const IT* row_ptrs, const IT* col_idxs,               \
        const VT* values, int64_t num_blocks, int32_t max_block_size, int64_t block_offset,    \
        int64_t group_offset, int32_t group_power, const IT* block_pointers, VT* blocks)       \
    {                                                                                          \
        B200_REQUIRE(ctx != nullptr, "ctx is null");                                           \
        B200_REQUIRE(max_block_size >= 1 && max_block_size <= 32, "max_block_size in [1, 32]"); \
        B200_REQUIRE(num_blocks >= 0 && num_rows >= 0, "negative size");                       \
        if (num_blocks == 0) return B200_OK;                                                   \
        B200_REQUIRE(row_ptrs && block_pointers && blocks, "null pointer");                    \
        b200::jacobi::generate_kernel<VT, IT>                                                  \
            <<<(unsigned)b200::ceildiv(num_blocks, b200::jacobi::kGenWarps),                   \
               b200::jacobi::kGenWarps * 32, 0, ctx->stream>>>(                                \
                num_blocks, row_ptrs, col_idxs, values, block_pointers, block_offset,          \
                group_offset, group_power, blocks);                                            \
        B200_LAUNCH_CHECK(ctx);                                                                \
        return B200_OK;                                                                        \
    }
B200_DEF_JACOBI_GENERATE(f64, double, i32, int32_t)
B200_DEF_JACOBI_GENERATE(f64, double, i64, int64_t)
B200_DEF_JACOBI_GENERATE(f32, float, i32, int32_t)
B200_DEF_JACOBI_GENERATE(f32, float, i64, int64_t)

}  // extern "C"
