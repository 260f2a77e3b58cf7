// jacobi::find_blocks on the device: natural-block detection and
// supervariable agglomeration, bit-exact against reference/preconditioner/
// jacobi_kernels.cpp:36-107 (has_same_nonzero_pattern, find_natural_blocks,
// agglomerate_supervariables) + :112-123 (find_blocks).
//
// Both reference loops are sequential scans with a tiny state (the size of the block being
// grown, 1..max_block_size <= 32):
//   natural blocks:   item i = "row i has the pattern of row i-1";
//                     size < max && same ? size + 1 : (new block at row i, size = 1)
//   agglomeration:    item j = size s_j of natural block j;
//                     size + s_j <= max ? size + s_j : (new block at natural block j, s_j)
// i.e. finite-state machines.  They are evaluated in parallel the classic way: per chunk of
// items and per possible entry state (one thread each) the exit state; a serial composition
// over the (few thousand) chunks; then every chunk replays from its true entry state and
// marks the block starts.  The reference's own CUDA backend runs these two loops as
// <<<1, 1>>> kernels (common/cuda_hip/preconditioner/jacobi_kernels.cpp:246, :261).
#include "scan.cuh"

namespace b200 {
namespace jacobi {

constexpr int kChunk = 2048;

template <typename I>
__global__ void same_pattern_kernel(int64_t num_rows, const I* __restrict__ rp,
                                    const I* __restrict__ ci, uint8_t* __restrict__ same)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= num_rows) return;
    uint8_t s = 0;
    if (i > 0) {
        const int64_t p = rp[i - 1], c = rp[i], n = rp[i + 1];
        if (n - c == c - p) {
            s = 1;
            for (int64_t k = 0; k < n - c; ++k)
                if (ci[c + k] != ci[p + k]) {
                    s = 0;
                    break;
                }
        }
    }
    same[i] = s;
}

// MODE 0: items are `same` flags (item 0 is the first row: always a start, state 1)
// MODE 1: items are natural-block sizes (item 0 always a start, state s_0)
template <int MODE, typename I>
__device__ __forceinline__ int step(int state, int max_bs, const uint8_t* same, const I* nat_ptrs,
                                    int64_t i, bool* start)
{
    if (MODE == 0) {
        if (state < max_bs && same[i]) {
            *start = false;
            return state + 1;
        }
        *start = true;
        return 1;
    } else {
        const int s = (int)(nat_ptrs[i + 1] - nat_ptrs[i]);
        if (state + s <= max_bs) {
            *start = false;
            return state + s;
        }
        *start = true;
        return s;
    }
}

// exit state of every chunk for every entry state 1..32 (thread = entry state - 1)
template <int MODE, typename I>
__global__ void __launch_bounds__(32)
    chunk_map_kernel(int64_t num_items, int max_bs, const uint8_t* __restrict__ same,
                     const I* __restrict__ nat_ptrs, uint8_t* __restrict__ exit_state)
{
    const int64_t c = blockIdx.x;
    const int64_t lo = c * kChunk > 1 ? c * kChunk : 1;  // item 0 is handled by the composition
    int64_t hi = (c + 1) * (int64_t)kChunk;
    if (hi > num_items) hi = num_items;
    int state = threadIdx.x + 1;
    bool st;
    if (state <= max_bs)
        for (int64_t i = lo; i < hi; ++i) state = step<MODE, I>(state, max_bs, same, nat_ptrs, i, &st);
    exit_state[c * 32 + threadIdx.x] = (uint8_t)state;
}

template <int MODE, typename I>
__global__ void compose_kernel(int64_t num_chunks, int max_bs, const I* __restrict__ nat_ptrs,
                               const uint8_t* __restrict__ exit_state,
                               uint8_t* __restrict__ entry_state)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int state = MODE == 0 ? 1 : (int)(nat_ptrs[1] - nat_ptrs[0]);  // after item 0
    for (int64_t c = 0; c < num_chunks; ++c) {
        entry_state[c] = (uint8_t)state;
        state = exit_state[c * 32 + state - 1];
    }
}

template <int MODE, typename I>
__global__ void emit_kernel(int64_t num_items, int64_t num_chunks, int max_bs,
                            const uint8_t* __restrict__ same, const I* __restrict__ nat_ptrs,
                            const uint8_t* __restrict__ entry_state, uint8_t* __restrict__ start)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= num_chunks) return;
    int64_t lo = c * (int64_t)kChunk;
    int64_t hi = lo + kChunk;
    if (hi > num_items) hi = num_items;
    int state = entry_state[c];
    if (lo == 0) {
        start[0] = 1;
        lo = 1;
    }
    for (int64_t i = lo; i < hi; ++i) {
        bool st;
        state = step<MODE, I>(state, max_bs, same, nat_ptrs, i, &st);
        start[i] = st ? 1 : 0;
    }
}

// ptrs[rank of i among the starts] = MODE 0 ? i : nat_ptrs[i]
template <int MODE, typename I>
__global__ void scatter_starts_kernel(int64_t num_items, const uint8_t* __restrict__ start,
                                      const int64_t* __restrict__ rank,
                                      const I* __restrict__ nat_ptrs, I* __restrict__ ptrs)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= num_items || !start[i]) return;
    ptrs[rank[i]] = MODE == 0 ? (I)i : nat_ptrs[i];
}

// one FSM pass: marks the starts among `num_items` items, compacts them into `out_ptrs`
// (count entries, caller appends the end); returns the count on the host
template <int MODE, typename I>
b200_status fsm_pass(b200_ctx* ctx, int64_t num_items, int max_bs, const uint8_t* same,
                     const I* nat_ptrs, uint8_t* start, uint8_t* exit_state, uint8_t* entry_state,
                     int64_t* rank, int64_t* tile_sums, I* out_ptrs, int64_t* count_host)
{
    const int64_t chunks = ceildiv(num_items, (int64_t)kChunk);
    chunk_map_kernel<MODE, I><<<(unsigned)chunks, 32, 0, ctx->stream>>>(num_items, max_bs, same, nat_ptrs,
                                                                       exit_state);
    B200_LAUNCH_CHECK(ctx);
    compose_kernel<MODE, I><<<1, 32, 0, ctx->stream>>>(chunks, max_bs, nat_ptrs, exit_state, entry_state);
    B200_LAUNCH_CHECK(ctx);
    emit_kernel<MODE, I><<<(unsigned)ceildiv(chunks, (int64_t)128), 128, 0, ctx->stream>>>(
        num_items, chunks, max_bs, same, nat_ptrs, entry_state, start);
    B200_LAUNCH_CHECK(ctx);
    const uint8_t* cs = start;
    b200_status st = scan::exclusive<int64_t>(
        ctx, num_items + 1,
        [=] __device__(int64_t i) -> int64_t { return i < num_items ? (int64_t)cs[i] : 0; }, rank,
        tile_sums);
    if (st != B200_OK) return st;
    scatter_starts_kernel<MODE, I><<<(unsigned)ceildiv(num_items, (int64_t)256), 256, 0, ctx->stream>>>(
        num_items, start, rank, nat_ptrs, out_ptrs);
    B200_LAUNCH_CHECK(ctx);
    B200_CUDA_CHECK(cudaMemcpyAsync(count_host, rank + num_items, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

template <typename I>
__global__ void set_last_kernel(I* ptrs, int64_t pos, I value)
{
    ptrs[pos] = value;
}

template <typename I>
b200_status find_blocks(b200_ctx* ctx, int64_t num_rows, const I* row_ptrs, const I* col_idxs,
                        int32_t max_block_size, I* block_ptrs, int64_t* num_blocks_host)
{
    B200_REQUIRE(ctx && block_ptrs && num_blocks_host, "null argument");
    B200_REQUIRE(max_block_size >= 1 && max_block_size <= 32, "max_block_size in [1, 32]");
    B200_REQUIRE(num_rows >= 0, "negative size");
    *num_blocks_host = 0;
    if (num_rows == 0) {
        set_last_kernel<I><<<1, 1, 0, ctx->stream>>>(block_ptrs, 0, I(0));
        B200_LAUNCH_CHECK(ctx);
        return B200_OK;
    }
    B200_REQUIRE(row_ptrs, "null pointer");
    const int64_t n = num_rows;
    const int64_t chunks = ceildiv(n, (int64_t)kChunk);
    // one scratch block: same[n] | start[n] | exit[chunks*32] | entry[chunks] | rank[n+1] |
    //                    tile sums | natural ptrs[n+1]
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t o_same = 0, o_start = o_same + al(n), o_exit = o_start + al(n),
                 o_entry = o_exit + al(chunks * 32), o_rank = o_entry + al(chunks),
                 o_sums = o_rank + al(sizeof(int64_t) * (n + 1)),
                 o_nat = o_sums + al(sizeof(int64_t) * scan::num_tiles(n + 1)),
                 total = o_nat + al(sizeof(I) * (n + 1));
    char* base = (char*)ctx->scratch(total);
    if (!base) return B200_ERR_ALLOC;
    uint8_t* same = (uint8_t*)(base + o_same);
    uint8_t* start = (uint8_t*)(base + o_start);
    uint8_t* exit_state = (uint8_t*)(base + o_exit);
    uint8_t* entry_state = (uint8_t*)(base + o_entry);
    int64_t* rank = (int64_t*)(base + o_rank);
    int64_t* sums = (int64_t*)(base + o_sums);
    I* nat = (I*)(base + o_nat);
    same_pattern_kernel<I><<<(unsigned)ceildiv(n, (int64_t)256), 256, 0, ctx->stream>>>(n, row_ptrs, col_idxs,
                                                                                      same);
    B200_LAUNCH_CHECK(ctx);
    int64_t num_nat = 0;
    b200_status st = fsm_pass<0, I>(ctx, n, max_block_size, same, (const I*)nullptr, start, exit_state,
                                    entry_state, rank, sums, nat, &num_nat);
    if (st != B200_OK) return st;
    set_last_kernel<I><<<1, 1, 0, ctx->stream>>>(nat, num_nat, (I)n);
    B200_LAUNCH_CHECK(ctx);
    int64_t num_blocks = 0;
    st = fsm_pass<1, I>(ctx, num_nat, max_block_size, (const uint8_t*)nullptr, nat, start, exit_state,
                        entry_state, rank, sums, block_ptrs, &num_blocks);
    if (st != B200_OK) return st;
    set_last_kernel<I><<<1, 1, 0, ctx->stream>>>(block_ptrs, num_blocks, (I)n);
    B200_LAUNCH_CHECK(ctx);
    *num_blocks_host = num_blocks;
    return B200_OK;
}

}  // namespace jacobi
}  // namespace b200

extern "C" {

b200_status b200_jacobi_find_blocks_i32(b200_ctx* ctx, int64_t num_rows, const int32_t* row_ptrs,
                                        const int32_t* col_idxs, int32_t max_block_size,
                                        int32_t* block_pointers, int64_t* num_blocks_host)
{
    return b200::jacobi::find_blocks<int32_t>(ctx, num_rows, row_ptrs, col_idxs, max_block_size,
                                              block_pointers, num_blocks_host);
}
b200_status b200_jacobi_find_blocks_i64(b200_ctx* ctx, int64_t num_rows, const int64_t* row_ptrs,
                                        const int64_t* col_idxs, int32_t max_block_size,
                                        int64_t* block_pointers, int64_t* num_blocks_host)
{
    return b200::jacobi::find_blocks<int64_t>(ctx, num_rows, row_ptrs, col_idxs, max_block_size,
                                              block_pointers, num_blocks_host);
}

}  // extern "C"
