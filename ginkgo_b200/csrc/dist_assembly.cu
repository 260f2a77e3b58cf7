// Distributed set-up on the device: the row/column partition, the split
// of a global matrix into the entries a rank owns, and the index map of the remote columns --
// what experimental::distributed::Matrix::read_distributed runs before the first apply
// (core/distributed/matrix.cpp:300-380).  Bit-exact against
//   partition::*                               reference/distributed/partition_kernels.cpp:18-160
//   find_range / map_to_local                  reference/distributed/partition_helpers.hpp:24-56
//   distributed_matrix::separate_local_nonlocal   reference/distributed/matrix_kernels.cpp:18-90
//   index_map::build_mapping / map_to_local    reference/distributed/index_map_kernels.cpp:20-212
//
// The reference walks its inputs sequentially (or, in its CUDA backend, sorts them with
// thrust).  Here everything is an element-wise pass plus exclusive scans (scan.cuh):
//   * separate_local_nonlocal = classify every entry (binary search of the row and column in the
//     range bounds), two scans for the stable positions in the local / non-local lists, one
//     scatter;
//   * the index map's set of remote columns is a BITMAP over the global column space with a
//     per-word running popcount: "sort + unique by (part id, global index)" becomes
//     rank(g) = offset of g's range in (part, range) order + set bits of that range below g.
//     1/8 byte + 1/4 byte per global column instead of a sort of the non-local entries.
// Element-wise lambdas only, so a copy of this file also compiles for the host and is checked
// there without a GPU (tests/test_dist_assembly_cpu.py).
#include "elementwise.cuh"
#include "scan.cuh"

namespace b200 {
namespace dist_assembly {

constexpr int32_t kLocal = 0, kNonLocal = 1, kCombined = 2;

// std::upper_bound over bounds[1 .. num_ranges]: the range that holds idx
template <typename G>
__device__ __forceinline__ int64_t find_range(const G* __restrict__ bounds, int64_t num_ranges, G idx)
{
    int64_t lo = 0, hi = num_ranges;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (bounds[1 + mid] <= idx)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool bit_set(const uint32_t* __restrict__ bitmap, int64_t g)
{
    return (bitmap[g >> 5] >> (g & 31)) & 1u;
}

// number of set bits below global index g
__device__ __forceinline__ int64_t rank_at(const uint32_t* __restrict__ bitmap,
                                           const int64_t* __restrict__ word_rank, int64_t g)
{
    const int b = (int)(g & 31);
    return word_rank[g >> 5] + (b ? __popc(bitmap[g >> 5] & ((1u << b) - 1u)) : 0);
}

inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

template <typename T>
b200_status read_back(b200_ctx* ctx, T* host, const T* dev, int64_t count)
{
    B200_CUDA_CHECK(cudaMemcpyAsync(host, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------- partition
template <typename G>
b200_status build_ranges_from_global_size(b200_ctx* ctx, int32_t num_parts, int64_t global_size, G* ranges)
{
    B200_REQUIRE(ctx && ranges, "null argument");
    B200_REQUIRE(num_parts >= 0 && global_size >= 0, "negative size");
    const int64_t per = num_parts ? global_size / num_parts : 0;
    const int64_t rest = global_size - (int64_t)num_parts * per;
    // ranges[i] = ranges[i-1] + per + (i-1 < rest) in closed form
    return launch_ew(ctx, (int64_t)num_parts + 1, 1, [=] __device__(int64_t i, int64_t) {
        ranges[i] = (G)(i * per + (i < rest ? i : rest));
    });
}

template <typename G>
b200_status build_from_contiguous(b200_ctx* ctx, int64_t num_ranges, const G* ranges,
                                  const int32_t* part_id_mapping, G* range_bounds, int32_t* part_ids)
{
    B200_REQUIRE(ctx && ranges && range_bounds && (part_ids || num_ranges == 0), "null argument");
    B200_REQUIRE(num_ranges >= 0, "negative size");
    return launch_ew(ctx, num_ranges + 1, 1, [=] __device__(int64_t i, int64_t) {
        range_bounds[i] = i == 0 ? G(0) : ranges[i];
        if (i < num_ranges) part_ids[i] = part_id_mapping ? part_id_mapping[i] : (int32_t)i;
    });
}

// a range starts wherever the owner changes (the reference starts from part -1)
__device__ __forceinline__ int64_t range_start_flag(const int32_t* __restrict__ mapping, int64_t i)
{
    return mapping[i] != (i == 0 ? -1 : mapping[i - 1]) ? 1 : 0;
}

inline b200_status count_ranges(b200_ctx* ctx, int64_t n, const int32_t* mapping, int64_t* num_ranges_host)
{
    B200_REQUIRE(ctx && num_ranges_host && (mapping || n == 0), "null argument");
    B200_REQUIRE(n >= 0, "negative size");
    *num_ranges_host = 0;
    if (n == 0) return B200_OK;
    const size_t o_sums = al256(sizeof(int64_t) * (n + 1));
    char* base = (char*)ctx->scratch(o_sums + al256(sizeof(int64_t) * scan::num_tiles(n + 1)));
    if (!base) return B200_ERR_ALLOC;
    int64_t* rank = (int64_t*)base;
    b200_status st = scan::exclusive<int64_t>(
        ctx, n + 1, [=] __device__(int64_t i) -> int64_t { return i < n ? range_start_flag(mapping, i) : 0; },
        rank, (int64_t*)(base + o_sums));
    if (st != B200_OK) return st;
    return read_back(ctx, num_ranges_host, rank + n, 1);
}

template <typename G>
b200_status build_from_mapping(b200_ctx* ctx, int64_t n, const int32_t* mapping, G* range_bounds,
                               int32_t* part_ids)
{
    B200_REQUIRE(ctx && range_bounds && (mapping || n == 0), "null argument");
    B200_REQUIRE(n >= 0, "negative size");
    if (n == 0)
        return launch_ew(ctx, 1, 1, [=] __device__(int64_t, int64_t) { range_bounds[0] = G(0); });
    const size_t o_sums = al256(sizeof(int64_t) * (n + 1));
    char* base = (char*)ctx->scratch(o_sums + al256(sizeof(int64_t) * scan::num_tiles(n + 1)));
    if (!base) return B200_ERR_ALLOC;
    int64_t* rank = (int64_t*)base;
    b200_status st = scan::exclusive<int64_t>(
        ctx, n + 1, [=] __device__(int64_t i) -> int64_t { return i < n ? range_start_flag(mapping, i) : 0; },
        rank, (int64_t*)(base + o_sums));
    if (st != B200_OK) return st;
    return launch_ew(ctx, n + 1, 1, [=] __device__(int64_t i, int64_t) {
        if (i == n) {
            range_bounds[rank[n]] = (G)n;
        } else if (range_start_flag(mapping, i)) {
            range_bounds[rank[i]] = (G)i;
            part_ids[rank[i]] = mapping[i];
        }
    });
}

// starting_indices[r] = rows of r's part in earlier ranges; one thread per part walks the ranges
// in order (the reference's loop, partitioned by owner)
template <typename L, typename G>
b200_status build_starting_indices(b200_ctx* ctx, int64_t num_ranges, int32_t num_parts,
                                   const G* range_bounds, const int32_t* part_ids, L* starting_indices,
                                   L* part_sizes, int32_t* num_empty_parts_host)
{
    B200_REQUIRE(ctx && num_empty_parts_host, "null argument");
    B200_REQUIRE(num_ranges >= 0 && num_parts >= 0, "negative size");
    *num_empty_parts_host = 0;
    if (num_parts == 0) return B200_OK;
    B200_REQUIRE(part_sizes && (num_ranges == 0 || (range_bounds && part_ids && starting_indices)),
                 "null argument");
    int32_t* empty = (int32_t*)ctx->scratch(256);
    if (!empty) return B200_ERR_ALLOC;
    b200_status st = launch_ew(ctx, 1, 1, [=] __device__(int64_t, int64_t) { *empty = 0; });
    if (st != B200_OK) return st;
    st = launch_ew(ctx, (int64_t)num_parts, 1, [=] __device__(int64_t p, int64_t) {
        L run = 0;
        for (int64_t r = 0; r < num_ranges; ++r)
            if (part_ids[r] == (int32_t)p) {
                starting_indices[r] = run;
                run += (L)(range_bounds[r + 1] - range_bounds[r]);
            }
        part_sizes[p] = run;
        if (run == 0) atomicAdd(empty, 1);
    });
    if (st != B200_OK) return st;
    return read_back(ctx, num_empty_parts_host, empty, 1);
}

inline b200_status has_ordered_parts(b200_ctx* ctx, int64_t num_ranges, const int32_t* part_ids,
                                     int32_t* result_host)
{
    B200_REQUIRE(ctx && result_host && (part_ids || num_ranges == 0), "null argument");
    *result_host = 1;
    if (num_ranges < 2) return B200_OK;
    int32_t* unordered = (int32_t*)ctx->scratch(256);
    if (!unordered) return B200_ERR_ALLOC;
    b200_status st = launch_ew(ctx, 1, 1, [=] __device__(int64_t, int64_t) { *unordered = 0; });
    if (st != B200_OK) return st;
    st = launch_ew(ctx, num_ranges - 1, 1, [=] __device__(int64_t i, int64_t) {
        if (part_ids[i + 1] < part_ids[i]) atomicAdd(unordered, 1);
    });
    if (st != B200_OK) return st;
    int32_t u = 0;
    st = read_back(ctx, &u, unordered, 1);
    *result_host = u ? 0 : 1;
    return st;
}

// ----------------------------------------------------------------- separate_local_nonlocal
template <typename G>
b200_status classify_entries(b200_ctx* ctx, int64_t nnz, const G* row_idxs, const G* col_idxs,
                             int64_t row_num_ranges, const G* row_bounds, const int32_t* row_part_ids,
                             int64_t col_num_ranges, const G* col_bounds, const int32_t* col_part_ids,
                             int32_t local_part, uint8_t* cls, int64_t* local_rank,
                             int64_t* non_local_rank, int64_t* num_local_host, int64_t* num_non_local_host)
{
    B200_REQUIRE(ctx && num_local_host && num_non_local_host && local_rank && non_local_rank, "null argument");
    B200_REQUIRE(nnz >= 0 && row_num_ranges >= 0 && col_num_ranges >= 0, "negative size");
    B200_REQUIRE(nnz == 0 || (row_idxs && col_idxs && cls && row_bounds && row_part_ids && col_bounds &&
                              col_part_ids),
                 "null argument");
    B200_REQUIRE(nnz == 0 || (row_num_ranges > 0 && col_num_ranges > 0), "empty partition");
    b200_status st = launch_ew(ctx, nnz, 1, [=] __device__(int64_t i, int64_t) {
        uint8_t c = 0;
        if (row_part_ids[find_range(row_bounds, row_num_ranges, row_idxs[i])] == local_part)
            c = col_part_ids[find_range(col_bounds, col_num_ranges, col_idxs[i])] == local_part ? 1 : 2;
        cls[i] = c;
    });
    if (st != B200_OK) return st;
    int64_t* sums = (int64_t*)ctx->scratch(sizeof(int64_t) * scan::num_tiles(nnz + 1));
    if (!sums) return B200_ERR_ALLOC;
    const uint8_t* c = cls;
    st = scan::exclusive<int64_t>(
        ctx, nnz + 1, [=] __device__(int64_t i) -> int64_t { return i < nnz && c[i] == 1 ? 1 : 0; },
        local_rank, sums);
    if (st != B200_OK) return st;
    st = scan::exclusive<int64_t>(
        ctx, nnz + 1, [=] __device__(int64_t i) -> int64_t { return i < nnz && c[i] == 2 ? 1 : 0; },
        non_local_rank, sums);
    if (st != B200_OK) return st;
    st = read_back(ctx, num_local_host, local_rank + nnz, 1);
    if (st != B200_OK) return st;
    return read_back(ctx, num_non_local_host, non_local_rank + nnz, 1);
}

template <typename V, typename L, typename G>
b200_status separate_fill(b200_ctx* ctx, int64_t nnz, const G* row_idxs, const G* col_idxs, const V* values,
                          int64_t row_num_ranges, const G* row_bounds, const L* row_starting,
                          int64_t col_num_ranges, const G* col_bounds, const L* col_starting,
                          const uint8_t* cls, const int64_t* local_rank, const int64_t* non_local_rank,
                          L* local_rows, L* local_cols, V* local_vals, L* non_local_rows, G* non_local_cols,
                          V* non_local_vals)
{
    B200_REQUIRE(ctx, "null argument");
    B200_REQUIRE(nnz >= 0, "negative size");
    return launch_ew(ctx, nnz, 1, [=] __device__(int64_t i, int64_t) {
        const uint8_t c = cls[i];
        if (!c) return;
        const G row = row_idxs[i], col = col_idxs[i];
        const int64_t rr = find_range(row_bounds, row_num_ranges, row);
        const L lrow = (L)(row - row_bounds[rr]) + row_starting[rr];
        if (c == 1) {
            const int64_t cr = find_range(col_bounds, col_num_ranges, col);
            const int64_t k = local_rank[i];
            local_rows[k] = lrow;
            local_cols[k] = (L)(col - col_bounds[cr]) + col_starting[cr];
            local_vals[k] = values[i];
        } else {
            const int64_t k = non_local_rank[i];
            non_local_rows[k] = lrow;
            non_local_cols[k] = col;
            non_local_vals[k] = values[i];
        }
    });
}

template <typename V, typename L, typename G>
b200_status kept_fill(b200_ctx* ctx, int64_t nnz, const G* row_idxs, const G* col_idxs, const V* values,
                      int64_t row_num_ranges, const G* row_bounds, const L* row_starting, const uint8_t* cls,
                      const int64_t* local_rank, const int64_t* non_local_rank, L* rows, G* cols, V* vals)
{
    B200_REQUIRE(ctx, "null argument");
    B200_REQUIRE(nnz >= 0, "negative size");
    return launch_ew(ctx, nnz, 1, [=] __device__(int64_t i, int64_t) {
        if (!cls[i]) return;
        const G row = row_idxs[i];
        const int64_t rr = find_range(row_bounds, row_num_ranges, row);
        const int64_t k = local_rank[i] + non_local_rank[i];
        rows[k] = (L)(row - row_bounds[rr]) + row_starting[rr];
        cols[k] = col_idxs[i];
        vals[k] = values[i];
    });
}

// distributed_vector::build_local (reference/distributed/vector_kernels.cpp:15-40): scatter the
// entries of the owned rows into the pre-zeroed row-major local block.  Like the reference's device
// backends this is a plain scatter: (row, column) pairs must be unique (sum_duplicates first).
template <typename V, typename L, typename G>
b200_status vector_build_local(b200_ctx* ctx, int64_t nnz, const G* row_idxs, const G* col_idxs, const V* values,
                               int64_t num_ranges, const G* bounds, const int32_t* part_ids, const L* starting,
                               int32_t local_part, V* local_values, int64_t local_stride)
{
    B200_REQUIRE(ctx, "null argument");
    B200_REQUIRE(nnz >= 0 && num_ranges >= 0 && local_stride >= 0, "negative size");
    B200_REQUIRE(nnz == 0 || (row_idxs && col_idxs && values && bounds && part_ids && starting && local_values &&
                              num_ranges > 0),
                 "null argument");
    return launch_ew(ctx, nnz, 1, [=] __device__(int64_t i, int64_t) {
        const G row = row_idxs[i];
        const int64_t r = find_range(bounds, num_ranges, row);
        if (part_ids[r] != local_part) return;
        const int64_t lrow = (int64_t)((L)(row - bounds[r]) + starting[r]);
        local_values[lrow * local_stride + (int64_t)col_idxs[i]] = values[i];
    });
}

// --------------------------------------------------------------------------------- index map
inline int64_t num_words(int64_t global_size) { return (global_size + 31) / 32; }

template <typename G>
b200_status index_map_mark(b200_ctx* ctx, int64_t global_size, int64_t num_ranges, const G* bounds,
                           const int32_t* part_ids, int32_t skip_part, int64_t m, const G* global_idxs,
                           uint32_t* bitmap)
{
    B200_REQUIRE(ctx && bitmap, "null argument");
    B200_REQUIRE(global_size >= 0 && m >= 0 && num_ranges >= 0, "negative size");
    B200_REQUIRE(m == 0 || global_idxs, "null argument");
    B200_REQUIRE(skip_part < 0 || num_ranges == 0 || (bounds && part_ids), "null argument");
    b200_status st = launch_ew(ctx, num_words(global_size) + 1, 1,
                               [=] __device__(int64_t w, int64_t) { bitmap[w] = 0u; });
    if (st != B200_OK) return st;
    return launch_ew(ctx, m, 1, [=] __device__(int64_t i, int64_t) {
        const G g = global_idxs[i];
        if (g < 0 || (int64_t)g >= global_size) return;
        if (skip_part >= 0 && part_ids[find_range(bounds, num_ranges, g)] == skip_part) return;
        atomicOr(bitmap + (g >> 5), 1u << (g & 31));
    });
}

template <typename G>
b200_status index_map_rank(b200_ctx* ctx, int64_t global_size, int64_t num_ranges, int32_t num_parts,
                           const G* bounds, const int32_t* part_ids, const uint32_t* bitmap,
                           int64_t* word_rank, int64_t* range_offsets, int64_t* remote_sizes,
                           int64_t* num_remote_host)
{
    B200_REQUIRE(ctx && bitmap && word_rank && num_remote_host, "null argument");
    B200_REQUIRE(global_size >= 0 && num_ranges >= 0 && num_parts >= 0, "negative size");
    B200_REQUIRE(num_ranges == 0 || (bounds && part_ids && range_offsets), "null argument");
    B200_REQUIRE(num_parts == 0 || remote_sizes, "null argument");
    const int64_t words = num_words(global_size);
    const size_t o_off = al256(sizeof(int64_t) * scan::num_tiles(words + 1));
    char* base = (char*)ctx->scratch(o_off + al256(sizeof(int64_t) * (size_t)(num_parts + 1)));
    if (!base) return B200_ERR_ALLOC;
    int64_t* part_off = (int64_t*)(base + o_off);
    b200_status st = scan::exclusive<int64_t>(
        ctx, words + 1,
        [=] __device__(int64_t w) -> int64_t { return w < words ? (int64_t)__popc(bitmap[w]) : 0; }, word_rank,
        (int64_t*)base);
    if (st != B200_OK) return st;
    // inside a part the ranges keep their order: one thread per part accumulates the counts
    st = launch_ew(ctx, (int64_t)num_parts, 1, [=] __device__(int64_t p, int64_t) {
        int64_t run = 0;
        for (int64_t r = 0; r < num_ranges; ++r)
            if (part_ids[r] == (int32_t)p) {
                range_offsets[r] = run;
                run += rank_at(bitmap, word_rank, (int64_t)bounds[r + 1]) -
                       rank_at(bitmap, word_rank, (int64_t)bounds[r]);
            }
        remote_sizes[p] = run;
    });
    if (st != B200_OK) return st;
    // parts in ascending order
    st = launch_ew(ctx, 1, 1, [=] __device__(int64_t, int64_t) {
        int64_t run = 0;
        for (int32_t p = 0; p < num_parts; ++p) {
            part_off[p] = run;
            run += remote_sizes[p];
        }
    });
    if (st != B200_OK) return st;
    st = launch_ew(ctx, num_ranges, 1,
                   [=] __device__(int64_t r, int64_t) { range_offsets[r] += part_off[part_ids[r]]; });
    if (st != B200_OK) return st;
    return read_back(ctx, num_remote_host, word_rank + words, 1);
}

template <typename L, typename G>
b200_status index_map_fill(b200_ctx* ctx, int64_t global_size, int64_t num_ranges, const G* bounds,
                           const int32_t* part_ids, const L* starting, const uint32_t* bitmap,
                           const int64_t* word_rank, const int64_t* range_offsets, G* remote_global_idxs,
                           L* remote_local_idxs, int32_t* remote_part_ids)
{
    B200_REQUIRE(ctx && bitmap && word_rank, "null argument");
    B200_REQUIRE(global_size >= 0 && num_ranges >= 0, "negative size");
    return launch_ew(ctx, num_words(global_size), 1, [=] __device__(int64_t w, int64_t) {
        uint32_t bits = bitmap[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int64_t g = w * 32 + b;
            const int64_t r = find_range(bounds, num_ranges, (G)g);
            const int64_t k = range_offsets[r] + rank_at(bitmap, word_rank, g) -
                              rank_at(bitmap, word_rank, (int64_t)bounds[r]);
            remote_global_idxs[k] = (G)g;
            remote_local_idxs[k] = (L)(g - (int64_t)bounds[r]) + starting[r];
            if (remote_part_ids) remote_part_ids[k] = part_ids[r];
        }
    });
}

template <typename L, typename G>
b200_status index_map_map_to_local(b200_ctx* ctx, int64_t global_size, int64_t num_ranges, const G* bounds,
                                   const int32_t* part_ids, const L* starting, const uint32_t* bitmap,
                                   const int64_t* word_rank, const int64_t* range_offsets, int32_t rank,
                                   L local_size, int32_t index_space, int64_t m, const G* global_ids,
                                   L* local_ids)
{
    B200_REQUIRE(ctx, "null argument");
    B200_REQUIRE(global_size >= 0 && num_ranges >= 0 && m >= 0, "negative size");
    B200_REQUIRE(index_space >= kLocal && index_space <= kCombined, "index_space: 0 local, 1 non_local, 2 combined");
    B200_REQUIRE(m == 0 || (global_ids && local_ids && bounds && part_ids && starting), "null argument");
    B200_REQUIRE(index_space == kLocal || m == 0 || (bitmap && word_rank && range_offsets), "null argument");
    return launch_ew(ctx, m, 1, [=] __device__(int64_t i, int64_t) {
        const G gid = global_ids[i];
        L res = (L)-1;
        if (gid >= 0 && (int64_t)gid < global_size) {
            const int64_t r = find_range(bounds, num_ranges, gid);
            if (part_ids[r] == rank) {
                if (index_space != kNonLocal) res = (L)(gid - bounds[r]) + starting[r];
            } else if (index_space != kLocal && bit_set(bitmap, (int64_t)gid)) {
                const int64_t k = range_offsets[r] + rank_at(bitmap, word_rank, (int64_t)gid) -
                                  rank_at(bitmap, word_rank, (int64_t)bounds[r]);
                res = (L)(index_space == kCombined ? k + (int64_t)local_size : k);
            }
        }
        local_ids[i] = res;
    });
}

}  // namespace dist_assembly
}  // namespace b200

extern "C" {

b200_status b200_partition_count_ranges(b200_ctx* ctx, int64_t n, const int32_t* mapping,
                                        int64_t* num_ranges_host)
{
    return b200::dist_assembly::count_ranges(ctx, n, mapping, num_ranges_host);
}
b200_status b200_partition_has_ordered_parts(b200_ctx* ctx, int64_t num_ranges, const int32_t* part_ids,
                                             int32_t* result_host)
{
    return b200::dist_assembly::has_ordered_parts(ctx, num_ranges, part_ids, result_host);
}

#define B200_DEF_DIST_G(G, GT)                                                                          \
    b200_status b200_partition_build_ranges_from_global_size_##G(b200_ctx* ctx, int32_t num_parts,       \
                                                                 int64_t global_size, GT* ranges)        \
    {                                                                                                    \
        return b200::dist_assembly::build_ranges_from_global_size<GT>(ctx, num_parts, global_size,       \
                                                                      ranges);                           \
    }                                                                                                    \
    b200_status b200_partition_build_from_contiguous_##G(b200_ctx* ctx, int64_t num_ranges,              \
                                                         const GT* ranges,                               \
                                                         const int32_t* part_id_mapping,                 \
                                                         GT* range_bounds, int32_t* part_ids)            \
    {                                                                                                    \
        return b200::dist_assembly::build_from_contiguous<GT>(ctx, num_ranges, ranges, part_id_mapping,  \
                                                              range_bounds, part_ids);                   \
    }                                                                                                    \
    b200_status b200_partition_build_from_mapping_##G(b200_ctx* ctx, int64_t n, const int32_t* mapping,  \
                                                      GT* range_bounds, int32_t* part_ids)               \
    {                                                                                                    \
        return b200::dist_assembly::build_from_mapping<GT>(ctx, n, mapping, range_bounds, part_ids);     \
    }                                                                                                    \
    b200_status b200_dist_classify_entries_##G(                                                          \
        b200_ctx* ctx, int64_t nnz, const GT* row_idxs, const GT* col_idxs, int64_t row_num_ranges,      \
        const GT* row_bounds, const int32_t* row_part_ids, int64_t col_num_ranges, const GT* col_bounds, \
        const int32_t* col_part_ids, int32_t local_part, uint8_t* cls, int64_t* local_rank,              \
        int64_t* non_local_rank, int64_t* num_local_host, int64_t* num_non_local_host)                   \
    {                                                                                                    \
        return b200::dist_assembly::classify_entries<GT>(                                                \
            ctx, nnz, row_idxs, col_idxs, row_num_ranges, row_bounds, row_part_ids, col_num_ranges,      \
            col_bounds, col_part_ids, local_part, cls, local_rank, non_local_rank, num_local_host,       \
            num_non_local_host);                                                                         \
    }                                                                                                    \
    b200_status b200_index_map_mark_##G(b200_ctx* ctx, int64_t global_size, int64_t num_ranges,          \
                                        const GT* bounds, const int32_t* part_ids, int32_t skip_part,    \
                                        int64_t m, const GT* global_idxs, uint32_t* bitmap)              \
    {                                                                                                    \
        return b200::dist_assembly::index_map_mark<GT>(ctx, global_size, num_ranges, bounds, part_ids,   \
                                                       skip_part, m, global_idxs, bitmap);               \
    }                                                                                                    \
    b200_status b200_index_map_rank_##G(b200_ctx* ctx, int64_t global_size, int64_t num_ranges,          \
                                        int32_t num_parts, const GT* bounds, const int32_t* part_ids,    \
                                        const uint32_t* bitmap, int64_t* word_rank,                      \
                                        int64_t* range_offsets, int64_t* remote_sizes,                   \
                                        int64_t* num_remote_host)                                        \
    {                                                                                                    \
        return b200::dist_assembly::index_map_rank<GT>(ctx, global_size, num_ranges, num_parts, bounds,  \
                                                       part_ids, bitmap, word_rank, range_offsets,       \
                                                       remote_sizes, num_remote_host);                   \
    }
B200_DEF_DIST_G(i32, int32_t)
B200_DEF_DIST_G(i64, int64_t)

#define B200_DEF_DIST_LG(L, LT, G, GT)                                                                   \
    b200_status b200_partition_build_starting_indices_##L##_##G(                                         \
        b200_ctx* ctx, int64_t num_ranges, int32_t num_parts, const GT* range_bounds,                    \
        const int32_t* part_ids, LT* starting_indices, LT* part_sizes, int32_t* num_empty_parts_host)    \
    {                                                                                                    \
        return b200::dist_assembly::build_starting_indices<LT, GT>(ctx, num_ranges, num_parts,           \
                                                                   range_bounds, part_ids,               \
                                                                   starting_indices, part_sizes,         \
                                                                   num_empty_parts_host);                \
    }                                                                                                    \
    b200_status b200_index_map_fill_##L##_##G(                                                           \
        b200_ctx* ctx, int64_t global_size, int64_t num_ranges, const GT* bounds,                        \
        const int32_t* part_ids, const LT* starting, const uint32_t* bitmap, const int64_t* word_rank,   \
        const int64_t* range_offsets, GT* remote_global_idxs, LT* remote_local_idxs,                     \
        int32_t* remote_part_ids)                                                                        \
    {                                                                                                    \
        return b200::dist_assembly::index_map_fill<LT, GT>(ctx, global_size, num_ranges, bounds,         \
                                                           part_ids, starting, bitmap, word_rank,        \
                                                           range_offsets, remote_global_idxs,            \
                                                           remote_local_idxs, remote_part_ids);          \
    }                                                                                                    \
    b200_status b200_index_map_map_to_local_##L##_##G(                                                   \
        b200_ctx* ctx, int64_t global_size, int64_t num_ranges, const GT* bounds,                        \
        const int32_t* part_ids, const LT* starting, const uint32_t* bitmap, const int64_t* word_rank,   \
        const int64_t* range_offsets, int32_t rank, LT local_size, int32_t index_space, int64_t m,       \
        const GT* global_ids, LT* local_ids)                                                             \
    {                                                                                                    \
        return b200::dist_assembly::index_map_map_to_local<LT, GT>(                                      \
            ctx, global_size, num_ranges, bounds, part_ids, starting, bitmap, word_rank, range_offsets,  \
            rank, local_size, index_space, m, global_ids, local_ids);                                    \
    }
B200_DEF_DIST_LG(i32, int32_t, i32, int32_t)
B200_DEF_DIST_LG(i32, int32_t, i64, int64_t)
B200_DEF_DIST_LG(i64, int64_t, i64, int64_t)

#define B200_DEF_DIST_VLG(V, VT, L, LT, G, GT)                                                           \
    b200_status b200_dist_separate_fill_##V##_##L##_##G(                                                 \
        b200_ctx* ctx, int64_t nnz, const GT* row_idxs, const GT* col_idxs, const VT* values,            \
        int64_t row_num_ranges, const GT* row_bounds, const LT* row_starting, int64_t col_num_ranges,    \
        const GT* col_bounds, const LT* col_starting, const uint8_t* cls, const int64_t* local_rank,     \
        const int64_t* non_local_rank, LT* local_rows, LT* local_cols, VT* local_vals,                   \
        LT* non_local_rows, GT* non_local_cols, VT* non_local_vals)                                      \
    {                                                                                                    \
        return b200::dist_assembly::separate_fill<VT, LT, GT>(                                           \
            ctx, nnz, row_idxs, col_idxs, values, row_num_ranges, row_bounds, row_starting,              \
            col_num_ranges, col_bounds, col_starting, cls, local_rank, non_local_rank, local_rows,       \
            local_cols, local_vals, non_local_rows, non_local_cols, non_local_vals);                     \
    }                                                                                                    \
    b200_status b200_dist_vector_build_local_##V##_##L##_##G(                                            \
        b200_ctx* ctx, int64_t nnz, const GT* row_idxs, const GT* col_idxs, const VT* values,            \
        int64_t num_ranges, const GT* bounds, const int32_t* part_ids, const LT* starting,               \
        int32_t local_part, VT* local_values, int64_t local_stride)                                      \
    {                                                                                                    \
        return b200::dist_assembly::vector_build_local<VT, LT, GT>(ctx, nnz, row_idxs, col_idxs, values, \
                                                                   num_ranges, bounds, part_ids,         \
                                                                   starting, local_part, local_values,   \
                                                                   local_stride);                        \
    }                                                                                                    \
    b200_status b200_dist_kept_fill_##V##_##L##_##G(                                                     \
        b200_ctx* ctx, int64_t nnz, const GT* row_idxs, const GT* col_idxs, const VT* values,            \
        int64_t row_num_ranges, const GT* row_bounds, const LT* row_starting, const uint8_t* cls,        \
        const int64_t* local_rank, const int64_t* non_local_rank, LT* rows, GT* cols, VT* vals)          \
    {                                                                                                    \
        return b200::dist_assembly::kept_fill<VT, LT, GT>(ctx, nnz, row_idxs, col_idxs, values,          \
                                                          row_num_ranges, row_bounds, row_starting, cls, \
                                                          local_rank, non_local_rank, rows, cols, vals); \
    }
#define B200_DEF_DIST_VLG_ALL(V, VT)              \
    B200_DEF_DIST_VLG(V, VT, i32, int32_t, i32, int32_t) \
    B200_DEF_DIST_VLG(V, VT, i32, int32_t, i64, int64_t) \
    B200_DEF_DIST_VLG(V, VT, i64, int64_t, i64, int64_t)
B200_DEF_DIST_VLG_ALL(f64, double)
B200_DEF_DIST_VLG_ALL(f32, float)

}  // extern "C"
