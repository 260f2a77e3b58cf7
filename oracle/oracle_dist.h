/* TEST INFRASTRUCTURE ONLY -- sequential restatement of the reference's distributed set-up
 * kernels, included by oracle.c:
 *   partition::*                              reference/distributed/partition_kernels.cpp:18-160
 *   find_range / map_to_local                 reference/distributed/partition_helpers.hpp:24-56
 *   distributed_matrix::separate_local_nonlocal  reference/distributed/matrix_kernels.cpp:18-90
 *   index_map::build_mapping / map_to_local   reference/distributed/index_map_kernels.cpp:20-212
 * The entry points have the argument lists of the b200_* functions of include/ginkgo_b200.h
 * (minus the context).  separate_local_nonlocal is split into classify + fill and the index
 * map keeps its set of remote indices as a bitmap over the global index space with a
 * per-word rank; everything the reference returns is computed here the way the reference
 * computes it (sequential walk, sort + unique by (part id, global index)). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- index-type independent */
#ifndef ORC_DIST_COMMON
#define ORC_DIST_COMMON
/* partition::count_ranges (:18-29) */
void orc_partition_count_ranges(int64_t n, const int32_t* mapping, int64_t* num_ranges_host)
{
    int64_t num = 0;
    int32_t prev = -1;
    for (int64_t i = 0; i < n; ++i) {
        num += mapping[i] != prev;
        prev = mapping[i];
    }
    *num_ranges_host = num;
}

/* partition::has_ordered_parts (:138-153) */
void orc_partition_has_ordered_parts(int64_t num_ranges, const int32_t* part_ids, int32_t* result_host)
{
    *result_host = 1;
    for (int64_t i = 1; i < num_ranges; ++i)
        if (part_ids[i] < part_ids[i - 1]) {
            *result_host = 0;
            return;
        }
}

static int orc_popc32(uint32_t w)
{
    int c = 0;
    for (; w; w &= w - 1) ++c;
    return c;
}
#endif

/* ---------------------------------------------------------------- per global index type */
#if defined(G) && !defined(L) && !defined(V)
#define ORC_CAT3(a, b, c) a##b##c
#define ORC_GN2(name, gs) ORC_CAT3(name, _, gs)
#define ORC_GN(name) ORC_GN2(name, GS)

/* std::upper_bound over range_bounds[1 .. num_ranges] (partition_helpers.hpp:24-40) */
static int64_t ORC_GN(orc_find_range)(const G* bounds, int64_t num_ranges, G idx)
{
    int64_t lo = 0, hi = num_ranges; /* search in bounds + 1 */
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (bounds[1 + mid] <= idx)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* partition::build_ranges_from_global_size (:75-92) */
void ORC_GN(orc_partition_build_ranges_from_global_size)(int32_t num_parts, int64_t global_size, G* ranges)
{
    ranges[0] = 0;
    if (num_parts == 0) return;
    const int64_t size_per_part = global_size / num_parts;
    const int64_t rest = global_size - num_parts * size_per_part;
    for (int i = 1; i < num_parts + 1; ++i)
        ranges[i] = (G)(ranges[i - 1] + size_per_part + ((i - 1) < rest ? 1 : 0));
}

/* partition::build_from_contiguous (:32-47) */
void ORC_GN(orc_partition_build_from_contiguous)(int64_t num_ranges, const G* ranges,
                                                 const int32_t* part_id_mapping, G* range_bounds,
                                                 int32_t* part_ids)
{
    range_bounds[0] = 0;
    for (int64_t i = 0; i < num_ranges; ++i) {
        range_bounds[i + 1] = ranges[i + 1];
        part_ids[i] = part_id_mapping ? part_id_mapping[i] : (int32_t)i;
    }
}

/* partition::build_from_mapping (:52-70) */
void ORC_GN(orc_partition_build_from_mapping)(int64_t n, const int32_t* mapping, G* range_bounds,
                                              int32_t* part_ids)
{
    int64_t range_idx = 0;
    int32_t range_part = -1;
    for (int64_t i = 0; i < n; ++i) {
        if (mapping[i] != range_part) {
            range_bounds[range_idx] = (G)i;
            part_ids[range_idx] = mapping[i];
            ++range_idx;
            range_part = mapping[i];
        }
    }
    range_bounds[range_idx] = (G)n;
}

/* distributed_matrix::separate_local_nonlocal, the walk of :43-65 as a classification:
 * cls 0 = row not owned, 1 = local entry, 2 = non-local entry; *_rank = exclusive counts */
void ORC_GN(orc_dist_classify_entries)(int64_t nnz, const G* row_idxs, const G* col_idxs,
                                       int64_t row_num_ranges, const G* row_bounds,
                                       const int32_t* row_part_ids, int64_t col_num_ranges,
                                       const G* col_bounds, const int32_t* col_part_ids,
                                       int32_t local_part, uint8_t* cls, int64_t* local_rank,
                                       int64_t* non_local_rank, int64_t* num_local_host,
                                       int64_t* num_non_local_host)
{
    int64_t nl = 0, nn = 0;
    for (int64_t i = 0; i < nnz; ++i) {
        local_rank[i] = nl;
        non_local_rank[i] = nn;
        const int64_t rr = ORC_GN(orc_find_range)(row_bounds, row_num_ranges, row_idxs[i]);
        uint8_t c = 0;
        if (row_part_ids[rr] == local_part) {
            const int64_t cr = ORC_GN(orc_find_range)(col_bounds, col_num_ranges, col_idxs[i]);
            c = col_part_ids[cr] == local_part ? 1 : 2;
        }
        cls[i] = c;
        nl += c == 1;
        nn += c == 2;
    }
    local_rank[nnz] = nl;
    non_local_rank[nnz] = nn;
    *num_local_host = nl;
    *num_non_local_host = nn;
}

/* the index map's set of remote indices: bit g of the bitmap <=> g is connected.  skip_part
 * >= 0 ignores indices owned by that part (so all columns of the kept entries can be passed) */
void ORC_GN(orc_index_map_mark)(int64_t global_size, int64_t num_ranges, const G* bounds,
                                const int32_t* part_ids, int32_t skip_part, int64_t m,
                                const G* global_idxs, uint32_t* bitmap)
{
    const int64_t words = (global_size + 31) / 32 + 1;
    memset(bitmap, 0, sizeof(uint32_t) * (size_t)words);
    for (int64_t i = 0; i < m; ++i) {
        const G g = global_idxs[i];
        if (g < 0 || (int64_t)g >= global_size) continue;
        if (skip_part >= 0 && part_ids[ORC_GN(orc_find_range)(bounds, num_ranges, g)] == skip_part) continue;
        bitmap[g >> 5] |= 1u << (g & 31);
    }
}

/* sizes and offsets of index_map::build_mapping (:39-62 sort + unique by (part, global),
 * :92-101 sizes per part): word_rank = exclusive count of set bits before each word;
 * range_offsets[r] = position of range r's first remote index in the sorted unique list */
void ORC_GN(orc_index_map_rank)(int64_t global_size, int64_t num_ranges, int32_t num_parts,
                                const G* bounds, const int32_t* part_ids, const uint32_t* bitmap,
                                int64_t* word_rank, int64_t* range_offsets, int64_t* remote_sizes,
                                int64_t* num_remote_host)
{
    const int64_t words = (global_size + 31) / 32;
    int64_t run = 0;
    for (int64_t w = 0; w < words; ++w) {
        word_rank[w] = run;
        run += orc_popc32(bitmap[w]);
    }
    word_rank[words] = run;
    *num_remote_host = run;
    for (int32_t p = 0; p < num_parts; ++p) remote_sizes[p] = 0;
    /* count per range by walking the set bits (the reference counts part ids of the list) */
    int64_t* count = (int64_t*)calloc((size_t)(num_ranges > 0 ? num_ranges : 1), sizeof(int64_t));
    for (int64_t g = 0; g < global_size; ++g)
        if (bitmap[g >> 5] >> (g & 31) & 1u) count[ORC_GN(orc_find_range)(bounds, num_ranges, (G)g)]++;
    for (int64_t r = 0; r < num_ranges; ++r) remote_sizes[part_ids[r]] += count[r];
    /* (part, global) order: parts ascending, inside a part ranges ascending */
    int64_t pos = 0;
    for (int32_t p = 0; p < num_parts; ++p)
        for (int64_t r = 0; r < num_ranges; ++r)
            if (part_ids[r] == p) {
                range_offsets[r] = pos;
                pos += count[r];
            }
    free(count);
}
#undef ORC_GN
#undef ORC_GN2
#undef ORC_CAT3
#endif

/* ---------------------------------------------------------------- per (local, global) */
#if defined(G) && defined(L) && !defined(V)
#define ORC_CAT5(a, b, c, d, e) a##b##c##d##e
#define ORC_LGN2(name, ls, gs) ORC_CAT5(name, _, ls, _, gs)
#define ORC_LGN(name) ORC_LGN2(name, LS, GS)
#define ORC_CAT3(a, b, c) a##b##c
#define ORC_GN2(name, gs) ORC_CAT3(name, _, gs)
#define ORC_GN(name) ORC_GN2(name, GS)

/* partition::build_starting_indices (:97-113) */
void ORC_LGN(orc_partition_build_starting_indices)(int64_t num_ranges, int32_t num_parts,
                                                   const G* range_bounds, const int32_t* part_ids,
                                                   L* starting_indices, L* part_sizes,
                                                   int32_t* num_empty_parts_host)
{
    for (int32_t p = 0; p < num_parts; ++p) part_sizes[p] = 0;
    for (int64_t r = 0; r < num_ranges; ++r) {
        const int32_t part = part_ids[r];
        starting_indices[r] = part_sizes[part];
        part_sizes[part] += (L)(range_bounds[r + 1] - range_bounds[r]);
    }
    int32_t empty = 0;
    for (int32_t p = 0; p < num_parts; ++p) empty += part_sizes[p] == 0;
    *num_empty_parts_host = empty;
}

typedef struct {
    int32_t part;
    G gid;
} ORC_LGN(orc_pg);
static int ORC_LGN(orc_pg_cmp)(const void* a, const void* b)
{
    const ORC_LGN(orc_pg)* x = (const ORC_LGN(orc_pg)*)a;
    const ORC_LGN(orc_pg)* y = (const ORC_LGN(orc_pg)*)b;
    if (x->part != y->part) return x->part < y->part ? -1 : 1;
    return x->gid < y->gid ? -1 : (x->gid > y->gid);
}

/* index_map::build_mapping (:20-105): the connected indices, sorted and made unique by
 * (part id, global index); remote_local_idxs = map_to_local of each; remote_part_ids (may be
 * NULL) = the owner of each entry (the reference's full_part_ids) */
void ORC_LGN(orc_index_map_fill)(int64_t global_size, int64_t num_ranges, const G* bounds,
                                 const int32_t* part_ids, const L* starting, const uint32_t* bitmap,
                                 const int64_t* word_rank, const int64_t* range_offsets,
                                 G* remote_global_idxs, L* remote_local_idxs, int32_t* remote_part_ids)
{
    (void)word_rank;
    (void)range_offsets;
    int64_t n = 0;
    for (int64_t g = 0; g < global_size; ++g) n += bitmap[g >> 5] >> (g & 31) & 1u;
    ORC_LGN(orc_pg)* list = (ORC_LGN(orc_pg)*)malloc(sizeof(ORC_LGN(orc_pg)) * (size_t)(n > 0 ? n : 1));
    int64_t k = 0;
    for (int64_t g = 0; g < global_size; ++g)
        if (bitmap[g >> 5] >> (g & 31) & 1u) {
            list[k].gid = (G)g;
            list[k].part = part_ids[ORC_GN(orc_find_range)(bounds, num_ranges, (G)g)];
            ++k;
        }
    qsort(list, (size_t)n, sizeof(list[0]), ORC_LGN(orc_pg_cmp));
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = ORC_GN(orc_find_range)(bounds, num_ranges, list[i].gid);
        remote_global_idxs[i] = list[i].gid;
        remote_local_idxs[i] = (L)(list[i].gid - bounds[r]) + starting[r];
        if (remote_part_ids) remote_part_ids[i] = list[i].part;
    }
    free(list);
}

/* index_map::map_to_local (:108-212); index_space 0 = local, 1 = non_local, 2 = combined;
 * the non-local index of a connected gid is its position in the (part, global)-sorted list */
void ORC_LGN(orc_index_map_map_to_local)(int64_t global_size, int64_t num_ranges, const G* bounds,
                                         const int32_t* part_ids, const L* starting,
                                         const uint32_t* bitmap, const int64_t* word_rank,
                                         const int64_t* range_offsets, int32_t rank, L local_size,
                                         int32_t index_space, int64_t m, const G* global_ids,
                                         L* local_ids)
{
    (void)word_rank;
    for (int64_t i = 0; i < m; ++i) {
        const G gid = global_ids[i];
        L res = (L)-1;
        if (gid >= 0 && (int64_t)gid < global_size) {
            const int64_t r = ORC_GN(orc_find_range)(bounds, num_ranges, gid);
            const int is_local = part_ids[r] == rank;
            L loc = (L)-1, nloc = (L)-1;
            if (is_local) loc = (L)(gid - bounds[r]) + starting[r];
            if (!is_local && (bitmap[gid >> 5] >> (gid & 31) & 1u)) {
                int64_t before = 0; /* connected indices of this range below gid */
                for (int64_t g = (int64_t)bounds[r]; g < (int64_t)gid; ++g)
                    before += bitmap[g >> 5] >> (g & 31) & 1u;
                nloc = (L)(range_offsets[r] + before);
            }
            if (index_space == 0) res = loc;
            if (index_space == 1) res = nloc;
            if (index_space == 2) res = is_local ? loc : (nloc == (L)-1 ? nloc : (L)(nloc + local_size));
        }
        local_ids[i] = res;
    }
}
#undef ORC_LGN
#undef ORC_LGN2
#undef ORC_CAT5
#undef ORC_GN
#undef ORC_GN2
#undef ORC_CAT3
#endif

/* ---------------------------------------------------------------- per (value, local, global) */
#if defined(G) && defined(L) && defined(V)
#define ORC_CAT7(a, b, c, d, e, f, g) a##b##c##d##e##f##g
#define ORC_VLGN2(name, vs, ls, gs) ORC_CAT7(name, _, vs, _, ls, _, gs)
#define ORC_VLGN(name) ORC_VLGN2(name, VS, LS, GS)
#define ORC_CAT3(a, b, c) a##b##c
#define ORC_GN2(name, gs) ORC_CAT3(name, _, gs)
#define ORC_GN(name) ORC_GN2(name, GS)

/* distributed_matrix::separate_local_nonlocal (:43-88): the six output arrays */
void ORC_VLGN(orc_dist_separate_fill)(int64_t nnz, const G* row_idxs, const G* col_idxs,
                                      const V* values, int64_t row_num_ranges, const G* row_bounds,
                                      const L* row_starting, int64_t col_num_ranges,
                                      const G* col_bounds, const L* col_starting, const uint8_t* cls,
                                      const int64_t* local_rank, const int64_t* non_local_rank,
                                      L* local_rows, L* local_cols, V* local_vals, L* non_local_rows,
                                      G* non_local_cols, V* non_local_vals)
{
    for (int64_t i = 0; i < nnz; ++i) {
        if (!cls[i]) continue;
        const int64_t rr = ORC_GN(orc_find_range)(row_bounds, row_num_ranges, row_idxs[i]);
        const L lrow = (L)(row_idxs[i] - row_bounds[rr]) + row_starting[rr];
        if (cls[i] == 1) {
            const int64_t cr = ORC_GN(orc_find_range)(col_bounds, col_num_ranges, col_idxs[i]);
            const int64_t k = local_rank[i];
            local_rows[k] = lrow;
            local_cols[k] = (L)(col_idxs[i] - col_bounds[cr]) + col_starting[cr];
            local_vals[k] = values[i];
        } else {
            const int64_t k = non_local_rank[i];
            non_local_rows[k] = lrow;
            non_local_cols[k] = col_idxs[i];
            non_local_vals[k] = values[i];
        }
    }
}

/* all entries of the owned rows in input order, rows local, columns still global (what the
 * combined-index-space matrix is assembled from) */
void ORC_VLGN(orc_dist_kept_fill)(int64_t nnz, const G* row_idxs, const G* col_idxs, const V* values,
                                  int64_t row_num_ranges, const G* row_bounds, const L* row_starting,
                                  const uint8_t* cls, const int64_t* local_rank,
                                  const int64_t* non_local_rank, L* rows, G* cols, V* vals)
{
    for (int64_t i = 0; i < nnz; ++i) {
        if (!cls[i]) continue;
        const int64_t rr = ORC_GN(orc_find_range)(row_bounds, row_num_ranges, row_idxs[i]);
        const int64_t k = local_rank[i] + non_local_rank[i];
        rows[k] = (L)(row_idxs[i] - row_bounds[rr]) + row_starting[rr];
        cols[k] = col_idxs[i];
        vals[k] = values[i];
    }
}

/* distributed_vector::build_local (reference/distributed/vector_kernels.cpp:15-40): the entries
 * of the owned rows into the (pre-zeroed) row-major local block; later duplicates win */
void ORC_VLGN(orc_dist_vector_build_local)(int64_t nnz, const G* row_idxs, const G* col_idxs, const V* values,
                                           int64_t num_ranges, const G* bounds, const int32_t* part_ids,
                                           const L* starting, int32_t local_part, V* local_values,
                                           int64_t local_stride)
{
    for (int64_t i = 0; i < nnz; ++i) {
        const int64_t r = ORC_GN(orc_find_range)(bounds, num_ranges, row_idxs[i]);
        if (part_ids[r] != local_part) continue;
        const int64_t lrow = (int64_t)((L)(row_idxs[i] - bounds[r]) + starting[r]);
        local_values[lrow * local_stride + (int64_t)col_idxs[i]] = values[i];
    }
}
#undef ORC_VLGN
#undef ORC_VLGN2
#undef ORC_CAT7
#undef ORC_GN
#undef ORC_GN2
#undef ORC_CAT3
#endif
