/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (ginkgo_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU
 * baseline.  See oracle_impl.h for what is restated and how it is pinned.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_precision.h"

/* jacobi::initialize_precisions, reference/preconditioner/jacobi_kernels.cpp:453-461 */
void orc_jacobi_initialize_precisions(const uint8_t* source, int64_t source_size, uint8_t* precisions,
                                      int64_t size)
{
    orc_jacobi_initialize_precisions_impl(source, source_size, precisions, size);
}
/* conversions of the adaptive block-Jacobi storage types, exported for the tests */
void orc_float_to_gko_half(const float* in, int64_t n, uint16_t* out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = orc_float_to_half(in[i]);
}
void orc_gko_half_to_float(const uint16_t* in, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = orc_half_to_float(in[i]);
}

/* reference/components/format_conversion_kernels.cpp: convert_ptrs_to_idxs / convert_idxs_to_ptrs */
#define ORC_CONVERT(IS, I)                                                                  \
    void orc_convert_ptrs_to_idxs_##IS(const I* ptrs, int64_t num_rows, I* idxs)            \
    {                                                                                       \
        for (int64_t r = 0; r < num_rows; ++r)                                              \
            for (int64_t k = ptrs[r]; k < (int64_t)ptrs[r + 1]; ++k) idxs[k] = (I)r;        \
    }                                                                                       \
    void orc_convert_idxs_to_ptrs_##IS(const I* idxs, int64_t nnz, int64_t num_rows, I* ptrs)\
    {                                                                                       \
        for (int64_t r = 0; r <= num_rows; ++r) ptrs[r] = 0;                                \
        for (int64_t k = 0; k < nnz; ++k) ptrs[idxs[k] + 1]++;                              \
        for (int64_t r = 0; r < num_rows; ++r) ptrs[r + 1] += ptrs[r];                      \
    }
ORC_CONVERT(i32, int32_t)
ORC_CONVERT(i64, int64_t)

static int orc_cmp_i64(const void* a, const void* b)
{
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* reference/matrix/ell_kernels.cpp:130-140 compute_max_row_nnz;
 * reference/matrix/sellp_kernels.cpp:107-130 compute_slice_sets (per-slice maximum rounded up
 * to stride_factor, then prefix_sum_nonnegative over num_slices + 1 entries);
 * core/matrix/csr.cpp:419-441 + reference `compute_hybrid_coo_row_ptrs`: prefix sum of
 * max(row_nnz - ell_lim, 0);
 * include/ginkgo/core/matrix/hybrid.hpp:222-243 imbalance_limit: std::sort of the row
 * lengths, value at position k. */
#define ORC_CONVERT_FMT(IS, I)                                                              \
    void orc_ell_compute_max_row_nnz_##IS(const I* ptrs, int64_t num_rows, int64_t* max_nnz) \
    {                                                                                       \
        *max_nnz = 0;                                                                       \
        for (int64_t i = 1; i <= num_rows; ++i) {                                           \
            const int64_t len = (int64_t)ptrs[i] - (int64_t)ptrs[i - 1];                    \
            if (len > *max_nnz) *max_nnz = len;                                             \
        }                                                                                   \
    }                                                                                       \
    void orc_sellp_compute_slice_sets_##IS(const I* ptrs, int64_t num_rows,                 \
                                           int64_t slice_size, int64_t stride_factor,       \
                                           uint64_t* slice_sets, uint64_t* slice_lengths)   \
    {                                                                                       \
        const int64_t num_slices = (num_rows + slice_size - 1) / slice_size;                \
        for (int64_t slice = 0; slice < num_slices; ++slice) {                              \
            uint64_t slice_length = 0;                                                      \
            for (int64_t lr = 0; lr < slice_size; ++lr) {                                   \
                const int64_t row = slice * slice_size + lr;                                \
                const int64_t len = row < num_rows ? (int64_t)ptrs[row + 1] - ptrs[row] : 0;\
                const uint64_t padded =                                                     \
                    (uint64_t)((len + stride_factor - 1) / stride_factor * stride_factor);  \
                if (padded > slice_length) slice_length = padded;                           \
            }                                                                               \
            slice_lengths[slice] = slice_length;                                            \
        }                                                                                   \
        uint64_t run = 0;                                                                   \
        for (int64_t slice = 0; slice < num_slices; ++slice) {                              \
            slice_sets[slice] = run;                                                        \
            run += slice_lengths[slice];                                                    \
        }                                                                                   \
        slice_sets[num_slices] = run;                                                       \
    }                                                                                       \
    void orc_csr_compute_hybrid_coo_row_ptrs_##IS(const I* ptrs, int64_t num_rows,          \
                                                  int64_t ell_lim, int64_t* coo_row_ptrs)   \
    {                                                                                       \
        int64_t run = 0;                                                                    \
        for (int64_t r = 0; r < num_rows; ++r) {                                            \
            const int64_t len = (int64_t)ptrs[r + 1] - (int64_t)ptrs[r];                    \
            coo_row_ptrs[r] = run;                                                          \
            run += len > ell_lim ? len - ell_lim : 0;                                       \
        }                                                                                   \
        coo_row_ptrs[num_rows] = run;                                                       \
    }                                                                                       \
    void orc_csr_row_nnz_order_statistic_##IS(const I* ptrs, int64_t num_rows, int64_t k,   \
                                              int64_t* value)                               \
    {                                                                                       \
        int64_t* len = (int64_t*)malloc(sizeof(int64_t) * (size_t)(num_rows > 0 ? num_rows : 1)); \
        for (int64_t r = 0; r < num_rows; ++r) len[r] = (int64_t)ptrs[r + 1] - (int64_t)ptrs[r]; \
        qsort(len, (size_t)num_rows, sizeof(int64_t), orc_cmp_i64);                         \
        *value = len[k];                                                                    \
        free(len);                                                                          \
    }
ORC_CONVERT_FMT(i32, int32_t)
ORC_CONVERT_FMT(i64, int64_t)

/* reference/preconditioner/jacobi_kernels.cpp:36-123: has_same_nonzero_pattern,
 * find_natural_blocks, agglomerate_supervariables, find_blocks */
#define ORC_FIND_BLOCKS(IS, I)                                                              \
    void orc_jacobi_find_blocks_##IS(int64_t rows, const I* row_ptrs, const I* col_idx,     \
                                     int32_t max_block_size, I* block_ptrs,                 \
                                     int64_t* num_blocks_out)                               \
    {                                                                                       \
        block_ptrs[0] = 0;                                                                  \
        *num_blocks_out = 0;                                                                \
        if (rows == 0) return;                                                              \
        int64_t num_blocks = 1;                                                             \
        int32_t current = 1;                                                                \
        for (int64_t i = 1; i < rows; ++i) {                                                \
            const I* prev = col_idx + row_ptrs[i - 1];                                      \
            const I* curr = col_idx + row_ptrs[i];                                          \
            const I* next = col_idx + row_ptrs[i + 1];                                      \
            int same = (next - curr) == (curr - prev);                                      \
            for (int64_t k = 0; same && k < next - curr; ++k) same = curr[k] == prev[k];    \
            if (current < max_block_size && same) {                                         \
                ++current;                                                                  \
            } else {                                                                        \
                block_ptrs[num_blocks] = block_ptrs[num_blocks - 1] + current;              \
                ++num_blocks;                                                               \
                current = 1;                                                                \
            }                                                                               \
        }                                                                                   \
        block_ptrs[num_blocks] = block_ptrs[num_blocks - 1] + current;                      \
        const int64_t num_natural = num_blocks;                                             \
        num_blocks = 1;                                                                     \
        I cur = block_ptrs[1] - block_ptrs[0];                                              \
        for (int64_t i = 1; i < num_natural; ++i) {                                         \
            const I size = block_ptrs[i + 1] - block_ptrs[i];                               \
            if (cur + size <= max_block_size) {                                             \
                cur += size;                                                                \
            } else {                                                                        \
                block_ptrs[num_blocks] = block_ptrs[i];                                     \
                ++num_blocks;                                                               \
                cur = size;                                                                 \
            }                                                                               \
        }                                                                                   \
        block_ptrs[num_blocks] = block_ptrs[num_natural];                                   \
        *num_blocks_out = num_blocks;                                                       \
    }
ORC_FIND_BLOCKS(i32, int32_t)
ORC_FIND_BLOCKS(i64, int64_t)

/* ---- distributed set-up (partition, separate_local_nonlocal, index_map) ---- */
#define G int32_t
#define GS i32
#include "oracle_dist.h"
#define L int32_t
#define LS i32
#include "oracle_dist.h"
#define V double
#define VS f64
#include "oracle_dist.h"
#undef V
#undef VS
#define V float
#define VS f32
#include "oracle_dist.h"
#undef V
#undef VS
#undef L
#undef LS
#undef G
#undef GS
#define G int64_t
#define GS i64
#include "oracle_dist.h"
#define L int32_t
#define LS i32
#include "oracle_dist.h"
#define V double
#define VS f64
#include "oracle_dist.h"
#undef V
#undef VS
#define V float
#define VS f32
#include "oracle_dist.h"
#undef V
#undef VS
#undef L
#undef LS
#define L int64_t
#define LS i64
#include "oracle_dist.h"
#define V double
#define VS f64
#include "oracle_dist.h"
#undef V
#undef VS
#define V float
#define VS f32
#include "oracle_dist.h"
#undef V
#undef VS
#undef L
#undef LS
#undef G
#undef GS

/* ---- double ---- */
#define V double
#define VS f64
#define SQRT sqrt
#define FABS fabs
#include "oracle_impl.h"
#define ORC_VALUE_PART_DONE
#define I int32_t
#define IS i32
#include "oracle_impl.h"
#undef I
#undef IS
#define I int64_t
#define IS i64
#include "oracle_impl.h"
#undef I
#undef IS
#undef ORC_VALUE_PART_DONE
#include "oracle_solvers.h"
#undef V
#undef VS
#undef SQRT
#undef FABS

/* ---- float ---- */
#define V float
#define VS f32
#define SQRT sqrtf
#define FABS fabsf
#include "oracle_impl.h"
#define ORC_VALUE_PART_DONE
#define I int32_t
#define IS i32
#include "oracle_impl.h"
#undef I
#undef IS
#define I int64_t
#define IS i64
#include "oracle_impl.h"
#undef I
#undef IS
#undef ORC_VALUE_PART_DONE
#include "oracle_solvers.h"
