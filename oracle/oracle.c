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
