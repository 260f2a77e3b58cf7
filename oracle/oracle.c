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
