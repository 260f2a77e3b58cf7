/*
 * oracle_precision.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The storage types of the adaptive-precision block-Jacobi preconditioner, restated in plain C:
 *   gko::half            include/ginkgo/core/base/half.hpp:399-448 (float2half: denormals flushed to a
 *                        signed zero, exponent overflow -> infinity, significand rounded to nearest even;
 *                        half2float: denormals -> signed zero)
 *   gko::truncated<T,N>  core/base/extended_float.hpp:52-116 (component 0 = the upper bits of T, the rest
 *                        is dropped; reading it back fills zeros)
 *   precision_reduction  include/ginkgo/core/base/types.hpp:239-350 (one byte: preserving << 4 | nonpreserving)
 *   resolution           core/preconditioner/jacobi_utils.hpp:15-40 (GKO_PRECONDITIONER_JACOBI_RESOLVE_PRECISION)
 *                        with reduce_precision / truncate_type of include/ginkgo/core/base/math.hpp:365-383,
 *                        :546-582:   double: (0,1) float, (0,2) half, (1,0) truncated<double,2>,
 *                        (1,1) truncated<float,2>, (2,0) truncated<double,4>;   float: (0,1) half,
 *                        (0,2) half, (1,0) truncated<float,2>, (1,1) half, (2,0) truncated<float,2>
 *                        (truncate_type never goes below 16 bits); anything else: the value type.
 */
#ifndef ORACLE_PRECISION_H
#define ORACLE_PRECISION_H
#include <stdint.h>
#include <string.h>

/* storage kinds */
enum {
    ORC_ST_F64 = 0,
    ORC_ST_F32 = 1,
    ORC_ST_F16 = 2,     /* gko::half */
    ORC_ST_T64_32 = 3,  /* truncated<double, 2, 0>: upper 32 bits of a double */
    ORC_ST_T64_16 = 4,  /* truncated<double, 4, 0>: upper 16 bits of a double */
    ORC_ST_T32_16 = 5   /* truncated<float, 2, 0>: upper 16 bits of a float */
};

static inline int orc_storage_kind(int value_is_double, uint8_t prec)
{
    if (value_is_double) {
        switch (prec) {
        case 0x01: return ORC_ST_F32;
        case 0x02: return ORC_ST_F16;
        case 0x10: return ORC_ST_T64_32;
        case 0x11: return ORC_ST_T32_16;
        case 0x20: return ORC_ST_T64_16;
        default: return ORC_ST_F64;
        }
    }
    switch (prec) {
    case 0x01: return ORC_ST_F16;
    case 0x02: return ORC_ST_F16;
    case 0x10: return ORC_ST_T32_16;
    case 0x11: return ORC_ST_F16;
    case 0x20: return ORC_ST_T32_16;
    default: return ORC_ST_F32;
    }
}

static inline int orc_storage_bytes(int kind)
{
    switch (kind) {
    case ORC_ST_F64: return 8;
    case ORC_ST_F32:
    case ORC_ST_T64_32: return 4;
    default: return 2;
    }
}

/* half.hpp:399-432 */
static inline uint16_t orc_float2half_bits(uint32_t d)
{
    const uint32_t sign = (d & 0x80000000u) >> 16;
    const uint32_t exp_mask = 0x7f800000u, sig_mask = 0x007fffffu;
    if ((d & exp_mask) == exp_mask && (d & sig_mask) == 0) return (uint16_t)(sign | 0x7c00u);
    if ((d & exp_mask) == exp_mask) return (uint16_t)(sign | 0x7c00u | 0x03ffu);
    /* shift_exponent: (exp field >> 13), minus the bias change, clamped */
    const uint32_t e = (d & exp_mask) >> 13;                 /* still carries the float bias */
    const uint32_t bias_change = (0x3f800000u >> 13) - 0x3c00u; /* (127 - 15) << 10 */
    uint32_t he;
    if (e <= bias_change)
        he = 0;
    else {
        he = e - bias_change;
        if (he >= 0x7c00u) he = 0x7c00u;
    }
    if ((he & 0x7c00u) == 0x7c00u && (he & 0x03ffu) == 0) return (uint16_t)(sign | he); /* is_inf */
    if ((he & 0x7c00u) == 0) return (uint16_t)sign;                                     /* is_denom */
    const uint32_t result = sign | he | ((d & sig_mask) >> 13);
    const uint32_t tail = d & 0x1fffu;
    const uint32_t half = 0x1000u;
    return (uint16_t)(result + ((tail > half || (tail == half && (result & 1u))) ? 1u : 0u));
}

/* half.hpp:434-448 */
static inline uint32_t orc_half2float_bits(uint16_t h)
{
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    if ((h & 0x7c00u) == 0x7c00u && (h & 0x03ffu) == 0) return sign | 0x7f800000u;
    if ((h & 0x7c00u) == 0x7c00u) return sign | 0x7f800000u | 0x007fffffu;
    if ((h & 0x7c00u) == 0) return sign;
    const uint32_t e = ((uint32_t)(h & 0x7c00u) << 13) + (0x3f800000u - (0x3c00u << 13));
    return sign | e | ((uint32_t)(h & 0x03ffu) << 13);
}

static inline uint16_t orc_float_to_half(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return orc_float2half_bits(u);
}
static inline float orc_half_to_float(uint16_t h)
{
    const uint32_t u = orc_half2float_bits(h);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* store `v` (given as the value type: double or float) into slot `idx` of a block stored as `kind` */
static inline void orc_store_f64(void* base, int64_t idx, int kind, double v)
{
    uint64_t u;
    uint32_t w;
    float f;
    switch (kind) {
    case ORC_ST_F64: ((double*)base)[idx] = v; break;
    case ORC_ST_F32: ((float*)base)[idx] = (float)v; break;
    case ORC_ST_F16: ((uint16_t*)base)[idx] = orc_float_to_half((float)v); break;
    case ORC_ST_T64_32:
        memcpy(&u, &v, 8);
        ((uint32_t*)base)[idx] = (uint32_t)(u >> 32);
        break;
    case ORC_ST_T64_16:
        memcpy(&u, &v, 8);
        ((uint16_t*)base)[idx] = (uint16_t)(u >> 48);
        break;
    default: /* ORC_ST_T32_16: double -> float (round to nearest), then the upper 16 bits */
        f = (float)v;
        memcpy(&w, &f, 4);
        ((uint16_t*)base)[idx] = (uint16_t)(w >> 16);
        break;
    }
}
static inline double orc_load_f64(const void* base, int64_t idx, int kind)
{
    uint64_t u;
    uint32_t w;
    double d;
    float f;
    switch (kind) {
    case ORC_ST_F64: return ((const double*)base)[idx];
    case ORC_ST_F32: return (double)((const float*)base)[idx];
    case ORC_ST_F16: return (double)orc_half_to_float(((const uint16_t*)base)[idx]);
    case ORC_ST_T64_32:
        u = (uint64_t)((const uint32_t*)base)[idx] << 32;
        memcpy(&d, &u, 8);
        return d;
    case ORC_ST_T64_16:
        u = (uint64_t)((const uint16_t*)base)[idx] << 48;
        memcpy(&d, &u, 8);
        return d;
    default:
        w = (uint32_t)((const uint16_t*)base)[idx] << 16;
        memcpy(&f, &w, 4);
        return (double)f;
    }
}
static inline void orc_store_f32(void* base, int64_t idx, int kind, float v)
{
    uint32_t w;
    switch (kind) {
    case ORC_ST_F32: ((float*)base)[idx] = v; break;
    case ORC_ST_F16: ((uint16_t*)base)[idx] = orc_float_to_half(v); break;
    default: /* ORC_ST_T32_16 */
        memcpy(&w, &v, 4);
        ((uint16_t*)base)[idx] = (uint16_t)(w >> 16);
        break;
    }
}
static inline float orc_load_f32(const void* base, int64_t idx, int kind)
{
    uint32_t w;
    float f;
    switch (kind) {
    case ORC_ST_F32: return ((const float*)base)[idx];
    case ORC_ST_F16: return orc_half_to_float(((const uint16_t*)base)[idx]);
    default:
        w = (uint32_t)((const uint16_t*)base)[idx] << 16;
        memcpy(&f, &w, 4);
        return f;
    }
}

/* precision_reduction_descriptor, core/preconditioner/jacobi_utils.hpp:52-77 */
enum { ORC_P0N0 = 0x00, ORC_P0N2 = 0x01, ORC_P1N1 = 0x02, ORC_P2N0 = 0x04, ORC_P0N1 = 0x08, ORC_P1N0 = 0x10 };
static inline uint32_t orc_prd_singleton(uint8_t pr)
{
    switch (pr) {
    case 0x01: return ORC_P0N1;
    case 0x02: return ORC_P0N2;
    case 0x10: return ORC_P1N0;
    case 0x11: return ORC_P1N1;
    case 0x20: return ORC_P2N0;
    default: return ORC_P0N0;
    }
}
/* get_optimal_storage_reduction, jacobi_utils.hpp:171-189 */
static inline uint8_t orc_optimal_reduction(uint32_t supported)
{
    if (supported & ORC_P0N2) return 0x02;
    if (supported & ORC_P1N1) return 0x11;
    if (supported & ORC_P2N0) return 0x20;
    if (supported & ORC_P0N1) return 0x01;
    if (supported & ORC_P1N0) return 0x10;
    return 0x00;
}

/* jacobi::initialize_precisions, reference/preconditioner/jacobi_kernels.cpp:453-461 */
static inline void orc_jacobi_initialize_precisions_impl(const uint8_t* source, int64_t source_size,
                                                         uint8_t* precisions, int64_t size)
{
    for (int64_t i = 0; i < size; ++i) precisions[i] = source[i % source_size];
}
#endif
