// jacobi_precision.cuh -- the storage types of the adaptive-precision block-Jacobi preconditioner on the
// device.  One byte per block (gko::precision_reduction, include/ginkgo/core/base/types.hpp:239-350:
// preserving << 4 | nonpreserving) selects how the inverted block is stored
// (GKO_PRECONDITIONER_JACOBI_RESOLVE_PRECISION, core/preconditioner/jacobi_utils.hpp:15-40, with
// reduce_precision / truncate_type of include/ginkgo/core/base/math.hpp:365-383, :546-582):
//   value type double: (0,1) float, (0,2) gko::half, (1,0) truncated<double,2> (upper 32 bits),
//                      (1,1) truncated<float,2> (upper 16 bits of the float), (2,0) truncated<double,4>
//   value type float:  (0,1), (0,2), (1,1) gko::half; (1,0), (2,0) truncated<float,2>
//   anything else: the value type.
// gko::half is NOT the IEEE conversion of the hardware (cvt.rn.f16.f32): denormal results are flushed to
// a signed zero, in both directions (include/ginkgo/core/base/half.hpp:399-448) -- restated here with
// integer operations so the stored bits match the reference's.
#pragma once
#include <stdint.h>

namespace b200 {
namespace jacobi {

enum StorageKind : int { kF64 = 0, kF32 = 1, kF16 = 2, kT64_32 = 3, kT64_16 = 4, kT32_16 = 5 };

template <typename V>
__host__ __device__ __forceinline__ int storage_kind(uint8_t prec);
template <>
__host__ __device__ __forceinline__ int storage_kind<double>(uint8_t prec)
{
    switch (prec) {
    case 0x01: return kF32;
    case 0x02: return kF16;
    case 0x10: return kT64_32;
    case 0x11: return kT32_16;
    case 0x20: return kT64_16;
    default: return kF64;
    }
}
template <>
__host__ __device__ __forceinline__ int storage_kind<float>(uint8_t prec)
{
    switch (prec) {
    case 0x01:
    case 0x02:
    case 0x11: return kF16;
    case 0x10:
    case 0x20: return kT32_16;
    default: return kF32;
    }
}
__host__ __device__ __forceinline__ int storage_bytes(int kind)
{
    return kind == kF64 ? 8 : ((kind == kF32 || kind == kT64_32) ? 4 : 2);
}

// half.hpp:399-432 (float2half)
__device__ __forceinline__ uint16_t float_to_gko_half(float f)
{
    const uint32_t d = __float_as_uint(f);
    const uint32_t sign = (d & 0x80000000u) >> 16;
    if ((d & 0x7f800000u) == 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((d & 0x007fffffu) ? 0x03ffu : 0u));
    const uint32_t e = (d & 0x7f800000u) >> 13;
    const uint32_t bias_change = (0x3f800000u >> 13) - 0x3c00u;
    if (e <= bias_change) return (uint16_t)sign;  // zero and everything that would be a half denormal
    uint32_t he = e - bias_change;
    if (he >= 0x7c00u) return (uint16_t)(sign | 0x7c00u);  // exponent overflow: infinity
    const uint32_t result = sign | he | ((d & 0x007fffffu) >> 13);
    const uint32_t tail = d & 0x1fffu;
    return (uint16_t)(result + ((tail > 0x1000u || (tail == 0x1000u && (result & 1u))) ? 1u : 0u));
}
// half.hpp:434-448 (half2float), branch-free
__device__ __forceinline__ float gko_half_to_float(uint16_t h)
{
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t ex = h & 0x7c00u, sig = h & 0x03ffu;
    const uint32_t normal = sign | (((uint32_t)ex << 13) + (0x3f800000u - (0x3c00u << 13))) | ((uint32_t)sig << 13);
    const uint32_t special = sign | 0x7f800000u | (sig ? 0x007fffffu : 0u);  // inf / nan
    return __uint_as_float(ex == 0x7c00u ? special : (ex == 0 ? sign : normal));
}

__device__ __forceinline__ void store_elem(void* base, int64_t idx, int kind, double v)
{
    switch (kind) {
    case kF64: reinterpret_cast<double*>(base)[idx] = v; break;
    case kF32: reinterpret_cast<float*>(base)[idx] = (float)v; break;
    case kF16: reinterpret_cast<uint16_t*>(base)[idx] = float_to_gko_half((float)v); break;
    case kT64_32: reinterpret_cast<uint32_t*>(base)[idx] = (uint32_t)__double2hiint(v); break;
    case kT64_16: reinterpret_cast<uint16_t*>(base)[idx] = (uint16_t)((uint32_t)__double2hiint(v) >> 16); break;
    default: reinterpret_cast<uint16_t*>(base)[idx] = (uint16_t)(__float_as_uint((float)v) >> 16); break;
    }
}
__device__ __forceinline__ void store_elem(void* base, int64_t idx, int kind, float v)
{
    switch (kind) {
    case kF32: reinterpret_cast<float*>(base)[idx] = v; break;
    case kF16: reinterpret_cast<uint16_t*>(base)[idx] = float_to_gko_half(v); break;
    default: reinterpret_cast<uint16_t*>(base)[idx] = (uint16_t)(__float_as_uint(v) >> 16); break;
    }
}
__device__ __forceinline__ double load_elem(const void* base, int64_t idx, int kind, double)
{
    switch (kind) {
    case kF64: return reinterpret_cast<const double*>(base)[idx];
    case kF32: return (double)reinterpret_cast<const float*>(base)[idx];
    case kF16: return (double)gko_half_to_float(reinterpret_cast<const uint16_t*>(base)[idx]);
    case kT64_32: return __hiloint2double((int)reinterpret_cast<const uint32_t*>(base)[idx], 0);
    case kT64_16: return __hiloint2double((int)((uint32_t) reinterpret_cast<const uint16_t*>(base)[idx] << 16), 0);
    default: return (double)__uint_as_float((uint32_t) reinterpret_cast<const uint16_t*>(base)[idx] << 16);
    }
}
__device__ __forceinline__ float load_elem(const void* base, int64_t idx, int kind, float)
{
    switch (kind) {
    case kF32: return reinterpret_cast<const float*>(base)[idx];
    case kF16: return gko_half_to_float(reinterpret_cast<const uint16_t*>(base)[idx]);
    default: return __uint_as_float((uint32_t) reinterpret_cast<const uint16_t*>(base)[idx] << 16);
    }
}
// value -> storage type -> value (validate_precision_reduction_feasibility's static_cast pair)
__device__ __forceinline__ double round_trip(double v, int kind)
{
    switch (kind) {
    case kF32: return (double)(float)v;
    case kF16: return (double)gko_half_to_float(float_to_gko_half((float)v));
    default: return v;
    }
}
__device__ __forceinline__ float round_trip(float v, int kind)
{
    return kind == kF16 ? gko_half_to_float(float_to_gko_half(v)) : v;
}

// precision_reduction_descriptor, get_optimal_storage_reduction (core/preconditioner/jacobi_utils.hpp:52-77,
// :171-189)
enum : uint32_t { kP0N0 = 0x00, kP0N2 = 0x01, kP1N1 = 0x02, kP2N0 = 0x04, kP0N1 = 0x08, kP1N0 = 0x10 };
__host__ __device__ __forceinline__ uint32_t prd_singleton(uint8_t pr)
{
    switch (pr) {
    case 0x01: return kP0N1;
    case 0x02: return kP0N2;
    case 0x10: return kP1N0;
    case 0x11: return kP1N1;
    case 0x20: return kP2N0;
    default: return kP0N0;
    }
}
__host__ __device__ __forceinline__ uint8_t optimal_reduction(uint32_t supported)
{
    if (supported & kP0N2) return 0x02;
    if (supported & kP1N1) return 0x11;
    if (supported & kP2N0) return 0x20;
    if (supported & kP0N1) return 0x01;
    if (supported & kP1N0) return 0x10;
    return 0x00;
}

}  // namespace jacobi
}  // namespace b200
