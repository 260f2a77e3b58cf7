// gko_b200_types.hpp -- basic types and the exception hierarchy shared by the device-facing
// host layer (gko_b200.hpp) and the CUDA-free I/O layer (gko_b200_io.hpp).
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace gko_b200 {

using size_type = std::size_t;
using uint8 = std::uint8_t;
using uint32 = std::uint32_t;
using int32 = std::int32_t;
using int64 = std::int64_t;

struct dim2 {
    size_type rows = 0, cols = 0;
    dim2() = default;
    dim2(size_type r, size_type c) : rows(r), cols(c) {}
    explicit dim2(size_type n) : rows(n), cols(n) {}
    size_type operator[](int i) const { return i == 0 ? rows : cols; }
    bool operator==(const dim2& o) const { return rows == o.rows && cols == o.cols; }
};

// ---- exceptions (include/ginkgo/core/base/exception.hpp) -------------------------------
class Error : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};
class CudaError : public Error {
    using Error::Error;
};
class AllocationError : public Error {
    using Error::Error;
};
class DimensionMismatch : public Error {
    using Error::Error;
};
class NotSupported : public Error {
    using Error::Error;
};
class BadDimension : public Error {
    using Error::Error;
};

class OutOfBounds : public Error {
    using Error::Error;
};
class StreamError : public Error {
    using Error::Error;
};

}  // namespace gko_b200
