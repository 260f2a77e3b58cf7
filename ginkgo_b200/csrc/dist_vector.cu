// dense::compute_sqrt (core/matrix/dense_kernels.hpp, reference/matrix/dense_kernels.cpp:440-448):
// the last step of a distributed norm -- local squared norms, sum over the ranks, square root
// (core/distributed/vector.cpp:520-534).  Element-wise, bit-exact (IEEE square root).
#include "elementwise.cuh"

namespace b200 {
namespace dense {

template <typename V>
b200_status compute_sqrt(b200_ctx* ctx, int64_t rows, int64_t cols, V* data, int64_t stride)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_REQUIRE(rows >= 0 && cols >= 0 && stride >= cols, "bad shape");
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        data[i * stride + j] = sqrt(data[i * stride + j]);
    });
}

}  // namespace dense
}  // namespace b200

extern "C" {
b200_status b200_dense_compute_sqrt_f64(b200_ctx* ctx, int64_t rows, int64_t cols, double* data, int64_t stride)
{
    return b200::dense::compute_sqrt<double>(ctx, rows, cols, data, stride);
}
b200_status b200_dense_compute_sqrt_f32(b200_ctx* ctx, int64_t rows, int64_t cols, float* data, int64_t stride)
{
    return b200::dense::compute_sqrt<float>(ctx, rows, cols, data, stride);
}
}
