// TEST INFRASTRUCTURE ONLY -- stands in for ginkgo_b200/csrc/scan.cuh when a copy of an
// element-wise .cu file is compiled for the host (see elementwise.cuh next to this file): the
// three-launch device scan becomes the sequential loop it must agree with.  The device scan
// itself is covered by the GPU tests of its other users (conversions, find_blocks).
#pragma once
#include "elementwise.cuh"

namespace b200 {
namespace scan {

constexpr int kTile = 2048;
inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }

template <typename T, typename Load>
inline b200_status exclusive(b200_ctx* ctx, int64_t n, Load load, T* out, T*)
{
    T run = T(0);
    for (int64_t i = 0; i < n; ++i) {
        const T v = load(i);
        out[i] = run;
        run += v;
    }
    ctx->launches += 3;
    return B200_OK;
}

}  // namespace scan
}  // namespace b200
