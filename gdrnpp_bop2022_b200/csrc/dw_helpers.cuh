// Device helper shared by the depthwise-conv + LayerNorm kernels (dense_ops.cu, dwconv_pp.cu): the warp-transposing
// reduction used for per-pixel channel sums (the packed-FP32 helpers live in common.cuh).
#pragma once
#include "common.cuh"

namespace {

// sum over the 32 lanes of N per-lane values, N in {32,16,8}: lane L ends up with element (L * N) >> 5
template <int N>
__device__ __forceinline__ float lane_transpose_reduce(float (&a)[N], int lane) {
  static_assert(N == 32 || N == 16 || N == 8, "N");
  constexpr int STEPS = (N == 32) ? 5 : (N == 16 ? 4 : 3);
#pragma unroll
  for (int st = 0; st < STEPS; ++st) {
    const int o = 16 >> st;        // lane bit
    const int n = (N / 2) >> st;   // values kept after this step
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = up ? a[j] : a[j + n];
      const float keep = up ? a[j + n] : a[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (int o = (16 >> STEPS); o > 0; o >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
  return a[0];
}

}  // namespace
