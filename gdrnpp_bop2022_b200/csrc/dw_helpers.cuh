// Device helpers shared by the depthwise-conv + LayerNorm kernels (dense_ops.cu, dwconv_pp.cu): packed-FP32 FMA
// (fma.rn.f32x2, SASS FFMA2) and the warp-transposing reduction used for per-pixel channel sums.
#pragma once
#include "common.cuh"

namespace {

typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(f32x2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

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
