// Online training-target path: rendered depth -> object-space xyz -> mask / nearest-FPS-region
// labels / normalised xyz, one fused kernel after the CUDA rasteriser.
//
// Replaces, in core/gdrn_modeling/engine/engine_utils.py:131-187 (batch_data, XYZ_BP branch), the per-ROI EGL render
// loop (rast_render_meshes does all ROIs in one launch) and the torch ops that follow it:
//   misc.calc_xyz_bp_batch(depth, R, T, K, fmt="BHWC")      lib/pysixd/misc.py:412-457
//       xyz_cam = ((x - cx) * d / fx, (y - cy) * d / fy, d)  with x, y the INTEGER pixel indices (the helper's convention),
//       xyz = R^T (xyz_cam - T) * (d != 0)
//   roi_mask_obj = (xyz_x != 0) & (xyz_y != 0) & (xyz_z != 0)                   engine_utils.py:171-173
//   xyz_to_region_batch(xyz, fps_points, mask)               core/utils/data_utils.py:283-301
//       region = argmin_f |xyz - fps_f| + 1, times the mask (0 = background), int64
//   roi_xyz = xyz / extent + 0.5  (b c h w)                  engine_utils.py:183
// Arithmetic in fp32 in the order torch evaluates it (products before the division, three-term dot products left to
// right); argmin takes the lowest index among equal distances like torch.
#include "common.cuh"

namespace {

constexpr int TG_MAX_FPS = 256;

__global__ void __launch_bounds__(256)
xyz_region_kernel(const float* __restrict__ depth, const float* __restrict__ R, const float* __restrict__ T,
                  const float* __restrict__ K, const float* __restrict__ fps, const float* __restrict__ extents, int H,
                  int W, int F, float* __restrict__ roi_xyz, float* __restrict__ xyz_raw, float* __restrict__ mask_obj,
                  long long* __restrict__ region) {
  __shared__ float s_fps[TG_MAX_FPS * 3];
  const int r = blockIdx.y;
  const int npix = H * W;
  if (fps)
    for (int i = threadIdx.x; i < F * 3; i += blockDim.x) s_fps[i] = fps[(size_t)r * F * 3 + i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float d = depth[(size_t)r * npix + i];
  const float* Rr = R + r * 9;
  const float* Kr = K + r * 9;
  const float X = __fsub_rn((float)(i % W), Kr[2]), Y = __fsub_rn((float)(i / W), Kr[5]);
  const float vx = __fsub_rn(__fdiv_rn(__fmul_rn(X, d), Kr[0]), T[r * 3]);
  const float vy = __fsub_rn(__fdiv_rn(__fmul_rn(Y, d), Kr[4]), T[r * 3 + 1]);
  const float vz = __fsub_rn(d, T[r * 3 + 2]);
  const float m = d != 0.f ? 1.f : 0.f;
  // R^T v: row i of R^T = column i of R
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    o[c] = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rr[c], vx), __fmul_rn(Rr[3 + c], vy)), __fmul_rn(Rr[6 + c], vz)), m);
  const float mo = (o[0] != 0.f && o[1] != 0.f && o[2] != 0.f) ? 1.f : 0.f;
  if (mask_obj) mask_obj[(size_t)r * npix + i] = mo;
  if (xyz_raw) {
#pragma unroll
    for (int c = 0; c < 3; ++c) xyz_raw[((size_t)r * npix + i) * 3 + c] = o[c];
  }
  if (roi_xyz) {
#pragma unroll
    for (int c = 0; c < 3; ++c) roi_xyz[((size_t)r * 3 + c) * npix + i] = __fadd_rn(__fdiv_rn(o[c], extents[r * 3 + c]), 0.5f);
  }
  if (region) {
    float best = INFINITY;
    int bi = 0;
    for (int f = 0; f < F; ++f) {
      const float dx = __fsub_rn(o[0], s_fps[f * 3]), dy = __fsub_rn(o[1], s_fps[f * 3 + 1]), dz = __fsub_rn(o[2], s_fps[f * 3 + 2]);
      const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (dd < best) { best = dd; bi = f; }
    }
    region[(size_t)r * npix + i] = mo != 0.f ? (long long)(bi + 1) : 0;
  }
}

}  // namespace

extern "C" int gdrn_xyz_region_targets(const float* depth, const float* R, const float* T, const float* K,
                                       const float* fps_points, const float* extents, int n, int H, int W, int F,
                                       float* roi_xyz, float* xyz_raw, float* mask_obj, long long* region, void* stream) {
  GDRN_REQUIRE(depth && R && T && K, "xyz_region_targets: null argument");
  GDRN_REQUIRE(n > 0 && H > 0 && W > 0, "xyz_region_targets: empty input");
  GDRN_REQUIRE(!region || (fps_points && F >= 1 && F <= TG_MAX_FPS), "xyz_region_targets: region labels need 1..256 fps points");
  GDRN_REQUIRE(!roi_xyz || extents, "xyz_region_targets: normalised xyz needs the extents");
  xyz_region_kernel<<<dim3((H * W + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(depth, R, T, K, region ? fps_points : nullptr, extents, H, W,
                                                                                  F, roi_xyz, xyz_raw, mask_obj, region);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}
