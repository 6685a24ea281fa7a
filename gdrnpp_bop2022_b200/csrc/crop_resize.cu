// ROI crop + resize on the GPU: the reference crops every ROI with
// crop_resize_by_warp_affine = cv2.warpAffine (core/utils/data_utils.py:115-133), one ROI at a time on the host
// (demo/predictor_gdrn.py:417-438; datasets/data_loader.py:758-797), for the image (uint8, bilinear, then
// normalize_image), the 2-D coordinate grid (float, bilinear, 64x64) and the depth crop (float, nearest).
// These kernels reproduce OpenCV's warpAffine arithmetic exactly (the CPU restatement used by the tests is
// pinned against cv2 itself): double-precision inverse of the forward matrix, 10-bit fixed-point source coordinates
// built from cvRound()ed per-column / per-row terms, 5 fractional bits, 15-bit integer bilinear weights with the
// (32767, 0, 0, 1) quirk at integer positions, (sum + 2^14) >> 15, zero border.  One thread per output pixel; the
// uint8 path writes the normalised NCHW fp32 crop the stem kernel consumes.  Memory-bound: 12 B written per pixel.
#include "common.cuh"

namespace {

struct SrcCoord { int X, Y; };

// every double operation with explicit rounding: nvcc must not contract a*b+c into an FMA (the host code does not)
__device__ __forceinline__ void invert_affine(const double* __restrict__ M, double (&iM)[6]) {
  double D = __dsub_rn(__dmul_rn(M[0], M[4]), __dmul_rn(M[1], M[3]));
  D = D != 0 ? __ddiv_rn(1.0, D) : 0.0;
  const double A11 = __dmul_rn(M[4], D), A22 = __dmul_rn(M[0], D);
  iM[0] = A11; iM[1] = __dmul_rn(M[1], -D); iM[3] = __dmul_rn(M[3], -D); iM[4] = A22;
  iM[2] = __dsub_rn(__dmul_rn(-iM[0], M[2]), __dmul_rn(iM[1], M[5]));
  iM[5] = __dsub_rn(__dmul_rn(-iM[3], M[2]), __dmul_rn(iM[4], M[5]));
}

__device__ __forceinline__ SrcCoord src_coord(const double (&iM)[6], int x, int y, bool nearest) {
  constexpr int AB_BITS = 10, AB_SCALE = 1 << 10, INTER_BITS = 5;
  const int round_delta = nearest ? AB_SCALE / 2 : AB_SCALE / 32 / 2;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(iM[0], (double)x), (double)AB_SCALE));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(iM[3], (double)x), (double)AB_SCALE));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[1], (double)y), iM[2]), (double)AB_SCALE)) + round_delta;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[4], (double)y), iM[5]), (double)AB_SCALE)) + round_delta;
  SrcCoord c;
  const int sh = nearest ? AB_BITS : AB_BITS - INTER_BITS;
  c.X = (X0 + adelta) >> sh;
  c.Y = (Y0 + bdelta) >> sh;
  return c;
}

__device__ __forceinline__ int sat_short(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

__global__ void __launch_bounds__(256)
crop_resize_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int C, const double* __restrict__ Ms, int n, int oh,
                      int ow, double m0, double m1, double m2, double s0, double s1, double s2, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)oh * ow;
  if (idx >= per * n) return;
  const int b = (int)(idx / per);
  const int y = (int)((idx % per) / ow), x = (int)(idx % ow);
  double iM[6];
  invert_affine(Ms + 6 * b, iM);
  const SrcCoord sc = src_coord(iM, x, y, false);
  const int sx = sat_short(sc.X >> 5), sy = sat_short(sc.Y >> 5);
  const int fx = sc.X & 31, fy = sc.Y & 31;
  // 15-bit weights: exact products of the 1/32 fractions; saturate_cast<short>(32768) = 32767, deficit to weight 3
  int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
  if (w[0] > 32767) { w[0] = 32767; w[3] += 1; }
  const bool in00 = sx >= 0 && sx < W && sy >= 0 && sy < H, in01 = sx + 1 >= 0 && sx + 1 < W && sy >= 0 && sy < H;
  const bool in10 = sx >= 0 && sx < W && sy + 1 >= 0 && sy + 1 < H, in11 = sx + 1 >= 0 && sx + 1 < W && sy + 1 >= 0 && sy + 1 < H;
  const uint8_t* p00 = img + ((long long)sy * W + sx) * C;
  const double mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (int c = 0; c < C; ++c) {
    int v = 0;
    if (in00) v += p00[c] * w[0];
    if (in01) v += p00[C + c] * w[1];
    if (in10) v += p00[(long long)W * C + c] * w[2];
    if (in11) v += p00[(long long)W * C + C + c] * w[3];
    v = (v + (1 << 14)) >> 15;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    // normalize_image: (uint8 - mean) / std in float64, then astype(float32) (predictor_gdrn.py normalize_image)
    const int cc = c < 3 ? c : 2;
    out[((long long)b * C + c) * per + (long long)y * ow + x] = __double2float_rn(__ddiv_rn(__dsub_rn((double)v, mean[cc]), stdv[cc]));
  }
}

__global__ void __launch_bounds__(256)
crop_resize_f32_kernel(const float* __restrict__ src, int H, int W, int C, const double* __restrict__ Ms, int n, int oh,
                       int ow, int nearest, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)oh * ow;
  if (idx >= per * n) return;
  const int b = (int)(idx / per);
  const int y = (int)((idx % per) / ow), x = (int)(idx % ow);
  double iM[6];
  invert_affine(Ms + 6 * b, iM);
  const SrcCoord sc = src_coord(iM, x, y, nearest != 0);
  float* o = out + (long long)b * C * per + (long long)y * ow + x;
  if (nearest) {
    const bool in = sc.X >= 0 && sc.X < W && sc.Y >= 0 && sc.Y < H;
    for (int c = 0; c < C; ++c) o[c * per] = in ? src[((long long)sc.Y * W + sc.X) * C + c] : 0.f;
    return;
  }
  const int sx = sc.X >> 5, sy = sc.Y >> 5;
  const float ax = (float)(sc.X & 31) * (1.f / 32), ay = (float)(sc.Y & 31) * (1.f / 32);
  const float w0 = __fmul_rn(1.f - ay, 1.f - ax), w1 = __fmul_rn(1.f - ay, ax), w2 = __fmul_rn(ay, 1.f - ax), w3 = __fmul_rn(ay, ax);
  const bool in00 = sx >= 0 && sx < W && sy >= 0 && sy < H, in01 = sx + 1 >= 0 && sx + 1 < W && sy >= 0 && sy < H;
  const bool in10 = sx >= 0 && sx < W && sy + 1 >= 0 && sy + 1 < H, in11 = sx + 1 >= 0 && sx + 1 < W && sy + 1 >= 0 && sy + 1 < H;
  const float* p00 = src + ((long long)sy * W + sx) * C;
  for (int c = 0; c < C; ++c) {
    const float a = in00 ? p00[c] : 0.f, bq = in01 ? p00[C + c] : 0.f;
    const float cq = in10 ? p00[(long long)W * C + c] : 0.f, d = in11 ? p00[(long long)W * C + C + c] : 0.f;
    // same association as OpenCV's scalar remapBilinear: ((a*w0 + b*w1) + c*w2) + d*w3, no FMA
    o[c * per] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, w0), __fmul_rn(bq, w1)), __fmul_rn(cq, w2)), __fmul_rn(d, w3));
  }
}

}  // namespace

extern "C" int gdrn_crop_resize_u8(const uint8_t* image, int H, int W, int C, const double* M, int n, int out_h, int out_w,
                                   const double* pixel_mean, const double* pixel_std, float* out, void* stream) {
  GDRN_REQUIRE(H > 0 && W > 0 && C >= 1 && C <= 4 && n >= 0 && out_h > 0 && out_w > 0, "crop_resize_u8: bad shape");
  if (n == 0) return GDRN_OK;   // empty detection list: nothing to do (the reference loop simply does not iterate)
  GDRN_REQUIRE(image && M && out && pixel_mean && pixel_std, "crop_resize_u8: null argument");
  const long long total = (long long)n * out_h * out_w;
  const double* mu = pixel_mean;
  const double* sd = pixel_std;
  crop_resize_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      image, H, W, C, M, n, out_h, out_w, mu[0], mu[C > 1 ? 1 : 0], mu[C > 2 ? 2 : 0], sd[0], sd[C > 1 ? 1 : 0],
      sd[C > 2 ? 2 : 0], out);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

extern "C" int gdrn_crop_resize_f32(const float* src, int H, int W, int C, const double* M, int n, int out_h, int out_w,
                                    int nearest, float* out, void* stream) {
  GDRN_REQUIRE(H > 0 && W > 0 && C >= 1 && n >= 0 && out_h > 0 && out_w > 0, "crop_resize_f32: bad shape");
  if (n == 0) return GDRN_OK;
  GDRN_REQUIRE(src && M && out, "crop_resize_f32: null argument");
  const long long total = (long long)n * out_h * out_w;
  crop_resize_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, H, W, C, M, n, out_h, out_w,
                                                                                       nearest, out);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}
