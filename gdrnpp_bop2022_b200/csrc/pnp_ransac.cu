// Batched RANSAC-PnP from dense 2D-3D correspondences: one CTA per ROI.
//
// Replaces the per-ROI host loop of the reference's TEST.USE_PNP path: get_pnp_ransac_pose
// (core/gdrn_modeling/engine/gdrn_evaluator.py:1122-1180: get_out_coor / get_out_mask, correspondence selection
// get_img_model_points_with_coords2d :1183-1221) followed by misc.pnp_v2 (lib/pysixd/misc.py:153-208) =
// cv2.solvePnPRansac(flags=SOLVEPNP_EPNP, reprojectionError=3, iterationsCount=100) -- which GdrnPredictor runs by
// default (core/gdrn_modeling/demo/predictor_gdrn.py:58,169).
//
// Per ROI (all on chip, fp64 geometry):
//   1. L1 mask -> per-ROI min-max normalisation (engine_utils.py:313-333), xyz = (coor - 0.5) * extent, image points =
//      roi_coord_2d * (im_W, im_H); selection mask > thr & |xyz_k| > 1e-4 * extent_k, compacted in row-major order
//      (the order numpy boolean indexing gives the reference).
//   2. `iters` hypotheses: 4 sampled correspondences each (counter-based hash RNG of (seed, roi, hypothesis), or caller
//      supplied indices), Grunert P3P on the first three (quartic by Ferrari + Newton polish), the fourth picks among
//      the <= 4 solutions; one thread per hypothesis.
//   3. inlier count of every hypothesis over all correspondences (reprojection error < thr px; a warp per hypothesis,
//      lanes over points), best = max count (lowest index on ties).
//   4. refit on the inliers of the best hypothesis: Levenberg-Marquardt on the reprojection error (6x6 normal equations
//      reduced over the CTA), then the inlier set is recomputed with the refined pose and refitted once more.
// OpenCV's final step is EPnP on the inliers (an algebraic solution + Gauss-Newton); both minimise the same residual, so
// poses agree to the tolerance tests/test_gpu_parity.py states, not bit-exactly.  < 4 correspondences -> the
// reference's sentinel pose (-100 everywhere, gdrn_evaluator.py:1178-1179).
#include "common.cuh"

namespace {

constexpr int PR_THREADS = 512;
constexpr int PR_MAX_PTS = 4096;   // 64 x 64 maps
constexpr int PR_MAX_HYP = 256;

struct PnpParams {
  const float* coor_x; const float* coor_y; const float* coor_z;   // [n, hw*hw] network outputs in [0,1]
  const float* mask;        // [n, hw*hw] raw visible-mask output, or null (then `sel` picks)
  const float* coord2d;     // [n, 2, hw*hw] normalised image coordinates of the ROI grid
  const float* im_hw;       // [n, 2] (im_H, im_W)
  const float* extents;     // [n, 3]
  const float* Ks;          // [n, 9]
  const int* idxs;          // [n, iters, 4] indices into the compacted correspondence list (taken modulo count), or null
  // generic-correspondence entry (pts given directly): pts3d [n, npts, 3], pts2d [n, npts, 2], npts <= PR_MAX_PTS
  const float* pts3d; const float* pts2d;
  int npts;                 // pixels per ROI (hw*hw) or points per problem
  int iters;
  float mask_thr, reproj_thr;
  unsigned seed;
  float* poses;             // [n, 12]
  int* n_inliers;           // [n] (or null)
  unsigned char* inlier_mask;   // [n, npts] over the INPUT pixels/points (or null)
};

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// real roots of a4 x^4 + a3 x^3 + a2 x^2 + a1 x + a0 (Ferrari via one positive root of the resolvent cubic), polished
__device__ int solve_quartic(double a4, double a3, double a2, double a1, double a0, double* out) {
  if (fabs(a4) < 1e-300) return 0;
  const double b = a3 / a4, c = a2 / a4, d = a1 / a4, e = a0 / a4;
  const double b2 = b * b;
  const double p = c - 3.0 * b2 / 8.0;
  const double q = d - b * c / 2.0 + b2 * b / 8.0;
  const double r = e - b * d / 4.0 + b2 * c / 16.0 - 3.0 * b2 * b2 / 256.0;
  double y[4];
  int n = 0;
  if (fabs(q) < 1e-14 * fmax(1.0, pow(fabs(p), 1.5))) {
    const double disc = p * p - 4.0 * r;
    if (disc >= 0.0) {
      const double s = sqrt(disc);
      const double z0 = (-p + s) / 2.0, z1 = (-p - s) / 2.0;
      if (z0 >= 0.0) { y[n++] = sqrt(z0); y[n++] = -sqrt(z0); }
      if (z1 >= 0.0) { y[n++] = sqrt(z1); y[n++] = -sqrt(z1); }
    }
  } else {
    const double A = p, B = p * p / 4.0 - r, C = -q * q / 8.0;
    const double P = B - A * A / 3.0, Q = 2.0 * A * A * A / 27.0 - A * B / 3.0 + C;
    const double D = Q * Q / 4.0 + P * P * P / 27.0;
    double m;
    if (D >= 0.0) {
      const double sD = sqrt(D);
      m = cbrt(-Q / 2.0 + sD) + cbrt(-Q / 2.0 - sD) - A / 3.0;
    } else {
      const double rho = sqrt(-P * P * P / 27.0);
      const double th = acos(fmax(-1.0, fmin(1.0, -Q / (2.0 * rho))));
      const double mm = 2.0 * sqrt(-P / 3.0);
      m = -1e300;
      for (int k = 0; k < 3; ++k) m = fmax(m, mm * cos((th + 6.283185307179586 * k) / 3.0) - A / 3.0);
    }
    for (int it = 0; it < 2; ++it) {
      const double fm = ((m + A) * m + B) * m + C, dfm = (3.0 * m + 2.0 * A) * m + B;
      if (dfm != 0.0) m -= fm / dfm;
    }
    if (m <= 0.0) return 0;
    const double s = sqrt(2.0 * m);
    for (int sg = 1; sg >= -1; sg -= 2) {
      const double cc = p / 2.0 + m - sg * q / (2.0 * s);
      const double disc = s * s - 4.0 * cc;
      if (disc >= 0.0) {
        const double sd = sqrt(disc);
        y[n++] = (-sg * s + sd) / 2.0;
        y[n++] = (-sg * s - sd) / 2.0;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - b / 4.0;
    for (int it = 0; it < 2; ++it) {
      const double fx = (((a4 * x + a3) * x + a2) * x + a1) * x + a0;
      const double dfx = ((4.0 * a4 * x + 3.0 * a3) * x + 2.0 * a2) * x + a1;
      if (dfx != 0.0) x -= fx / dfx;
    }
    out[i] = x;
  }
  return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// orthonormal frame (columns e1, e2, e3) of the triangle X0 X1 X2; false if degenerate
__device__ bool tri_frame(const double* X0, const double* X1, const double* X2, double* F /*[9] row-major, columns = e*/) {
  double e1[3] = {X1[0] - X0[0], X1[1] - X0[1], X1[2] - X0[2]};
  double n1 = norm3(e1);
  if (n1 < 1e-12) return false;
  e1[0] /= n1; e1[1] /= n1; e1[2] /= n1;
  double w[3] = {X2[0] - X0[0], X2[1] - X0[1], X2[2] - X0[2]};
  double e3[3];
  cross3(e1, w, e3);
  double n3 = norm3(e3);
  if (n3 < 1e-12) return false;
  e3[0] /= n3; e3[1] /= n3; e3[2] /= n3;
  double e2[3];
  cross3(e3, e1, e2);
  for (int i = 0; i < 3; ++i) { F[i * 3] = e1[i]; F[i * 3 + 1] = e2[i]; F[i * 3 + 2] = e3[i]; }
  return true;
}

// Grunert P3P: world points P[3][3], unit bearings f[3][3] -> up to 4 (R, t) with x_cam = R X + t; returns count
__device__ int p3p_grunert(const double (*P)[3], const double (*f)[3], double (*Rt)[12]) {
  double d12[3] = {P[1][0] - P[2][0], P[1][1] - P[2][1], P[1][2] - P[2][2]};
  double d02[3] = {P[0][0] - P[2][0], P[0][1] - P[2][1], P[0][2] - P[2][2]};
  double d01[3] = {P[0][0] - P[1][0], P[0][1] - P[1][1], P[0][2] - P[1][2]};
  const double a2 = d12[0] * d12[0] + d12[1] * d12[1] + d12[2] * d12[2];
  const double b2 = d02[0] * d02[0] + d02[1] * d02[1] + d02[2] * d02[2];
  const double c2 = d01[0] * d01[0] + d01[1] * d01[1] + d01[2] * d01[2];
  if (a2 < 1e-18 || b2 < 1e-18 || c2 < 1e-18) return 0;
  const double ca = f[1][0] * f[2][0] + f[1][1] * f[2][1] + f[1][2] * f[2][2];
  const double cb = f[0][0] * f[2][0] + f[0][1] * f[2][1] + f[0][2] * f[2][2];
  const double cg = f[0][0] * f[1][0] + f[0][1] * f[1][1] + f[0][2] * f[1][2];
  const double q = (a2 - c2) / b2, r = (a2 + c2) / b2;
  const double A4 = (q - 1.0) * (q - 1.0) - 4.0 * c2 / b2 * ca * ca;
  const double A3 = 4.0 * (q * (1.0 - q) * cb - (1.0 - r) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  const double A2 = 2.0 * (q * q - 1.0 + 2.0 * q * q * cb * cb + 2.0 * (b2 - c2) / b2 * ca * ca - 4.0 * r * ca * cb * cg +
                           2.0 * (b2 - a2) / b2 * cg * cg);
  const double A1 = 4.0 * (-q * (1.0 + q) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - r) * ca * cg);
  const double A0 = (1.0 + q) * (1.0 + q) - 4.0 * a2 / b2 * cg * cg;
  double roots[4];
  const int nr = solve_quartic(A4, A3, A2, A1, A0, roots);
  double Fw[9];
  if (!tri_frame(P[0], P[1], P[2], Fw)) return 0;
  int ns = 0;
  for (int i = 0; i < nr; ++i) {
    const double v = roots[i];
    if (!(v > 0.0)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (fabs(den) < 1e-12) continue;
    const double u = ((-1.0 + q) * v * v - 2.0 * q * cb * v + 1.0 + q) / den;
    if (!(u > 0.0)) continue;
    const double s1sq = c2 / (1.0 + u * u - 2.0 * u * cg);
    if (!(s1sq > 0.0)) continue;
    const double s1 = sqrt(s1sq), s2 = u * s1, s3 = v * s1;
    double Q0[3] = {s1 * f[0][0], s1 * f[0][1], s1 * f[0][2]};
    double Q1[3] = {s2 * f[1][0], s2 * f[1][1], s2 * f[1][2]};
    double Q2[3] = {s3 * f[2][0], s3 * f[2][1], s3 * f[2][2]};
    double Fc[9];
    if (!tri_frame(Q0, Q1, Q2, Fc)) continue;
    double* o = Rt[ns];
    for (int rr = 0; rr < 3; ++rr)
      for (int cc = 0; cc < 3; ++cc)
        o[rr * 4 + cc] = Fc[rr * 3] * Fw[cc * 3] + Fc[rr * 3 + 1] * Fw[cc * 3 + 1] + Fc[rr * 3 + 2] * Fw[cc * 3 + 2];
    for (int rr = 0; rr < 3; ++rr)
      o[rr * 4 + 3] = (rr == 0 ? Q0[0] : rr == 1 ? Q0[1] : Q0[2]) - (o[rr * 4] * P[0][0] + o[rr * 4 + 1] * P[0][1] + o[rr * 4 + 2] * P[0][2]);
    ++ns;
  }
  return ns;
}

__device__ __forceinline__ bool project(const double* Rt, const double* K, float X, float Y, float Z, double* u, double* v) {
  const double xc = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[3];
  const double yc = Rt[4] * X + Rt[5] * Y + Rt[6] * Z + Rt[7];
  const double zc = Rt[8] * X + Rt[9] * Y + Rt[10] * Z + Rt[11];
  if (!(zc > 1e-9)) return false;
  *u = K[0] * xc / zc + K[1] * yc / zc + K[2];
  *v = K[4] * yc / zc + K[5];
  return true;
}

// symmetric 6x6 solve (A + lambda diag(A)) x = g by Cholesky; false if not positive definite
__device__ bool solve6(const double* A, const double* g, double lambda, double* x) {
  double L[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * 6 + j];
      if (i == j) s += lambda * fmax(A[i * 6 + i], 1e-12);
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = sqrt(s);
      } else {
        L[i * 6 + j] = s / L[j * 6 + j];
      }
    }
  double yv[6];
  for (int i = 0; i < 6; ++i) {
    double s = g[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * yv[k];
    yv[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = yv[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
  return true;
}

// R <- exp([w]x) R  (Rodrigues), t <- t + dt
__device__ void apply_update(double* Rt, const double* d) {
  const double th = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (th > 1e-14) {
    const double kx = d[0] / th, ky = d[1] / th, kz = d[2] / th, c = cos(th), s = sin(th), C = 1.0 - c;
    E[0] = c + kx * kx * C; E[1] = kx * ky * C - kz * s; E[2] = kx * kz * C + ky * s;
    E[3] = ky * kx * C + kz * s; E[4] = c + ky * ky * C; E[5] = ky * kz * C - kx * s;
    E[6] = kz * kx * C - ky * s; E[7] = kz * ky * C + kx * s; E[8] = c + kz * kz * C;
  }
  double Rn[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = E[r * 3] * Rt[c] + E[r * 3 + 1] * Rt[4 + c] + E[r * 3 + 2] * Rt[8 + c];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rt[r * 4 + c] = Rn[r * 3 + c];
    Rt[r * 4 + 3] += d[3 + r];
  }
}

__global__ void __launch_bounds__(PR_THREADS)
pnp_ransac_kernel(const PnpParams p) {
  extern __shared__ float sm_f[];
  float* s3 = sm_f;                         // [PR_MAX_PTS][3] model points
  float* s2 = s3 + 3 * PR_MAX_PTS;          // [PR_MAX_PTS][2] image points
  unsigned short* s_src = reinterpret_cast<unsigned short*>(s2 + 2 * PR_MAX_PTS);   // [PR_MAX_PTS] source pixel of entry i
  unsigned char* s_inl = reinterpret_cast<unsigned char*>(s_src + PR_MAX_PTS);      // [PR_MAX_PTS] inlier flags
  __shared__ double s_hyp[PR_MAX_HYP][12];
  __shared__ int s_cnt[PR_MAX_HYP];
  __shared__ double s_red[PR_THREADS / 32][28];
  __shared__ double s_pose[12], s_K[9], s_step[8];
  __shared__ float s_mm[2];
  __shared__ int s_count, s_scan[PR_THREADS / 32 + 1], s_flag;

  const int roi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npts = p.npts;
  if (tid < 9) s_K[tid] = (double)p.Ks[roi * 9 + tid];

  // ---------------- 1. correspondences ----------------
  if (p.pts3d == nullptr) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = tid; i < npts; i += PR_THREADS) {
      const float m = p.mask[(long long)roi * npts + i];
      mn = fminf(mn, m); mx = fmaxf(mx, m);
    }
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { reinterpret_cast<float*>(s_red)[warp * 2] = mn; reinterpret_cast<float*>(s_red)[warp * 2 + 1] = mx; }
    __syncthreads();
    if (tid == 0) {
      float a = INFINITY, b = -INFINITY;
      for (int w = 0; w < PR_THREADS / 32; ++w) { a = fminf(a, reinterpret_cast<float*>(s_red)[w * 2]); b = fmaxf(b, reinterpret_cast<float*>(s_red)[w * 2 + 1]); }
      s_mm[0] = a; s_mm[1] = b;
    }
    __syncthreads();
  }
  {
    // ordered compaction: chunks of PR_THREADS pixels, block-wide exclusive scan of the selection flags
    const float ex = p.extents ? p.extents[roi * 3] : 0.f, ey = p.extents ? p.extents[roi * 3 + 1] : 0.f, ez = p.extents ? p.extents[roi * 3 + 2] : 0.f;
    const float imH = p.im_hw ? p.im_hw[roi * 2] : 1.f, imW = p.im_hw ? p.im_hw[roi * 2 + 1] : 1.f;
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int base = 0; base < npts; base += PR_THREADS) {
      const int i = base + tid;
      bool sel = false;
      float X = 0.f, Y = 0.f, Z = 0.f, U = 0.f, V = 0.f;
      if (i < npts) {
        if (p.pts3d) {
          X = p.pts3d[((long long)roi * npts + i) * 3]; Y = p.pts3d[((long long)roi * npts + i) * 3 + 1]; Z = p.pts3d[((long long)roi * npts + i) * 3 + 2];
          U = p.pts2d[((long long)roi * npts + i) * 2]; V = p.pts2d[((long long)roi * npts + i) * 2 + 1];
          sel = true;
        } else {
          // get_out_mask (L1): (m - min) / (max - min); xyz denormalisation and image points: gdrn_evaluator.py:1196-1212
          const float m = (p.mask[(long long)roi * npts + i] - s_mm[0]) / (s_mm[1] - s_mm[0]);
          X = (p.coor_x[(long long)roi * npts + i] - 0.5f) * ex;
          Y = (p.coor_y[(long long)roi * npts + i] - 0.5f) * ey;
          Z = (p.coor_z[(long long)roi * npts + i] - 0.5f) * ez;
          U = p.coord2d[((long long)roi * 2) * npts + i] * imW;
          V = p.coord2d[((long long)roi * 2 + 1) * npts + i] * imH;
          sel = (m > p.mask_thr) && (fabsf(X) > 0.0001f * ex) && (fabsf(Y) > 0.0001f * ey) && (fabsf(Z) > 0.0001f * ez);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, sel);
      if (lane == 0) s_scan[warp] = __popc(bal);
      __syncthreads();
      if (tid == 0) {
        int acc = s_count;
        for (int w = 0; w < PR_THREADS / 32; ++w) { const int c = s_scan[w]; s_scan[w] = acc; acc += c; }
        s_scan[PR_THREADS / 32] = acc;
      }
      __syncthreads();
      if (sel) {
        const int slot = s_scan[warp] + __popc(bal & ((1u << lane) - 1u));
        s3[slot * 3] = X; s3[slot * 3 + 1] = Y; s3[slot * 3 + 2] = Z;
        s2[slot * 2] = U; s2[slot * 2 + 1] = V;
        s_src[slot] = (unsigned short)i;
      }
      __syncthreads();
      if (tid == 0) s_count = s_scan[PR_THREADS / 32];
      __syncthreads();
    }
  }
  const int N = s_count;
  float* out = p.poses + (long long)roi * 12;
  if (p.inlier_mask) for (int i = tid; i < npts; i += PR_THREADS) p.inlier_mask[(long long)roi * npts + i] = 0;
  if (N < 4) {   // gdrn_evaluator.py:1178-1179
    if (tid < 12) out[tid] = -100.f;
    if (tid == 0 && p.n_inliers) p.n_inliers[roi] = 0;
    return;
  }

  // ---------------- 2. hypotheses: one thread each ----------------
  const int H = p.iters;
  const double ifx = 1.0 / s_K[0], ify = 1.0 / s_K[4];
  for (int h = tid; h < H; h += PR_THREADS) {
    int id[4];
    if (p.idxs) {
      for (int k = 0; k < 4; ++k) id[k] = (int)((unsigned)p.idxs[((long long)roi * H + h) * 4 + k] % (unsigned)N);
    } else {
      unsigned st = hash32(p.seed ^ hash32((unsigned)roi * 0x9E3779B9u + (unsigned)h));
      for (int k = 0; k < 4; ++k) {
        bool dup = true;
        for (int tries = 0; tries < 8 && dup; ++tries) {
          st = hash32(st + 0x6D2B79F5u);
          id[k] = (int)(st % (unsigned)N);
          dup = false;
          for (int j = 0; j < k; ++j) dup = dup || (id[j] == id[k]);
        }
      }
    }
    double P[3][3], f[3][3];
    for (int k = 0; k < 3; ++k) {
      P[k][0] = s3[id[k] * 3]; P[k][1] = s3[id[k] * 3 + 1]; P[k][2] = s3[id[k] * 3 + 2];
      const double yn = ((double)s2[id[k] * 2 + 1] - s_K[5]) * ify;
      const double xn = ((double)s2[id[k] * 2] - s_K[2] - s_K[1] * yn) * ifx;
      const double nn = sqrt(xn * xn + yn * yn + 1.0);
      f[k][0] = xn / nn; f[k][1] = yn / nn; f[k][2] = 1.0 / nn;
    }
    double sols[4][12];
    const int ns = p3p_grunert(P, f, sols);
    double best = 1e300;
    int bi = -1;
    for (int s = 0; s < ns; ++s) {
      double u, v;
      if (!project(sols[s], s_K, s3[id[3] * 3], s3[id[3] * 3 + 1], s3[id[3] * 3 + 2], &u, &v)) continue;
      const double du = u - s2[id[3] * 2], dv = v - s2[id[3] * 2 + 1];
      const double e = du * du + dv * dv;
      if (e < best) { best = e; bi = s; }
    }
    if (bi >= 0) {
      for (int k = 0; k < 12; ++k) s_hyp[h][k] = sols[bi][k];
      s_cnt[h] = 0;
    } else {
      s_cnt[h] = -1;   // no valid model from this sample
    }
  }
  __syncthreads();

  // ---------------- 3. inlier counts: a warp per hypothesis ----------------
  const double thr2 = (double)p.reproj_thr * (double)p.reproj_thr;
  for (int h = warp; h < H; h += PR_THREADS / 32) {
    if (s_cnt[h] < 0) continue;
    int c = 0;
    for (int i = lane; i < N; i += 32) {
      double u, v;
      if (project(s_hyp[h], s_K, s3[i * 3], s3[i * 3 + 1], s3[i * 3 + 2], &u, &v)) {
        const double du = u - s2[i * 2], dv = v - s2[i * 2 + 1];
        c += (du * du + dv * dv < thr2) ? 1 : 0;
      }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_cnt[h] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int bh = -1, bc = -1;
    for (int h = 0; h < H; ++h)
      if (s_cnt[h] > bc) { bc = s_cnt[h]; bh = h; }
    s_flag = bh;
    if (bh >= 0) for (int k = 0; k < 12; ++k) s_pose[k] = s_hyp[bh][k];
  }
  __syncthreads();
  if (s_flag < 0 || s_cnt[s_flag] < 4) {   // RANSAC found no model: the reference's solvePnPRansac returns garbage/false here
    if (tid < 12) out[tid] = -100.f;
    if (tid == 0 && p.n_inliers) p.n_inliers[roi] = 0;
    return;
  }

  // ---------------- 4. refit on the inliers (LM), twice ----------------
  for (int round = 0; round < 2; ++round) {
    for (int i = tid; i < N; i += PR_THREADS) {
      double u, v;
      bool in = false;
      if (project(s_pose, s_K, s3[i * 3], s3[i * 3 + 1], s3[i * 3 + 2], &u, &v)) {
        const double du = u - s2[i * 2], dv = v - s2[i * 2 + 1];
        in = du * du + dv * dv < thr2;
      }
      s_inl[i] = in ? 1 : 0;
    }
    __syncthreads();
    double lambda = 1e-4, prev_cost = -1.0;
    for (int it = 0; it < 12; ++it) {
      // normal equations over the inliers: 21 (upper JtJ) + 6 (Jt r) + 1 (cost) sums
      double acc[28];
      for (int k = 0; k < 28; ++k) acc[k] = 0.0;
      for (int i = tid; i < N; i += PR_THREADS) {
        if (!s_inl[i]) continue;
        const double X = s3[i * 3], Y = s3[i * 3 + 1], Z = s3[i * 3 + 2];
        const double rx = s_pose[0] * X + s_pose[1] * Y + s_pose[2] * Z;
        const double ry = s_pose[4] * X + s_pose[5] * Y + s_pose[6] * Z;
        const double rz = s_pose[8] * X + s_pose[9] * Y + s_pose[10] * Z;
        const double xc = rx + s_pose[3], yc = ry + s_pose[7], zc = rz + s_pose[11];
        if (!(zc > 1e-9)) continue;
        const double iz = 1.0 / zc;
        const double u = s_K[0] * xc * iz + s_K[1] * yc * iz + s_K[2], v = s_K[4] * yc * iz + s_K[5];
        const double eu = u - s2[i * 2], ev = v - s2[i * 2 + 1];
        // d(u,v)/d(xc,yc,zc)
        const double ux = s_K[0] * iz, uy = s_K[1] * iz, uz = -(s_K[0] * xc + s_K[1] * yc) * iz * iz;
        const double vy = s_K[4] * iz, vz = -s_K[4] * yc * iz * iz;
        // d xc / d w = -[R X]x  (left perturbation), d xc / d t = I
        double Ju[6], Jv[6];
        Ju[0] = uy * (-rz) + uz * ry;           // column 0 of -[r]x = (0, -rz... ) see below
        Ju[1] = ux * rz + uz * (-rx);
        Ju[2] = ux * (-ry) + uy * rx;
        Jv[0] = vy * (-rz) + vz * ry;
        Jv[1] = vz * (-rx);
        Jv[2] = vy * rx;
        Ju[3] = ux; Ju[4] = uy; Ju[5] = uz;
        Jv[3] = 0.0; Jv[4] = vy; Jv[5] = vz;
        int k = 0;
        for (int a = 0; a < 6; ++a)
          for (int b = a; b < 6; ++b) acc[k++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
        for (int a = 0; a < 6; ++a) acc[21 + a] += Ju[a] * eu + Jv[a] * ev;
        acc[27] += eu * eu + ev * ev;
      }
      for (int k = 0; k < 28; ++k)
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane == 0) for (int k = 0; k < 28; ++k) s_red[warp][k] = acc[k];
      __syncthreads();
      if (tid == 0) {
        double tot[28];
        for (int k = 0; k < 28; ++k) { tot[k] = 0.0; for (int w = 0; w < PR_THREADS / 32; ++w) tot[k] += s_red[w][k]; }
        double A[36], g[6], dx[6];
        int k = 0;
        for (int a = 0; a < 6; ++a)
          for (int b = a; b < 6; ++b) { A[a * 6 + b] = tot[k]; A[b * 6 + a] = tot[k]; ++k; }
        for (int a = 0; a < 6; ++a) g[a] = -tot[21 + a];
        const double cost = tot[27];
        int stop = 0;
        if (prev_cost >= 0.0 && cost > prev_cost) {
          // the previous step made it worse: undo it and raise the damping
          for (int a = 0; a < 12; ++a) s_pose[a] = s_hyp[0][a];   // s_hyp[0] doubles as the saved pose below
          lambda *= 10.0;
          s_step[6] = 0.0;   // do not update prev_cost
          stop = lambda > 1e6;
          s_step[7] = stop ? 1.0 : 2.0;   // 2 = retry with the saved pose (normal equations recomputed next iteration)
        } else {
          for (int a = 0; a < 12; ++a) s_hyp[0][a] = s_pose[a];   // save
          if (solve6(A, g, lambda, dx)) {
            apply_update(s_pose, dx);
            double nrm = 0.0;
            for (int a = 0; a < 6; ++a) nrm += dx[a] * dx[a];
            stop = nrm < 1e-20;
            lambda = fmax(lambda * 0.3, 1e-9);
          } else {
            lambda *= 10.0;
            stop = lambda > 1e6;
          }
          s_step[6] = 1.0;
          s_step[5] = cost;
          s_step[7] = stop ? 1.0 : 0.0;
        }
      }
      __syncthreads();
      if (s_step[6] != 0.0) prev_cost = s_step[5];
      const bool stop = s_step[7] == 1.0;
      __syncthreads();
      if (stop) break;
    }
    __syncthreads();
  }
  // final inlier set of the refined pose
  int cnt = 0;
  for (int i = tid; i < N; i += PR_THREADS) {
    double u, v;
    bool in = false;
    if (project(s_pose, s_K, s3[i * 3], s3[i * 3 + 1], s3[i * 3 + 2], &u, &v)) {
      const double du = u - s2[i * 2], dv = v - s2[i * 2 + 1];
      in = du * du + dv * dv < thr2;
    }
    cnt += in ? 1 : 0;
    if (in && p.inlier_mask) p.inlier_mask[(long long)roi * npts + s_src[i]] = 1;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_scan[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int w = 0; w < PR_THREADS / 32; ++w) c += s_scan[w];
    if (p.n_inliers) p.n_inliers[roi] = c;
  }
  if (tid < 12) out[tid] = (float)s_pose[tid];
}

int launch_pnp(const PnpParams& p, int n, cudaStream_t st) {
  const size_t smem = (size_t)PR_MAX_PTS * (5 * sizeof(float) + sizeof(unsigned short) + 1);
  GDRN_OPT_IN_SMEM(pnp_ransac_kernel, smem);
  pnp_ransac_kernel<<<n, PR_THREADS, smem, st>>>(p);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

}  // namespace

extern "C" int gdrn_pnp_ransac_maps(const float* coor_x, const float* coor_y, const float* coor_z, const float* mask,
                                    const float* roi_coord_2d, const float* im_hw, const float* extents, const float* Ks,
                                    const int* idxs, int n, int hw, int iters, float mask_thr, float reproj_thr,
                                    unsigned seed, float* poses, int* n_inliers, unsigned char* inlier_mask, void* stream) {
  GDRN_REQUIRE(coor_x && coor_y && coor_z && mask && roi_coord_2d && im_hw && extents && Ks && poses, "pnp_ransac: null argument");
  GDRN_REQUIRE(n > 0 && hw > 0 && hw * hw <= PR_MAX_PTS, "pnp_ransac: need n > 0 and hw*hw <= 4096");
  GDRN_REQUIRE(iters >= 1 && iters <= PR_MAX_HYP, "pnp_ransac: iters must be in [1, 256]");
  PnpParams p = {};
  p.coor_x = coor_x; p.coor_y = coor_y; p.coor_z = coor_z; p.mask = mask; p.coord2d = roi_coord_2d; p.im_hw = im_hw;
  p.extents = extents; p.Ks = Ks; p.idxs = idxs; p.npts = hw * hw; p.iters = iters; p.mask_thr = mask_thr;
  p.reproj_thr = reproj_thr; p.seed = seed; p.poses = poses; p.n_inliers = n_inliers; p.inlier_mask = inlier_mask;
  return launch_pnp(p, n, (cudaStream_t)stream);
}

extern "C" int gdrn_pnp_ransac_points(const float* pts3d, const float* pts2d, const float* Ks, const int* idxs, int n,
                                      int npts, int iters, float reproj_thr, unsigned seed, float* poses, int* n_inliers,
                                      unsigned char* inlier_mask, void* stream) {
  GDRN_REQUIRE(pts3d && pts2d && Ks && poses, "pnp_ransac: null argument");
  GDRN_REQUIRE(n > 0 && npts >= 1 && npts <= PR_MAX_PTS, "pnp_ransac: need n > 0 and 1 <= npts <= 4096");
  GDRN_REQUIRE(iters >= 1 && iters <= PR_MAX_HYP, "pnp_ransac: iters must be in [1, 256]");
  PnpParams p = {};
  p.pts3d = pts3d; p.pts2d = pts2d; p.Ks = Ks; p.idxs = idxs; p.npts = npts; p.iters = iters; p.reproj_thr = reproj_thr;
  p.seed = seed; p.poses = poses; p.n_inliers = n_inliers; p.inlier_mask = inlier_mask;
  return launch_pnp(p, n, (cudaStream_t)stream);
}
