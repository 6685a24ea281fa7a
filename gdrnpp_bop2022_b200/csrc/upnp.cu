// Uncertainty-weighted PnP refinement (6-dof angle-axis + t), batched on the GPU.
// Replaces core/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:7-92: residual functor (:16-34),
// ceres::AutoDiffCostFunction<.,2,6> + ceres::Solve with default options and DENSE_SCHUR.
//
// libceres is not part of the reference tree; its default trust-region Levenberg-Marquardt loop
// (Ceres 2.0.0: trust_region_minimizer.cc / levenberg_marquardt_strategy.cc, defaults from
// solver.h: max_num_iterations 50, initial_trust_region_radius 1e4, max radius 1e16,
// min_relative_decrease 1e-3, function_tolerance 1e-6, gradient_tolerance 1e-10,
// parameter_tolerance 1e-8, min/max_lm_diagonal 1e-6/1e32, jacobi_scaling on) is restated here.
// Derivatives use forward-mode duals with 6 partials, i.e. what AutoDiff evaluates.
// One warp per problem: lanes stride over the points, J^T J / J^T r / cost are reduced with a fixed
// shuffle tree (deterministic), every lane then solves the 6x6 system redundantly.
#include <math.h>

#include "common.cuh"

namespace {

struct Jet6 {
  double v;
  double d[6];
};
__device__ __forceinline__ Jet6 jconst(double c) { Jet6 r; r.v = c; for (int i = 0; i < 6; ++i) r.d[i] = 0.0; return r; }
__device__ __forceinline__ Jet6 jvar(double c, int k) { Jet6 r = jconst(c); r.d[k] = 1.0; return r; }
__device__ __forceinline__ Jet6 operator+(const Jet6& a, const Jet6& b) { Jet6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ Jet6 operator-(const Jet6& a, const Jet6& b) { Jet6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ Jet6 operator*(const Jet6& a, const Jet6& b) { Jet6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
__device__ __forceinline__ Jet6 operator/(const Jet6& a, const Jet6& b) {
  Jet6 r; const double inv = 1.0 / b.v; r.v = a.v * inv;
  for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ Jet6 jsqrt(const Jet6& a) { Jet6 r; r.v = sqrt(a.v); const double t = 1.0 / (2.0 * r.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * t; return r; }
__device__ __forceinline__ Jet6 jcos(const Jet6& a) { Jet6 r; r.v = cos(a.v); const double s = -sin(a.v); for (int i = 0; i < 6; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ Jet6 jsin(const Jet6& a) { Jet6 r; r.v = sin(a.v); const double c = cos(a.v); for (int i = 0; i < 6; ++i) r.d[i] = c * a.d[i]; return r; }

// residuals (2) and their 2x6 Jacobian for one point (uncertainty_pnp.cpp:16-34 with ceres::AngleAxisRotatePoint)
__device__ void point_residual(const double* pose, const double* p2, const double* p3, const double* w, double fx,
                               double fy, double px, double py, double* res, double* J) {
  Jet6 aa[3] = {jvar(pose[0], 0), jvar(pose[1], 1), jvar(pose[2], 2)};
  Jet6 t[3] = {jvar(pose[3], 3), jvar(pose[4], 4), jvar(pose[5], 5)};
  Jet6 pt[3] = {jconst(p3[0]), jconst(p3[1]), jconst(p3[2])};
  Jet6 q[3];
  const Jet6 theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2.v > 2.220446049250313e-16) {  // std::numeric_limits<double>::epsilon()
    const Jet6 theta = jsqrt(theta2);
    const Jet6 ct = jcos(theta), st = jsin(theta);
    const Jet6 itheta = jconst(1.0) / theta;
    const Jet6 w0 = aa[0] * itheta, w1 = aa[1] * itheta, w2 = aa[2] * itheta;
    const Jet6 c0 = w1 * pt[2] - w2 * pt[1], c1 = w2 * pt[0] - w0 * pt[2], c2 = w0 * pt[1] - w1 * pt[0];
    const Jet6 tmp = (w0 * pt[0] + w1 * pt[1] + w2 * pt[2]) * (jconst(1.0) - ct);
    q[0] = pt[0] * ct + c0 * st + w0 * tmp;
    q[1] = pt[1] * ct + c1 * st + w1 * tmp;
    q[2] = pt[2] * ct + c2 * st + w2 * tmp;
  } else {
    q[0] = pt[0] + (aa[1] * pt[2] - aa[2] * pt[1]);
    q[1] = pt[1] + (aa[2] * pt[0] - aa[0] * pt[2]);
    q[2] = pt[2] + (aa[0] * pt[1] - aa[1] * pt[0]);
  }
  q[0] = q[0] + t[0]; q[1] = q[1] + t[1]; q[2] = q[2] + t[2];
  const Jet6 prx = jconst(fx) * q[0] / q[2] + jconst(px);
  const Jet6 pry = jconst(fy) * q[1] / q[2] + jconst(py);
  const Jet6 dx = prx - jconst(p2[0]), dy = pry - jconst(p2[1]);
  const Jet6 r0 = jconst(w[0]) * dx + jconst(w[1]) * dy;
  const Jet6 r1 = jconst(w[1]) * dx + jconst(w[2]) * dy;
  res[0] = r0.v; res[1] = r1.v;
  for (int i = 0; i < 6; ++i) { J[i] = r0.d[i]; J[6 + i] = r1.d[i]; }
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cost = 0.5 * sum r^2 ; optionally g = J^T r and H = J^T J (upper, 21 entries row-major i<=j)
__device__ double evaluate(const double* pose, const double* p2, const double* p3, const double* w, const double* K,
                           int pn, int lane, double* g, double* Hm) {
  double cost = 0.0, lg[6], lH[21];
  for (int i = 0; i < 6; ++i) lg[i] = 0.0;
  for (int i = 0; i < 21; ++i) lH[i] = 0.0;
  for (int i = lane; i < pn; i += 32) {
    double res[2], J[12];
    point_residual(pose, p2 + 2 * i, p3 + 3 * i, w + 3 * i, K[0], K[4], K[2], K[5], res, J);
    cost += 0.5 * (res[0] * res[0] + res[1] * res[1]);
    if (g) {
      int k = 0;
      for (int a = 0; a < 6; ++a) {
        lg[a] += J[a] * res[0] + J[6 + a] * res[1];
        for (int b = a; b < 6; ++b) lH[k++] += J[a] * J[b] + J[6 + a] * J[6 + b];
      }
    }
  }
  cost = warp_sum(cost);
  if (g) {
    for (int a = 0; a < 6; ++a) g[a] = warp_sum(lg[a]);
    for (int k = 0; k < 21; ++k) Hm[k] = warp_sum(lH[k]);
  }
  return cost;
}

// solve (A) x = b for symmetric positive definite 6x6 A (full storage) by Cholesky; returns false if not SPD
__device__ bool chol_solve6(double A[6][6], const double* b, double* x) {
  double L[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
  return true;
}

__global__ void __launch_bounds__(128)
upnp_kernel(const double* __restrict__ pts2d, const double* __restrict__ pts3d, const double* __restrict__ wgt2d,
            const double* __restrict__ Kall, const double* __restrict__ init_rt, double* __restrict__ result_rt, int pn,
            int n_problems) {
  const int prob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (prob >= n_problems) return;
  const double* p2 = pts2d + (size_t)prob * pn * 2;
  const double* p3 = pts3d + (size_t)prob * pn * 3;
  const double* w = wgt2d + (size_t)prob * pn * 3;
  const double* K = Kall + (size_t)prob * 9;
  double x[6];
  for (int i = 0; i < 6; ++i) x[i] = init_rt[prob * 6 + i];

  double g[6], Hu[21];
  double cost = evaluate(x, p2, p3, w, K, pn, lane, g, Hu);
  // Jacobi scaling, fixed at the first Jacobian: s_j = 1 / (1 + ||J_j||)
  double scale[6];
  {
    int k = 0;
    for (int a = 0; a < 6; ++a) { scale[a] = 1.0 / (1.0 + sqrt(Hu[k])); k += 6 - a; }
  }
  double radius = 1e4, decrease_factor = 2.0;
  const double max_radius = 1e16, min_radius = 1e-32;
  const double min_rel_decrease = 1e-3, ftol = 1e-6, gtol = 1e-10, ptol = 1e-8;
  {  // gradient tolerance at start: max |g| (unconstrained)
    double gmax = 0.0;
    for (int a = 0; a < 6; ++a) gmax = fmax(gmax, fabs(g[a]));
    if (gmax <= gtol) { for (int i = 0; i < 6; ++i) if (lane == 0) result_rt[prob * 6 + i] = x[i]; return; }
  }
  for (int iter = 0; iter < 50; ++iter) {
    // scaled system: Js = J * diag(scale)
    double A[6][6], gs[6];
    {
      int k = 0;
      for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { A[a][b] = A[b][a] = Hu[k++] * scale[a] * scale[b]; }
      for (int a = 0; a < 6; ++a) gs[a] = g[a] * scale[a];
    }
    // LM diagonal: D^2 = clamp(diag(Js^T Js), min, max) / radius
    double lm[6];
    for (int a = 0; a < 6; ++a) lm[a] = fmin(fmax(A[a][a], 1e-6), 1e32) / radius;
    double Areg[6][6], rhs[6], step[6];
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b) Areg[a][b] = A[a][b];
      Areg[a][a] += lm[a];
      rhs[a] = -gs[a];
    }
    bool ok = chol_solve6(Areg, rhs, step);
    bool step_ok = false;
    double new_cost = cost, xn[6];
    if (ok) {
      // model cost change = -step^T (gs + 0.5 * A step)   (on the un-regularised Gauss-Newton model)
      double mc = 0.0;
      for (int a = 0; a < 6; ++a) {
        double As = 0.0;
        for (int b = 0; b < 6; ++b) As += A[a][b] * step[b];
        mc -= step[a] * (gs[a] + 0.5 * As);
      }
      double delta[6], dn = 0.0, xnorm = 0.0;
      for (int a = 0; a < 6; ++a) { delta[a] = step[a] * scale[a]; dn += delta[a] * delta[a]; xnorm += x[a] * x[a]; xn[a] = x[a] + delta[a]; }
      dn = sqrt(dn); xnorm = sqrt(xnorm);
      if (mc > 0.0) {
        if (dn <= ptol * (xnorm + ptol)) break;  // parameter tolerance
        new_cost = evaluate(xn, p2, p3, w, K, pn, lane, nullptr, nullptr);
        const double rel = (cost - new_cost) / mc;
        step_ok = rel > min_rel_decrease;
        if (step_ok) {
          const double cost_change = cost - new_cost;
          for (int a = 0; a < 6; ++a) x[a] = xn[a];
          radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3.0));
          radius = fmin(max_radius, radius);
          decrease_factor = 2.0;
          const double prev = cost;
          cost = evaluate(x, p2, p3, w, K, pn, lane, g, Hu);
          double gmax = 0.0;
          for (int a = 0; a < 6; ++a) gmax = fmax(gmax, fabs(g[a]));
          if (gmax <= gtol) break;
          if (fabs(cost_change) <= ftol * prev) break;  // function tolerance
        }
      }
    }
    if (!step_ok) {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      if (radius < min_radius) break;
    }
  }
  if (lane == 0)
    for (int i = 0; i < 6; ++i) result_rt[prob * 6 + i] = x[i];
}

int upnp_launch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* init_rt,
                double* result_rt, int pn, int n_problems, cudaStream_t st) {
  GDRN_REQUIRE(pn >= 1 && n_problems >= 1, "upnp: need pn >= 1 and n_problems >= 1");
  const int warps_per_block = 4;
  upnp_kernel<<<(n_problems + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, st>>>(
      pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn, n_problems);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

}  // namespace

extern "C" int upnp_batched(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                            const double* init_rt, double* result_rt, int pn, int n_problems, void* stream) {
  return upnp_launch(pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn, n_problems, (cudaStream_t)stream);
}

// Host-pointer entry with the reference's cffi signature (uncertainty_pnp/src/ext.h); blocking.
extern "C" void uncertainty_pnp(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_rt,
                                double* result_rt, int pn) {
  for (int i = 0; i < 6; ++i) result_rt[i] = init_rt[i];
  if (pn <= 0) return;
  const size_t n2 = (size_t)pn * 2, n3 = (size_t)pn * 3;
  const size_t total = n2 + n3 + n3 + 9 + 6 + 6;
  double* d = nullptr;
  if (cudaMalloc(&d, total * 8) != cudaSuccess) { fprintf(stderr, "gdrn uncertainty_pnp: cudaMalloc failed\n"); return; }
  double *d2 = d, *d3 = d2 + n2, *dw = d3 + n3, *dK = dw + n3, *di = dK + 9, *dr = di + 6;
  cudaMemcpy(d2, pts2d, n2 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d3, pts3d, n3 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, wgt2d, n3 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dK, K, 72, cudaMemcpyHostToDevice);
  cudaMemcpy(di, init_rt, 48, cudaMemcpyHostToDevice);
  int rc = upnp_launch(d2, d3, dw, dK, di, dr, pn, 1, 0);
  cudaError_t e = cudaMemcpy(result_rt, dr, 48, cudaMemcpyDeviceToHost);
  if (rc != GDRN_OK || e != cudaSuccess)
    fprintf(stderr, "gdrn uncertainty_pnp failed: %s\n", e != cudaSuccess ? cudaGetErrorString(e) : gdrn_last_error());
  cudaFree(d);
}
