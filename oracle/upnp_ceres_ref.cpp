// ORACLE (test infrastructure, NOT product code): uncertainty-PnP solved with the REFERENCE's own vendored Ceres 2.0
// machinery -- ceres::Jet forward-mode autodiff (include/ceres/jet.h), ceres::AngleAxisRotatePoint
// (include/ceres/rotation.h) and Ceres' Levenberg-Marquardt ceres::TinySolver (include/ceres/tiny_solver.h) -- compiled
// from the headers where they lie under /root/reference/core/csrc/uncertainty_pnp/include (oracle/build_ref.py; output
// oracle/_ref/libupnp_ceres_ref.so).  The reference's uncertainty_pnp.cpp itself links libceres.so (ceres::Problem /
// ceres::Solve, DENSE_SCHUR), which is not in the repository nor on the system; the residual below restates its
// ReprojectionErrorArray::operator() (core/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:16-34) term for term, stacked
// over the points.  Both solvers are LM on the same autodiff residual, so they agree at the minimum; this pins the
// numpy oracle (ops_oracle.uncertainty_pnp) and the CUDA kernel (csrc/upnp.cu) against Ceres' arithmetic.
#include <cstring>

#include "ceres/jet.h"
#include "ceres/rotation.h"
#include "ceres/tiny_solver.h"
#include "ceres/tiny_solver_autodiff_function.h"

namespace {

struct StackedReprojection {
  const double* pts2d; const double* pts3d; const double* wgt2d; const double* K; int pn;
  int NumResiduals() const { return 2 * pn; }
  template <typename T>
  bool operator()(const T* const pose, T* residuals) const {
    const double fx = K[0], fy = K[4], px = K[2], py = K[5];
    for (int i = 0; i < pn; ++i) {
      T p[3] = {T(pts3d[i * 3]), T(pts3d[i * 3 + 1]), T(pts3d[i * 3 + 2])};
      T q[3];
      ceres::AngleAxisRotatePoint(pose, p, q);
      q[0] += pose[3]; q[1] += pose[4]; q[2] += pose[5];
      const T proj_x = T(fx) * q[0] / q[2] + T(px);
      const T proj_y = T(fy) * q[1] / q[2] + T(py);
      const T dx = proj_x - T(pts2d[i * 2]), dy = proj_y - T(pts2d[i * 2 + 1]);
      const double wxx = wgt2d[i * 3], wxy = wgt2d[i * 3 + 1], wyy = wgt2d[i * 3 + 2];
      residuals[2 * i] = T(wxx) * dx + T(wxy) * dy;
      residuals[2 * i + 1] = T(wxy) * dx + T(wyy) * dy;
    }
    return true;
  }
};

}  // namespace

extern "C" int upnp_ceres_ref(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                              const double* init_rt, double* result_rt, int pn, double* final_cost) {
  StackedReprojection f{pts2d, pts3d, wgt2d, K, pn};
  typedef ceres::TinySolverAutoDiffFunction<StackedReprojection, Eigen::Dynamic, 6> AD;
  AD ad(f);
  ceres::TinySolver<AD> solver;
  solver.options.max_num_iterations = 100;
  solver.options.gradient_tolerance = 1e-12;
  solver.options.parameter_tolerance = 1e-14;
  solver.options.cost_threshold = 1e-30;
  Eigen::Matrix<double, 6, 1> x;
  for (int i = 0; i < 6; ++i) x[i] = init_rt[i];
  auto summary = solver.Solve(ad, &x);
  for (int i = 0; i < 6; ++i) result_rt[i] = x[i];
  if (final_cost) *final_cost = summary.final_cost;
  return summary.iterations;
}
