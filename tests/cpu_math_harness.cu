// TEST INFRASTRUCTURE (never part of the product library): compiles the engine's analytic
// device math (deeparc-sfm_b200/csrc/ba_math.cuh, __host__ __device__) for the HOST so the
// `-m "not gpu"` suite can pin the formulas — pose-row hoisting, closed-form rotation
// derivative, projection chain rule — against the 50-digit golden vectors and the autodiff
// oracle without a GPU.  The GPU suite then checks that the kernels produce the same numbers.
#include <cstring>

#include "../deeparc-sfm_b200/csrc/ba_math.cuh"
#include "../include/deeparc_ba.h"

extern "C" int harness_eval(const dba_problem* p, double* residuals, double* jac_pt, double* jac_pose_a,
                            double* jac_pose_b, double* jac_intr) {
  using namespace dba;
  for (int64_t i = 0; i < p->n_obs; ++i) {
    const int a = p->obs_pose_a[i];
    const int b = p->obs_pose_b ? p->obs_pose_b[i] : -1;
    const int it = p->obs_intr[i];
    PoseRow A, B;
    make_pose_row(p->ext_rot + 3 * a, p->ext_trans + 3 * a, 1.0, &A);
    if (b >= 0) make_pose_row(p->ext_rot + 3 * b, p->ext_trans + 3 * b, 1.0, &B);
    IntrRow I;
    const int nf = p->intr_nf[it], nd = p->intr_nd[it];
    I.fx = p->intr_focal[2 * it];
    I.fy = nf == 2 ? p->intr_focal[2 * it + 1] : p->intr_focal[2 * it];
    I.cx = p->intr_center[2 * it];
    I.cy = p->intr_center[2 * it + 1];
    I.k0 = nd >= 1 ? p->intr_dist[2 * it] : 0.0;
    I.k1 = nd >= 2 ? p->intr_dist[2 * it + 1] : 0.0;
    I.nf = nf;
    I.nd = nd;
    ObsJacobian j;
    std::memset(&j, 0, sizeof j);
    observation_jacobian(A, b >= 0 ? &B : nullptr, I, p->pts + 3 * (size_t)p->obs_pt[i], p->obs_xy[2 * i],
                         p->obs_xy[2 * i + 1], true, j);
    if (residuals) {
      residuals[2 * i] = j.r0;
      residuals[2 * i + 1] = j.r1;
    }
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 3; ++c) {
        if (jac_pt) jac_pt[6 * i + 3 * r + c] = j.Jp[r][c];
        if (jac_pose_a) {
          jac_pose_a[12 * i + 6 * r + c] = j.JwA[r][c];
          jac_pose_a[12 * i + 6 * r + 3 + c] = j.JtA[r][c];
        }
        if (jac_pose_b) {
          jac_pose_b[12 * i + 6 * r + c] = b >= 0 ? j.JwB[r][c] : 0.0;
          jac_pose_b[12 * i + 6 * r + 3 + c] = b >= 0 ? j.JtB[r][c] : 0.0;
        }
      }
    if (jac_intr)
      for (int r = 0; r < 2; ++r) {
        jac_intr[6 * i + 3 * r + 0] = j.df[r];
        jac_intr[6 * i + 3 * r + 1] = j.dk0[r];
        jac_intr[6 * i + 3 * r + 2] = j.dk1[r];
      }
  }
  return 0;
}
