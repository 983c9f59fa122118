// TEST INFRASTRUCTURE — the CPU oracle.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; it is never
// linked into, imported by, or called from the product (deeparc-sfm_b200/).
//
// CPU restatement of the reference's bundle-adjustment hot path, in terms of the same
// flat SoA problem image the C ABI takes (include/deeparc_ba.h, struct dba_problem):
//   ReprojectionFunctor   <- src/snavely_reprojection_error.hh:38-118 (projectPoint,
//                            rotatePoint, operator())
//   MakeReprojectionCost  <- src/snavely_reprojection_error.hh:121-141 (Create)
//   SphereFunctor         <- src/hemisphere_radius.hh:18-28
//   oracle_solve          <- src/sfm.cc:31-75 (solve(): problem construction, gauge and
//                            freeze rules, solver options)
//   oracle_fit_hemisphere <- src/sfm.cc:86-103
//   oracle_filter_mse     <- src/DeepArcManager.cc:332-348
// The optimiser behind it is oracle/mini_ceres.cc (a restatement of ceres-solver, which
// is not installable here).  PARITY UNPINNED for solver semantics: the reference has no
// tests or golden vectors.  What IS pinned: the residual formula is
// checked against the reference's own functor compiled unmodified (oracle/_ref, see
// ref_bridge.cc) and against 50-digit mpmath vectors (tests/golden/).
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "ceres/ceres.h"
#include "ceres/rotation.h"
#include "deeparc_ba.h"

namespace {

constexpr int kAutodiffStride = 10;  // SNAVELY_REPROJECTION_KSTRIDE, snavely_reprojection_error.hh:4

// One observation.  Parameter-block order == ParameterBlock::get() (ParameterBlock.hh:68-94):
// [0] point(3) [1] centre(2) [2] focal(nf) [3] distortion(nd) [4] rot [5] trans ([6] rot_b [7] trans_b)
struct ReprojectionFunctor {
  double obs_x, obs_y;
  int nf, nd;
  bool two_poses;

  // rotatePoint, snavely_reprojection_error.hh:80-91
  template <typename T>
  static void Transform(const T* rot, const T* trans, const T* in, T* out) {
    ceres::AngleAxisRotatePoint(rot, in, out);
    out[0] = out[0] + trans[0];
    out[1] = out[1] + trans[1];
    out[2] = out[2] + trans[2];
  }

  template <typename T>
  bool operator()(T const* const* blk, T* res) const {
    T cam[3];
    if (two_poses) {  // :96-108 — the [6],[7] pose (ring) acts first, then [4],[5] (arc)
      T mid[3];
      Transform(blk[6], blk[7], blk[0], mid);
      Transform(blk[4], blk[5], mid, cam);
    } else {  // :110-115
      Transform(blk[4], blk[5], blk[0], cam);
    }
    // projectPoint :38-78.  No sign flip on the perspective division (:49-50).
    const T* centre = blk[1];
    const T* focal = blk[2];
    const T* dist = blk[3];
    const T u = cam[0] / cam[2];
    const T v = cam[1] / cam[2];
    const T fx = focal[0];
    const T fy = (nf == 2) ? focal[1] : focal[0];  // :53-55
    T radial = T(1.0);
    const T rr = u * u + v * v;
    if (nd == 2) radial = 1.0 + rr * (dist[0] + dist[1] * rr);  // :61-63
    if (nd == 1) radial = 1.0 + rr * dist[0];                   // :65-67
    const T px = fx * radial * u + centre[0];                   // :71-72
    const T py = fy * radial * v + centre[1];
    res[0] = px - obs_x;                                        // :75-76
    res[1] = py - obs_y;
    return true;
  }
};

typedef ceres::DynamicAutoDiffCostFunction<ReprojectionFunctor, kAutodiffStride> ReprojectionCost;

// snavely_reprojection_error.hh:121-141
ReprojectionCost* MakeReprojectionCost(double x, double y, int nf, int nd, bool two_poses) {
  ReprojectionCost* c = new ReprojectionCost(new ReprojectionFunctor{x, y, nf, nd, two_poses});
  c->AddParameterBlock(3);
  c->AddParameterBlock(2);
  c->AddParameterBlock(nf);
  c->AddParameterBlock(nd);
  c->AddParameterBlock(3);
  c->AddParameterBlock(3);
  if (two_poses) {
    c->AddParameterBlock(3);
    c->AddParameterBlock(3);
  }
  c->SetNumResiduals(2);
  return c;
}

// hemisphere_radius.hh:18-28 — note rho is the squared radius.
struct SphereFunctor {
  double pos[3];
  template <typename T>
  bool operator()(const T* centre, const T* rho, T* res) const {
    T acc = T(0.0);
    for (int i = 0; i < 3; ++i) {
      const T d = centre[i] - pos[i];
      acc += d * d;
    }
    res[0] = acc - rho[0];
    return true;
  }
};

// Mutable host copy of a dba_problem.
struct HostProblem {
  int64_t n_obs;
  int n_pts, n_ext, n_intr;
  std::vector<double> obs_xy, pts, ext_rot, ext_trans, centre, focal, dist;
  std::vector<int32_t> obs_pt, pose_a, pose_b, obs_intr, nf, nd;
  std::vector<uint8_t> ext_const;
  bool freeze_camera, free_intrinsics;

  explicit HostProblem(const dba_problem& p)
      : n_obs(p.n_obs), n_pts(p.n_pts), n_ext(p.n_ext), n_intr(p.n_intr),
        obs_xy(p.obs_xy, p.obs_xy + 2 * p.n_obs), pts(p.pts, p.pts + 3 * (size_t)p.n_pts),
        ext_rot(p.ext_rot, p.ext_rot + 3 * (size_t)p.n_ext),
        ext_trans(p.ext_trans, p.ext_trans + 3 * (size_t)p.n_ext),
        centre(p.intr_center, p.intr_center + 2 * (size_t)p.n_intr),
        focal(p.intr_focal, p.intr_focal + 2 * (size_t)p.n_intr),
        dist(p.intr_dist, p.intr_dist + 2 * (size_t)p.n_intr),
        obs_pt(p.obs_pt, p.obs_pt + p.n_obs), pose_a(p.obs_pose_a, p.obs_pose_a + p.n_obs),
        pose_b(p.n_obs, -1), obs_intr(p.obs_intr, p.obs_intr + p.n_obs),
        nf(p.intr_nf, p.intr_nf + p.n_intr), nd(p.intr_nd, p.intr_nd + p.n_intr),
        ext_const(p.n_ext, 0), freeze_camera(p.freeze_camera != 0),
        free_intrinsics(p.free_intrinsics != 0) {
    if (p.obs_pose_b) pose_b.assign(p.obs_pose_b, p.obs_pose_b + p.n_obs);
    if (p.ext_const) ext_const.assign(p.ext_const, p.ext_const + p.n_ext);
  }

  // the get() pointer list of observation i
  std::vector<double*> Blocks(int64_t i) {
    const int it = obs_intr[i];
    std::vector<double*> b = {&pts[3 * (size_t)obs_pt[i]], &centre[2 * it], &focal[2 * it], &dist[2 * it],
                              &ext_rot[3 * (size_t)pose_a[i]], &ext_trans[3 * (size_t)pose_a[i]]};
    if (pose_b[i] >= 0) {
      b.push_back(&ext_rot[3 * (size_t)pose_b[i]]);
      b.push_back(&ext_trans[3 * (size_t)pose_b[i]]);
    }
    return b;
  }
};

bool Validate(const dba_problem* p) {
  if (!p || p->n_obs < 0 || p->n_pts < 0 || p->n_ext < 0 || p->n_intr < 0) return false;
  if (p->n_obs > 0 && (!p->obs_xy || !p->obs_pt || !p->obs_pose_a || !p->obs_intr)) return false;
  for (int64_t i = 0; i < p->n_obs; ++i) {
    if (p->obs_pt[i] < 0 || p->obs_pt[i] >= p->n_pts) return false;
    if (p->obs_pose_a[i] < 0 || p->obs_pose_a[i] >= p->n_ext) return false;
    if (p->obs_intr[i] < 0 || p->obs_intr[i] >= p->n_intr) return false;
    if (p->obs_pose_b && (p->obs_pose_b[i] < -1 || p->obs_pose_b[i] >= p->n_ext)) return false;
  }
  return true;
}

void CopySummary(const ceres::Solver::Summary& cs, dba_summary* s) {
  if (!s) return;
  dba_iteration* it_buf = s->iterations;
  const int cap = s->iterations_capacity;
  std::memset(s, 0, sizeof *s);
  s->iterations = it_buf;
  s->iterations_capacity = cap;
  s->termination = cs.termination_type == ceres::CONVERGENCE
                       ? DBA_CONVERGENCE
                       : (cs.termination_type == ceres::NO_CONVERGENCE ? DBA_NO_CONVERGENCE : DBA_FAILURE);
  s->num_successful_steps = cs.num_successful_steps;
  s->num_unsuccessful_steps = cs.num_unsuccessful_steps;
  s->linear_solver_used = cs.linear_solver_type_used == ceres::ITERATIVE_SCHUR ? DBA_LS_PCG : DBA_LS_DENSE;
  s->reduced_system_size = cs.reduced_system_size;
  s->initial_cost = cs.initial_cost;
  s->final_cost = cs.final_cost;
  s->total_time_in_seconds = cs.total_time_in_seconds;
  s->jacobian_evaluations = cs.num_jacobian_evaluations;
  s->residual_evaluations = cs.num_residual_evaluations;
  std::snprintf(s->message, sizeof s->message, "%s", cs.message.c_str());
  int n = 0;
  for (const auto& ci : cs.iterations) {
    s->pcg_iterations_total += ci.linear_solver_iterations;
    if (it_buf && n < cap) {
      dba_iteration& d = it_buf[n];
      d.iteration = ci.iteration;
      d.step_is_valid = ci.step_is_valid;
      d.step_is_successful = ci.step_is_successful;
      d.linear_solver_iterations = ci.linear_solver_iterations;
      d.cost = ci.cost;
      d.cost_change = ci.cost_change;
      d.gradient_max_norm = ci.gradient_max_norm;
      d.gradient_norm = ci.gradient_norm;
      d.step_norm = ci.step_norm;
      d.relative_decrease = ci.relative_decrease;
      d.trust_region_radius = ci.trust_region_radius;
      d.model_cost_change = ci.model_cost_change;
      d.iteration_time_in_seconds = ci.iteration_time_in_seconds;
      ++n;
    }
  }
  s->num_iterations = n;
}

ceres::Solver::Options ToCeres(const dba_solve_options* o, int num_threads) {
  ceres::Solver::Options c;
  c.linear_solver_type = ceres::DENSE_SCHUR;  // sfm.cc:67
  c.num_threads = num_threads;
  if (!o) return c;
  c.max_num_iterations = o->max_num_iterations;
  c.max_solver_time_in_seconds = o->max_solver_time_in_seconds;
  c.initial_trust_region_radius = o->initial_trust_region_radius;
  c.max_trust_region_radius = o->max_trust_region_radius;
  c.min_trust_region_radius = o->min_trust_region_radius;
  c.min_relative_decrease = o->min_relative_decrease;
  c.min_lm_diagonal = o->min_lm_diagonal;
  c.max_lm_diagonal = o->max_lm_diagonal;
  c.function_tolerance = o->function_tolerance;
  c.gradient_tolerance = o->gradient_tolerance;
  c.parameter_tolerance = o->parameter_tolerance;
  c.jacobi_scaling = o->jacobi_scaling != 0;
  c.max_num_consecutive_invalid_steps = o->max_num_consecutive_invalid_steps;
  c.minimizer_progress_to_stdout = o->progress_to_stdout != 0;
  if (o->linear_solver == DBA_LS_PCG) {
    c.linear_solver_type = ceres::ITERATIVE_SCHUR;
    c.preconditioner_type = ceres::SCHUR_JACOBI;
    c.max_linear_solver_iterations = o->pcg_max_iterations;
    c.min_linear_solver_iterations = o->pcg_min_iterations;
    c.shim_pcg_rel_tol = o->pcg_rel_tolerance;
  }
  return c;
}

}  // namespace

extern "C" {

// Residuals and autodiff Jacobians of every observation, in the caller's order.
// Layouts as dba_eval (include/deeparc_ba.h).  Raw (unscaled, unmasked) derivatives.
int oracle_eval(const dba_problem* p, double* cost, double* residuals, double* jac_pt,
                double* jac_pose_a, double* jac_pose_b, double* jac_intr) {
  if (!Validate(p)) return DBA_ERR_INVALID_ARGUMENT;
  HostProblem hp(*p);
  const bool want_jac = jac_pt || jac_pose_a || jac_pose_b || jac_intr;
  double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int64_t i = 0; i < hp.n_obs; ++i) {
    const int it = hp.obs_intr[i];
    const bool two = hp.pose_b[i] >= 0;
    std::unique_ptr<ReprojectionCost> c(MakeReprojectionCost(hp.obs_xy[2 * i], hp.obs_xy[2 * i + 1], hp.nf[it], hp.nd[it], two));
    std::vector<double*> blk = hp.Blocks(i);
    double r[2];
    double J0[6], J1[4], J2[4], J3[4], J4[6], J5[6], J6[6], J7[6];
    double* jac[8] = {J0, J1, J2, J3, J4, J5, J6, J7};
    c->Evaluate(blk.data(), r, want_jac ? jac : NULL);
    total += r[0] * r[0] + r[1] * r[1];
    if (residuals) {
      residuals[2 * i] = r[0];
      residuals[2 * i + 1] = r[1];
    }
    if (jac_pt) std::memcpy(jac_pt + 6 * i, J0, 6 * sizeof(double));
    if (jac_pose_a)
      for (int k = 0; k < 2; ++k)
        for (int c2 = 0; c2 < 3; ++c2) {
          jac_pose_a[12 * i + 6 * k + c2] = J4[3 * k + c2];
          jac_pose_a[12 * i + 6 * k + 3 + c2] = J5[3 * k + c2];
        }
    if (jac_pose_b)
      for (int k = 0; k < 2; ++k)
        for (int c2 = 0; c2 < 3; ++c2) {
          jac_pose_b[12 * i + 6 * k + c2] = two ? J6[3 * k + c2] : 0.0;
          jac_pose_b[12 * i + 6 * k + 3 + c2] = two ? J7[3 * k + c2] : 0.0;
        }
    if (jac_intr) {
      const int nf = hp.nf[it], nd = hp.nd[it];
      for (int k = 0; k < 2; ++k) {
        // nf = 2: column 0 carries the diagonal of the 2x2 focal Jacobian, (d r0/d fx, d r1/d fy);
        // its off-diagonal entries d r0/d fy, d r1/d fx are identically zero (snavely...hh:53-55, :71-72)
        jac_intr[6 * i + 3 * k + 0] = J2[nf * k + (nf == 2 ? k : 0)];
        jac_intr[6 * i + 3 * k + 1] = nd >= 1 ? J3[nd * k + 0] : 0.0;
        jac_intr[6 * i + 3 * k + 2] = nd >= 2 ? J3[nd * k + 1] : 0.0;
      }
    }
  }
  if (cost) *cost = 0.5 * total;
  return DBA_OK;
}

// sfm.cc:31-75 on the flat problem image.  Outputs as dba_params_get.
int oracle_solve(const dba_problem* p, const dba_solve_options* o, dba_summary* s, int num_threads,
                 double* pts, double* ext_rot, double* ext_trans, double* intr_focal,
                 double* intr_dist) {
  if (!Validate(p)) return DBA_ERR_INVALID_ARGUMENT;
  HostProblem hp(*p);
  ceres::Problem problem;
  for (int64_t i = 0; i < hp.n_obs; ++i) {
    const int it = hp.obs_intr[i];
    const bool two = hp.pose_b[i] >= 0;
    ceres::CostFunction* c = MakeReprojectionCost(hp.obs_xy[2 * i], hp.obs_xy[2 * i + 1], hp.nf[it], hp.nd[it], two);
    std::vector<double*> blk = hp.Blocks(i);
    // sfm.cc:48 passes NULL; sfm.cc:49 keeps `new ceres::CauchyLoss(0.5)` in a comment
    ceres::LossFunction* loss = (o && o->loss_type == DBA_LOSS_CAUCHY) ? new ceres::CauchyLoss(o->loss_scale) : NULL;
    problem.AddResidualBlock(c, loss, blk);
    if (hp.ext_const[hp.pose_a[i]]) {        // gauge rule, resolved by the caller into a mask
      problem.SetParameterBlockConstant(blk[4]);
      problem.SetParameterBlockConstant(blk[5]);
    }
    if (two && hp.ext_const[hp.pose_b[i]]) {
      problem.SetParameterBlockConstant(blk[6]);
      problem.SetParameterBlockConstant(blk[7]);
    }
    // implicit-Schur PCG mode only (a shim extension, used as the same-algorithm CPU baseline and for the
    // fixed-K parity test): the preconditioner blocks are per camera, as in the GPU engine — rotation and
    // translation of one extrinsic together, plus focal / distortion of its intrinsic when those are free
    problem.SetParameterBlockPreconditionerGroup(blk[4], hp.pose_a[i]);
    problem.SetParameterBlockPreconditionerGroup(blk[5], hp.pose_a[i]);
    if (two) {
      problem.SetParameterBlockPreconditionerGroup(blk[6], hp.pose_b[i]);
      problem.SetParameterBlockPreconditionerGroup(blk[7], hp.pose_b[i]);
    }
    if (hp.free_intrinsics) {
      problem.SetParameterBlockPreconditionerGroup(blk[2], hp.pose_a[i]);
      if (hp.nd[it] > 0) problem.SetParameterBlockPreconditionerGroup(blk[3], hp.pose_a[i]);
    }
    if (hp.freeze_camera) {  // sfm.cc:54-57
      for (size_t j = 1; j < blk.size(); ++j) problem.SetParameterBlockConstant(blk[j]);
    } else {                 // sfm.cc:58-63
      problem.SetParameterBlockConstant(blk[1]);
      if (!hp.free_intrinsics) {
        problem.SetParameterBlockConstant(blk[2]);
        problem.SetParameterBlockConstant(blk[3]);
      }
    }
  }
  ceres::Solver::Options options = ToCeres(o, num_threads > 0 ? num_threads : omp_get_num_procs());
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  CopySummary(summary, s);
  if (pts) std::memcpy(pts, hp.pts.data(), hp.pts.size() * sizeof(double));
  if (ext_rot) std::memcpy(ext_rot, hp.ext_rot.data(), hp.ext_rot.size() * sizeof(double));
  if (ext_trans) std::memcpy(ext_trans, hp.ext_trans.data(), hp.ext_trans.size() * sizeof(double));
  if (intr_focal) std::memcpy(intr_focal, hp.focal.data(), hp.focal.size() * sizeof(double));
  if (intr_dist) std::memcpy(intr_dist, hp.dist.data(), hp.dist.size() * sizeof(double));
  return DBA_OK;
}

// sfm.cc:86-103
int oracle_fit_hemisphere(const double* centres, int n, double centre_io[3], double* rho_io,
                          const dba_solve_options* o, dba_summary* s, int num_threads) {
  if (!centres || n < 0 || !centre_io || !rho_io) return DBA_ERR_INVALID_ARGUMENT;
  ceres::Problem problem;
  for (int i = 0; i < n; ++i) {
    SphereFunctor* f = new SphereFunctor{{centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]}};
    problem.AddResidualBlock(new ceres::AutoDiffCostFunction<SphereFunctor, 1, 3, 1>(f), NULL, centre_io, rho_io);
  }
  ceres::Solver::Options options = ToCeres(o, num_threads > 0 ? num_threads : 1);
  options.linear_solver_type = ceres::DENSE_SCHUR;  // sfm.cc:95
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  CopySummary(summary, s);
  return DBA_OK;
}

// DeepArcManager.cc:335-347: mse = (r0^2 + r1^2) / 2 per observation, plain doubles.
int oracle_filter_mse(const dba_problem* p, double* mse) {
  if (!Validate(p) || !mse) return DBA_ERR_INVALID_ARGUMENT;
  HostProblem hp(*p);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < hp.n_obs; ++i) {
    const int it = hp.obs_intr[i];
    ReprojectionFunctor f{hp.obs_xy[2 * i], hp.obs_xy[2 * i + 1], hp.nf[it], hp.nd[it], hp.pose_b[i] >= 0};
    std::vector<double*> blk = hp.Blocks(i);
    double r[2];
    f(blk.data(), r);
    mse[i] = (r[0] * r[0] + r[1] * r[1]) / 2.0;
  }
  return DBA_OK;
}

// rotation helpers (column-major 3x3; quaternion w,x,y,z) for the loader tests
void oracle_angle_axis_rotate_point(const double* aa, const double* pt, double* out) {
  ceres::AngleAxisRotatePoint(aa, pt, out);
}
void oracle_angle_axis_to_rotation_matrix(const double* aa, double* R) { ceres::AngleAxisToRotationMatrix(aa, R); }
void oracle_rotation_matrix_to_angle_axis(const double* R, double* aa) { ceres::RotationMatrixToAngleAxis(R, aa); }
void oracle_quaternion_to_angle_axis(const double* q, double* aa) { ceres::QuaternionToAngleAxis(q, aa); }

int oracle_num_procs(void) { return omp_get_num_procs(); }

}  // extern "C"
