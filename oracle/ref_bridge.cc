// TEST INFRASTRUCTURE — bridge that exposes the reference's OWN code through a C API.
//
// This translation unit is compiled together with the reference sources where they lie
// (/root/reference/src/sfm.cc and src/DeepArcManager.cc, unmodified, never copied into
// this repository) against the mini-Ceres / mini-Eigen shim in oracle/ceres_shim, into
// oracle/_ref/libdeeparc_ref.so (recipe: oracle/Makefile, target `ref`).  It lets the
// tests run, on the same inputs as the CUDA engine:
//   - the reference's SnavelyReprojectionError functor and its autodiff cost function
//     (snavely_reprojection_error.hh:93-141), HemisphereRadius (hemisphere_radius.hh),
//   - the reference's DeepArcManager::read / write / writePly / filterPoint3d /
//     getCameraCenter (DeepArcManager.cc), and
//   - the reference's solve() (sfm.cc:31-75) — with the optimiser being the shim.
// Only the Ceres/Eigen layer underneath is a restatement; everything above it is the
// reference itself.  /root/reference does not exist on the GPU box: the prebuilt .so
// travels there, nothing here opens the reference tree at run time.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ceres/ceres.h"
#include "ceres/rotation.h"

// The manager keeps its intrinsic/extrinsic tables private and offers no accessor;
// the bridge needs their file order to export a flat problem image.
#define private public
#include "DeepArcManager.hh"
#undef private
#include "hemisphere_radius.hh"
#include "snavely_reprojection_error.hh"

#include "deeparc_ba.h"

// defined in the reference's sfm.cc (default arguments live on the definition only)
void solve(DeepArcManager& deeparcManager, int max_iteration, int max_second, bool freeze_camera);

namespace {

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

}  // namespace

extern "C" {

// ---- shim control ---------------------------------------------------------------
#ifndef DEEPARC_REAL_CERES
void ref_set_overrides(int quiet, int num_threads, int max_num_iterations, double function_tolerance,
                       double gradient_tolerance, double parameter_tolerance) {
  ceres::shim::Overrides& o = ceres::shim::GlobalOverrides();
  o.quiet = quiet;
  o.num_threads = num_threads;
  o.max_num_iterations = max_num_iterations;
  o.function_tolerance = function_tolerance;
  o.gradient_tolerance = gradient_tolerance;
  o.parameter_tolerance = parameter_tolerance;
}
void ref_last_summary(dba_summary* s) { CopySummary(ceres::shim::LastSummary(), s); }
int ref_is_real_ceres(void) { return 0; }
#else
// WITH_CERES build (oracle/Makefile target ref_ceres): the reference's solve() runs on the real
// ceres-solver exactly as shipped.  Its options cannot be overridden and its summary is a local of
// solve() (sfm.cc:72-74, printed, never returned), so only the resulting scene is comparable:
// tests then check final parameters / files against the restatement at the reference's own settings.
void ref_set_overrides(int, int, int, double, double, double) {}
void ref_last_summary(dba_summary* s) {
  if (!s) return;
  dba_iteration* it = s->iterations;
  const int cap = s->iterations_capacity;
  std::memset(s, 0, sizeof *s);
  s->iterations = it;
  s->iterations_capacity = cap;
  s->termination = DBA_FAILURE;
  std::snprintf(s->message, sizeof s->message, "summary not available: the reference's solve() keeps it local (real Ceres build)");
}
int ref_is_real_ceres(void) { return 1; }
#endif

// ---- the reference functor, called directly ---------------------------------------
// params = ParameterBlock::get() order; jac may be NULL, else 8 pointers (row-major blocks).
int ref_functor(double x, double y, int nf, int nd, int compose, const double* const* params,
                double* residuals, double** jac) {
  if (jac == NULL) {
    SnavelyReprojectionError f(x, y, nf, nd, compose != 0);
    return f(params, residuals) ? 0 : -1;
  }
  SnavelyCostFunction* c = SnavelyReprojectionError::Create(x, y, nf, nd, compose != 0);
  const bool ok = c->Evaluate(params, residuals, jac);
  delete c;
  return ok ? 0 : -1;
}

// Same contract as oracle_eval / dba_eval, evaluated with the reference's cost function.
int ref_eval(const dba_problem* p, double* cost, double* residuals, double* jac_pt, double* jac_pose_a,
             double* jac_pose_b, double* jac_intr) {
  const bool want_jac = jac_pt || jac_pose_a || jac_pose_b || jac_intr;
  double total = 0.0;
  for (int64_t i = 0; i < p->n_obs; ++i) {
    const int it = p->obs_intr[i];
    const int pb = p->obs_pose_b ? p->obs_pose_b[i] : -1;
    const double* blk[8] = {p->pts + 3 * (size_t)p->obs_pt[i], p->intr_center + 2 * it, p->intr_focal + 2 * it,
                            p->intr_dist + 2 * it, p->ext_rot + 3 * (size_t)p->obs_pose_a[i],
                            p->ext_trans + 3 * (size_t)p->obs_pose_a[i],
                            pb >= 0 ? p->ext_rot + 3 * (size_t)pb : NULL, pb >= 0 ? p->ext_trans + 3 * (size_t)pb : NULL};
    double r[2];
    double J0[6], J1[4], J2[4], J3[4], J4[6], J5[6], J6[6], J7[6];
    double* jac[8] = {J0, J1, J2, J3, J4, J5, J6, J7};
    const int nf = p->intr_nf[it], nd = p->intr_nd[it];
    if (ref_functor(p->obs_xy[2 * i], p->obs_xy[2 * i + 1], nf, nd, pb >= 0, blk, r, want_jac ? jac : NULL) != 0)
      return DBA_ERR_NUMERIC;
    total += r[0] * r[0] + r[1] * r[1];
    if (residuals) {
      residuals[2 * i] = r[0];
      residuals[2 * i + 1] = r[1];
    }
    if (jac_pt) std::memcpy(jac_pt + 6 * i, J0, 6 * sizeof(double));
    if (jac_pose_a)
      for (int k = 0; k < 2; ++k)
        for (int c = 0; c < 3; ++c) {
          jac_pose_a[12 * i + 6 * k + c] = J4[3 * k + c];
          jac_pose_a[12 * i + 6 * k + 3 + c] = J5[3 * k + c];
        }
    if (jac_pose_b)
      for (int k = 0; k < 2; ++k)
        for (int c = 0; c < 3; ++c) {
          jac_pose_b[12 * i + 6 * k + c] = pb >= 0 ? J6[3 * k + c] : 0.0;
          jac_pose_b[12 * i + 6 * k + 3 + c] = pb >= 0 ? J7[3 * k + c] : 0.0;
        }
    if (jac_intr)
      for (int k = 0; k < 2; ++k) {
        // nf = 2: column 0 carries the diagonal of the 2x2 focal Jacobian, (d r0/d fx, d r1/d fy);
        // its off-diagonal entries d r0/d fy, d r1/d fx are identically zero (snavely...hh:53-55, :71-72)
        jac_intr[6 * i + 3 * k + 0] = J2[nf * k + (nf == 2 ? k : 0)];
        jac_intr[6 * i + 3 * k + 1] = nd >= 1 ? J3[nd * k + 0] : 0.0;
        jac_intr[6 * i + 3 * k + 2] = nd >= 2 ? J3[nd * k + 1] : 0.0;
      }
  }
  if (cost) *cost = 0.5 * total;
  return DBA_OK;
}

// The reference's hemisphere problem (sfm.cc:86-101) with the reference's cost function.
int ref_fit_hemisphere(const double* centres, int n, double centre_io[3], double* rho_io, dba_summary* s) {
  ceres::Problem problem;
  for (int i = 0; i < n; ++i) {
    double pos[3] = {centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]};
    ceres::CostFunction* cost_fn = HemisphereRadius::Create(pos);
    problem.AddResidualBlock(cost_fn, NULL, centre_io, rho_io);
  }
  ceres::Solver::Options options;
  options.linear_solver_type = ceres::DENSE_SCHUR;
  options.minimizer_progress_to_stdout = true;
  options.max_num_iterations = 1000;
  options.num_threads = 20;
  options.max_solver_time_in_seconds = 3600;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  CopySummary(summary, s);
  return DBA_OK;
}

// ---- the reference DeepArcManager ---------------------------------------------------
void* ref_manager_read(const char* path) {
  DeepArcManager* m = new DeepArcManager();
  try {
    m->read(path);
  } catch (const char*) {
    delete m;
    return NULL;
  }
  return m;
}
void ref_manager_free(void* h) { delete static_cast<DeepArcManager*>(h); }
int ref_manager_is_shared(void* h) { return static_cast<DeepArcManager*>(h)->isShareExtrinsic() ? 1 : 0; }
void ref_manager_counts(void* h, int64_t* n_obs, int* n_pts, int* n_ext, int* n_intr, int* n_arc, int* n_ring) {
  DeepArcManager* m = static_cast<DeepArcManager*>(h);
  *n_obs = static_cast<int64_t>(m->parameters()->size());
  *n_pts = static_cast<int>(m->point3ds()->size());
  *n_ext = static_cast<int>(m->extrinsics_.size());
  *n_intr = static_cast<int>(m->intrinsics_.size());
  *n_arc = m->arc_size_;
  *n_ring = m->share_extrinsic_ ? m->ring_size_ : 0;
}

// Flat export by POINTER IDENTITY (point ids go stale after filterPoint3d).  Arrays sized by
// ref_manager_counts.  ext_const follows the gauge rule of sfm.cc:50-53.
int ref_manager_export(void* h, double* obs_xy, int32_t* obs_pt, int32_t* obs_pose_a, int32_t* obs_pose_b,
                       int32_t* obs_intr, double* pts, int32_t* pts_rgb, double* ext_rot, double* ext_trans,
                       double* intr_center, double* intr_focal, double* intr_dist, int32_t* intr_nf,
                       int32_t* intr_nd, uint8_t* ext_const) {
  DeepArcManager* m = static_cast<DeepArcManager*>(h);
  std::map<Point3d*, int> pt_index;
  std::map<Extrinsic*, int> ext_index;
  std::map<Intrinsic*, int> intr_index;
  for (size_t i = 0; i < m->point3ds()->size(); ++i) {
    Point3d* q = m->point3ds()->at(i);
    pt_index[q] = static_cast<int>(i);
    for (int k = 0; k < 3; ++k) pts[3 * i + k] = q->position()[k];
    if (pts_rgb) {
      pts_rgb[3 * i] = q->r();
      pts_rgb[3 * i + 1] = q->g();
      pts_rgb[3 * i + 2] = q->b();
    }
  }
  for (size_t i = 0; i < m->extrinsics_.size(); ++i) {
    Extrinsic* e = m->extrinsics_[i];
    ext_index[e] = static_cast<int>(i);
    for (int k = 0; k < 3; ++k) {
      ext_rot[3 * i + k] = e->rotation()[k];
      ext_trans[3 * i + k] = e->translation()[k];
    }
    if (ext_const) ext_const[i] = 0;
  }
  for (size_t i = 0; i < m->intrinsics_.size(); ++i) {
    Intrinsic* q = m->intrinsics_[i];
    intr_index[q] = static_cast<int>(i);
    intr_center[2 * i] = q->center()[0];
    intr_center[2 * i + 1] = q->center()[1];
    intr_nf[i] = q->focal_size();
    intr_nd[i] = q->distrotion_size();
    for (int k = 0; k < 2; ++k) {
      intr_focal[2 * i + k] = k < q->focal_size() ? q->focal()[k] : 0.0;
      intr_dist[2 * i + k] = k < q->distrotion_size() ? q->distrotion()[k] : 0.0;
    }
  }
  std::map<double*, int> rot_owner;
  for (auto& kv : ext_index) rot_owner[kv.first->rotation()] = kv.second;
  for (size_t i = 0; i < m->parameters()->size(); ++i) {
    ParameterBlock* b = m->parameters()->at(i);
    obs_xy[2 * i] = b->point2d()->x();
    obs_xy[2 * i + 1] = b->point2d()->y();
    obs_pt[i] = pt_index.at(b->point3d());
    obs_intr[i] = intr_index.at(b->intrinsic());
    std::vector<double*> blk = b->get();
    obs_pose_a[i] = rot_owner.at(blk[4]);
    obs_pose_b[i] = blk.size() > 6 ? rot_owner.at(blk[6]) : -1;
    if (ext_const && b->pos_arc() == 0 && b->pos_ring() == 0) ext_const[obs_pose_a[i]] = 1;
  }
  return 0;
}

// reference solve() (sfm.cc:31-75), unmodified; trace via ref_last_summary
void ref_manager_solve(void* h, int max_iteration, int max_second, int freeze_camera) {
  solve(*static_cast<DeepArcManager*>(h), max_iteration, max_second, freeze_camera != 0);
}
void ref_manager_filter(void* h, double error_boundary, double* centre, double radius) {
  static_cast<DeepArcManager*>(h)->filterPoint3d(error_boundary, centre, radius);
}
void ref_manager_write(void* h, const char* path) { static_cast<DeepArcManager*>(h)->write(path); }
void ref_manager_write_ply(void* h, const char* path) { static_cast<DeepArcManager*>(h)->writePly(path); }
int ref_manager_camera_centers(void* h, double* out, int capacity) {
  std::vector<std::vector<double> > c = static_cast<DeepArcManager*>(h)->getCameraCenter();
  const int n = static_cast<int>(c.size());
  for (int i = 0; i < n && i < capacity; ++i)
    for (int k = 0; k < 3; ++k) out[3 * i + k] = c[i][k];
  return n;
}

}  // extern "C"
