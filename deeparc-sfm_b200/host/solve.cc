// solve(): drop-in for the reference's solve() (reference src/sfm.cc:31-75).  Same signature
// and defaults; instead of building a ceres::Problem and calling ceres::Solve it gathers the
// scene into the flat SoA image, runs the GPU Levenberg-Marquardt engine through the C ABI and
// scatters the result back into the objects the raw double* of ParameterBlock::get() point at —
// which is what Ceres' in-place optimisation leaves behind for filterPoint3d / write / writePly.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>

#include "ba_client.hh"

namespace {
dba_summary g_summary;
std::vector<dba_iteration> g_iterations;
double g_cauchy_scale = 0.0;  // <= 0: no loss, as the reference ships (sfm.cc:48)

void print_full_report(const dba_summary& s, const deeparc::FlatProblem& f, bool freeze) {
  std::printf("\nSolver Summary (deeparc B200 engine)\n\n");
  std::printf("Observations       %12zu\nPoints             %12zu\nExtrinsics         %12zu\nIntrinsics         %12zu (constant)\n",
              f.block_of.size(), f.point_of.size(), f.ext_of.size(), f.intr_of.size());
  std::printf("Camera parameters  %12s\n", freeze ? "constant" : "6 per free extrinsic");
  std::printf("Linear solver      %s, reduced system %d\n",
              s.linear_solver_used == DBA_LS_DENSE ? "explicit Schur + dense Cholesky" : "implicit Schur + block-Jacobi PCG",
              s.reduced_system_size);
  if (s.linear_solver_failures) std::printf("Linear solver failures (step rejected) %d\n", s.linear_solver_failures);
  if (s.pcg_unconverged_solves) std::printf("WARNING: %d PCG solve(s) stopped at the iteration cap before reaching the tolerance\n", s.pcg_unconverged_solves);
  std::printf("\nCost:\nInitial          %30e\nFinal            %30e\nChange           %30e\n\n", s.initial_cost,
              s.final_cost, s.initial_cost - s.final_cost);
  std::printf("Minimizer iterations %d\nSuccessful steps     %d\nUnsuccessful steps   %d\n\n", s.num_iterations,
              s.num_successful_steps, s.num_unsuccessful_steps);
  std::printf("Time (in seconds):\n  Device LM loop   %12.6f\n  Total            %12.6f\nKernel launches    %lld\n\n",
              s.device_time_in_seconds, s.total_time_in_seconds, static_cast<long long>(s.kernel_launches));
  std::printf("Termination: %s (%s)\n",
              s.termination == DBA_CONVERGENCE ? "CONVERGENCE" : (s.termination == DBA_NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE"),
              s.message);
}
}  // namespace

void solve_set_cauchy_loss(double scale) { g_cauchy_scale = scale; }
const dba_summary& last_solve_summary() { return g_summary; }
const std::vector<dba_iteration>& last_solve_iterations() { return g_iterations; }

void solve(DeepArcManager& deeparcManager, int max_iteration, int max_second, bool freeze_camera) {
  dba_handle* h = deeparc::engine();
  deeparc::Resident& res = deeparc::resident();
  if (deeparc::resident_usable(deeparcManager)) {
    // the engine still holds this scene (left there by the previous solve() / filterPoint3d()): drop what
    // the filter removed and switch the freeze flag on the device side, no gather / sort / upload
    if (res.pending || res.freeze_camera != (freeze_camera ? 1 : 0)) {
      deeparc::check(dba_problem_update(h, res.pending ? res.obs_remove.data() : nullptr, res.pending ? res.pt_remove.data() : nullptr,
                                        freeze_camera ? 1 : 0, nullptr, nullptr),
                     "dba_problem_update");
      deeparc::resident_compact();
      res.freeze_camera = freeze_camera ? 1 : 0;
      res.flat.freeze_camera = res.freeze_camera;
    }
  } else {
    deeparc::resident_invalidate();
    deeparc::flatten(deeparcManager, freeze_camera, &res.flat);  // sfm.cc:36-65
    dba_problem view = res.flat.view();
    deeparc::check(dba_problem_set(h, &view), "dba_problem_set");
    res.manager = &deeparcManager;
    res.freeze_camera = freeze_camera ? 1 : 0;
    res.valid = true;
    // the observation arrays live on the device now; the maps stay for the scatter-back and the filter
    std::vector<double>().swap(res.flat.obs_xy);
    std::vector<int32_t>().swap(res.flat.obs_pose_a);
    std::vector<int32_t>().swap(res.flat.obs_pose_b);
    std::vector<int32_t>().swap(res.flat.obs_intr);
  }
  deeparc::FlatProblem& flat = res.flat;

  dba_solve_options options;
  dba_solve_options_default(&options);            // Ceres defaults (not overridden at sfm.cc:66-71)
  // sfm.cc:67 DENSE_SCHUR: the explicit reduced system + device Cholesky whenever the camera side
  // fits it (every rig the reference was written for: 19..50 pose blocks); only scenes with
  // hundreds of free cameras fall back to the PCG, driven to the exact step, and the report says so
  options.linear_solver = DBA_LS_AUTO;
  options.progress_to_stdout = 1;                 // sfm.cc:68
  options.max_num_iterations = max_iteration;     // sfm.cc:69
  options.max_solver_time_in_seconds = max_second;  // sfm.cc:71
  options.pcg_rel_tolerance = 1e-13;
  options.pcg_max_iterations = 4000;
  if (g_cauchy_scale > 0.0) {                     // sfm.cc:49 (commented out in the reference)
    options.loss_type = DBA_LOSS_CAUCHY;
    options.loss_scale = g_cauchy_scale;
  }
  // (sfm.cc:70 num_threads has no meaning here)
  g_iterations.assign(static_cast<size_t>(max_iteration) + 2, dba_iteration());
  std::memset(&g_summary, 0, sizeof g_summary);
  g_summary.iterations = g_iterations.data();
  g_summary.iterations_capacity = static_cast<int32_t>(g_iterations.size());
  deeparc::check(dba_solve(h, &options, &g_summary), "dba_solve");  // sfm.cc:73
  g_iterations.resize(static_cast<size_t>(g_summary.num_iterations));

  // scatter-back; the resident image remembers the values so that a later call can tell whether the
  // scene graph was edited behind the engine's back
  deeparc::check(dba_params_get(h, flat.pts.data(), flat.ext_rot.data(), flat.ext_trans.data(), nullptr, nullptr), "dba_params_get");
  deeparc::scatter(flat, flat.pts, flat.ext_rot, flat.ext_trans);
  print_full_report(g_summary, flat, freeze_camera);  // sfm.cc:74
}
