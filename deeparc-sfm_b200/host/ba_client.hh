// C++ binding of the C ABI (include/deeparc_ba.h) for the host mirror: the gather from the
// AoS pointer graph into the flat SoA image (by POINTER IDENTITY — Point3d ids go stale after
// filterPoint3d, reference DeepArcManager.cc:368-378), the scatter-back that emulates Ceres
// optimising in place through the raw double* (reference sfm.cc:47-48), and a process-wide
// engine handle.  This is the only place where the host code talks to CUDA, and it does so
// exclusively through the C ABI.
#ifndef DEEPARC_B200_BA_CLIENT_HH_
#define DEEPARC_B200_BA_CLIENT_HH_

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/deeparc_ba.h"
#include "DeepArcManager.hh"

namespace deeparc {

struct FlatProblem {
  std::vector<double> obs_xy, pts, ext_rot, ext_trans, intr_center, intr_focal, intr_dist;
  std::vector<int32_t> obs_pt, obs_pose_a, obs_pose_b, obs_intr, intr_nf, intr_nd;
  std::vector<uint8_t> ext_const;
  std::vector<ParameterBlock*> block_of;  // flat observation index -> object
  std::vector<Point3d*> point_of;      // flat point index -> object
  std::vector<Extrinsic*> ext_of;
  std::vector<Intrinsic*> intr_of;
  int freeze_camera = 0;
  dba_problem view() const;
};

// Problem construction of solve() (reference sfm.cc:36-65): one observation per
// ParameterBlock, parameter blocks from ParameterBlock::get(), the gauge rule (:50-53) as
// ext_const, freeze_camera (:54-57); intrinsics stay constant as shipped (:60-62).
void flatten(DeepArcManager& manager, bool freeze_camera, FlatProblem* out);

// Writes optimised values back into Point3d::position(), Extrinsic::rotation()/translation().
void scatter(const FlatProblem& flat, const std::vector<double>& pts, const std::vector<double>& ext_rot,
             const std::vector<double>& ext_trans);

// ---- device-resident scene: what the engine currently holds, in terms of the scene graph.
// solve() leaves the scene on the device; filterPoint3d() then decides on that copy without a second
// gather / upload, and the next solve() lets the engine drop the filtered observations / points itself
// (dba_problem_update) instead of flattening and uploading the scene again.  The shortcut is taken only
// when the scene graph still is what the engine was given: same manager, same object sequences (minus
// what the filter removed), parameters unchanged since the last scatter (all extrinsics and a sample of
// the points are compared).  DEEPARC_RESIDENT=0 disables it.
struct Resident {
  DeepArcManager* manager = nullptr;
  bool valid = false;
  int freeze_camera = 0;
  FlatProblem flat;                       // maps + last scattered parameter values
  bool pending = false;                   // flags decided by filterPoint3d, not yet applied on the device
  std::vector<uint8_t> obs_remove, pt_remove;
};
Resident& resident();
// true when the engine's copy can serve `manager` (after applying the pending removal, if any)
bool resident_usable(DeepArcManager& manager);
void resident_invalidate();
// applies the pending removal to the maps of the resident image (the engine side is dba_problem_update)
void resident_compact();

// Process-wide engine (device DEEPARC_DEVICE, default 0).  Throws std::runtime_error if the
// engine cannot be created (no GPU => no solve; there is no CPU fallback).
dba_handle* engine();
void engine_release();
void check(int status, const char* what);

}  // namespace deeparc

// Drop-in for the reference's solve() (sfm.cc:31-75).
void solve(DeepArcManager& deeparcManager, int max_iteration = 1000, int max_second = 3600,
           bool freeze_camera = false);

// Optional robust loss for the following solve() calls (the reference builds its problem with a NULL
// loss, sfm.cc:48, and keeps `new ceres::CauchyLoss(0.5)` in a comment, :49): scale <= 0 switches it off.
void solve_set_cauchy_loss(double scale);

// Last summary of solve() / of the hemisphere fit, for drivers that want more than stdout.
const dba_summary& last_solve_summary();
const std::vector<dba_iteration>& last_solve_iterations();

#endif  // DEEPARC_B200_BA_CLIENT_HH_
