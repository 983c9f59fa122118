// See ba_client.hh.
#include "ba_client.hh"

#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <unordered_map>

namespace deeparc {

dba_problem FlatProblem::view() const {
  dba_problem p;
  std::memset(&p, 0, sizeof p);
  p.n_obs = static_cast<int64_t>(obs_pt.size());
  p.n_pts = static_cast<int32_t>(point_of.size());
  p.n_ext = static_cast<int32_t>(ext_of.size());
  p.n_intr = static_cast<int32_t>(intr_of.size());
  p.obs_xy = obs_xy.data();
  p.obs_pt = obs_pt.data();
  p.obs_pose_a = obs_pose_a.data();
  p.obs_pose_b = obs_pose_b.data();
  p.obs_intr = obs_intr.data();
  p.pts = pts.data();
  p.ext_rot = ext_rot.data();
  p.ext_trans = ext_trans.data();
  p.intr_center = intr_center.data();
  p.intr_focal = intr_focal.data();
  p.intr_dist = intr_dist.data();
  p.intr_nf = intr_nf.data();
  p.intr_nd = intr_nd.data();
  p.ext_const = ext_const.data();
  p.freeze_camera = freeze_camera;
  p.free_intrinsics = 0;  // principal point, focal, distortion stay constant (sfm.cc:60-62)
  return p;
}

void flatten(DeepArcManager& m, bool freeze_camera, FlatProblem* out) {
  FlatProblem& f = *out;
  f = FlatProblem();
  f.freeze_camera = freeze_camera ? 1 : 0;
  std::vector<ParameterBlock*>& blocks = *m.parameters();
  std::vector<Point3d*>& points = *m.point3ds();
  std::vector<Extrinsic*>& exts = *m.extrinsics();
  std::vector<Intrinsic*>& intrs = *m.intrinsics();
  std::unordered_map<const Point3d*, int> pt_index;
  std::unordered_map<const Extrinsic*, int> ext_index;
  std::unordered_map<const Intrinsic*, int> intr_index;
  pt_index.reserve(points.size() * 2);
  f.point_of = points;
  f.pts.resize(3 * points.size());
  for (size_t i = 0; i < points.size(); ++i) {
    pt_index[points[i]] = static_cast<int>(i);
    for (int k = 0; k < 3; ++k) f.pts[3 * i + k] = points[i]->position()[k];
  }
  f.ext_of = exts;
  f.ext_rot.resize(3 * exts.size());
  f.ext_trans.resize(3 * exts.size());
  f.ext_const.assign(exts.size(), 0);
  for (size_t i = 0; i < exts.size(); ++i) {
    ext_index[exts[i]] = static_cast<int>(i);
    for (int k = 0; k < 3; ++k) {
      f.ext_rot[3 * i + k] = exts[i]->rotation()[k];
      f.ext_trans[3 * i + k] = exts[i]->translation()[k];
    }
  }
  f.intr_of = intrs;
  f.intr_center.resize(2 * intrs.size());
  f.intr_focal.assign(2 * intrs.size(), 0.0);
  f.intr_dist.assign(2 * intrs.size(), 0.0);
  f.intr_nf.resize(intrs.size());
  f.intr_nd.resize(intrs.size());
  for (size_t i = 0; i < intrs.size(); ++i) {
    intr_index[intrs[i]] = static_cast<int>(i);
    f.intr_center[2 * i] = intrs[i]->center()[0];
    f.intr_center[2 * i + 1] = intrs[i]->center()[1];
    f.intr_nf[i] = intrs[i]->focal_size();
    f.intr_nd[i] = intrs[i]->distrotion_size();
    for (int k = 0; k < 2; ++k) {
      if (k < intrs[i]->focal_size()) f.intr_focal[2 * i + k] = intrs[i]->focal()[k];
      if (k < intrs[i]->distrotion_size()) f.intr_dist[2 * i + k] = intrs[i]->distrotion()[k];
    }
  }
  const int64_t n = static_cast<int64_t>(blocks.size());
  f.block_of = blocks;
  f.obs_xy.resize(2 * static_cast<size_t>(n));
  f.obs_pt.resize(static_cast<size_t>(n));
  f.obs_pose_a.resize(static_cast<size_t>(n));
  f.obs_pose_b.resize(static_cast<size_t>(n));
  f.obs_intr.resize(static_cast<size_t>(n));
  // the per-observation gather runs on all host cores (the index maps are only read); an
  // observation that points outside the scene is reported like std::unordered_map::at()
  int64_t first_bad = n;
  const int n_thr = std::max(1, omp_get_max_threads());
  std::vector<std::vector<uint8_t> > gauge(static_cast<size_t>(n_thr), std::vector<uint8_t>(exts.size(), 0));
#pragma omp parallel for schedule(static) reduction(min : first_bad) num_threads(n_thr)
  for (int64_t i = 0; i < n; ++i) {
    ParameterBlock* b = blocks[i];
    f.obs_xy[2 * i] = b->point2d()->x();
    f.obs_xy[2 * i + 1] = b->point2d()->y();
    // the poses behind params[4],[5] (and [6],[7]) exactly as ParameterBlock::get() picks them
    Extrinsic *pa, *pb = nullptr;
    if (!b->share_extrinsic())
      pa = b->extrinsic();
    else if (b->pos_ring() == 0)
      pa = b->arc();
    else if (b->pos_arc() == 0)
      pa = b->ring();
    else {
      pa = b->arc();
      pb = b->ring();
    }
    const auto ip = pt_index.find(b->point3d());
    const auto ii = intr_index.find(b->intrinsic());
    const auto ia = ext_index.find(pa);
    const auto ib = pb ? ext_index.find(pb) : ext_index.end();
    if (ip == pt_index.end() || ii == intr_index.end() || ia == ext_index.end() || (pb && ib == ext_index.end())) {
      first_bad = std::min(first_bad, i);
      continue;
    }
    f.obs_pt[i] = ip->second;
    f.obs_intr[i] = ii->second;
    f.obs_pose_a[i] = ia->second;
    f.obs_pose_b[i] = pb ? ib->second : -1;
    // gauge: blocks at (arc 0, ring 0) pin params[4],[5] (sfm.cc:50-53)
    if (b->pos_arc() == 0 && b->pos_ring() == 0) gauge[omp_get_thread_num()][ia->second] = 1;
  }
  if (first_bad < n)
    throw std::out_of_range("flatten: observation " + std::to_string(first_bad) + " refers to an object outside the scene");
  for (const auto& g : gauge)
    for (size_t e = 0; e < exts.size(); ++e) f.ext_const[e] |= g[e];
}

void scatter(const FlatProblem& f, const std::vector<double>& pts, const std::vector<double>& ext_rot,
             const std::vector<double>& ext_trans) {
  for (size_t i = 0; i < f.point_of.size(); ++i)
    for (int k = 0; k < 3; ++k) f.point_of[i]->position()[k] = pts[3 * i + k];
  for (size_t i = 0; i < f.ext_of.size(); ++i) {
    f.ext_of[i]->rotation(&ext_rot[3 * i]);
    f.ext_of[i]->translation(ext_trans[3 * i], ext_trans[3 * i + 1], ext_trans[3 * i + 2]);
  }
}

namespace {
dba_handle* g_engine = nullptr;
Resident g_resident;
}

Resident& resident() { return g_resident; }
void resident_invalidate() { g_resident = Resident(); }

void resident_compact() {
  Resident& r = g_resident;
  if (!r.pending) return;
  FlatProblem& f = r.flat;
  const size_t n = f.block_of.size(), np = f.point_of.size();
  std::vector<int> new_pt(np + 1, 0);
  for (size_t i = 0; i < np; ++i) new_pt[i + 1] = new_pt[i] + (r.pt_remove[i] ? 0 : 1);
  size_t w = 0;
  for (size_t i = 0; i < n; ++i) {
    if (r.obs_remove[i] || r.pt_remove[f.obs_pt[i]]) continue;
    f.block_of[w] = f.block_of[i];
    f.obs_pt[w] = new_pt[f.obs_pt[i]];
    ++w;
  }
  f.block_of.resize(w);
  f.obs_pt.resize(w);
  size_t wp = 0;
  for (size_t i = 0; i < np; ++i) {
    if (r.pt_remove[i]) continue;
    f.point_of[wp] = f.point_of[i];
    for (int k = 0; k < 3; ++k) f.pts[3 * wp + k] = f.pts[3 * i + k];
    ++wp;
  }
  f.point_of.resize(wp);
  f.pts.resize(3 * wp);
  r.pending = false;
  r.obs_remove.clear();
  r.pt_remove.clear();
}

bool resident_usable(DeepArcManager& m) {
  Resident& r = g_resident;
  const char* env = std::getenv("DEEPARC_RESIDENT");
  if (env && env[0] == '0') return false;
  if (!r.valid || r.manager != &m || !g_engine) return false;
  const FlatProblem& f = r.flat;
  const std::vector<ParameterBlock*>& blocks = *m.parameters();
  const std::vector<Point3d*>& points = *m.point3ds();
  if (*m.extrinsics() != f.ext_of || *m.intrinsics() != f.intr_of) return false;
  // object sequences: the scene must hold exactly the survivors, in order
  size_t w = 0;
  for (size_t i = 0; i < f.block_of.size(); ++i) {
    if (r.pending && (r.obs_remove[i] || r.pt_remove[f.obs_pt[i]])) continue;
    if (w >= blocks.size() || blocks[w] != f.block_of[i]) return false;
    ++w;
  }
  if (w != blocks.size()) return false;
  size_t wp = 0;
  for (size_t i = 0; i < f.point_of.size(); ++i) {
    if (r.pending && r.pt_remove[i]) continue;
    if (wp >= points.size() || points[wp] != f.point_of[i]) return false;
    // parameters unchanged since the last scatter (sampled)
    if (i % 61 == 0)
      for (int k = 0; k < 3; ++k)
        if (points[wp]->position()[k] != f.pts[3 * i + k]) return false;
    ++wp;
  }
  if (wp != points.size()) return false;
  for (size_t i = 0; i < f.ext_of.size(); ++i)
    for (int k = 0; k < 3; ++k)
      if (f.ext_of[i]->rotation()[k] != f.ext_rot[3 * i + k] || f.ext_of[i]->translation()[k] != f.ext_trans[3 * i + k]) return false;
  return true;
}

dba_handle* engine() {
  if (g_engine) return g_engine;
  dba_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  const char* dev = std::getenv("DEEPARC_DEVICE");
  cfg.device = dev ? std::atoi(dev) : 0;
  cfg.rank = 0;
  cfg.world_size = 1;
  const int st = dba_create(&g_engine, &cfg);
  if (st != DBA_OK) {
    const std::string msg = std::string("deeparc: cannot create the GPU engine: ") + dba_last_error(nullptr);
    g_engine = nullptr;
    throw std::runtime_error(msg);
  }
  return g_engine;
}

void engine_release() {
  if (g_engine) dba_destroy(g_engine);
  g_engine = nullptr;
  g_resident = Resident();
}

void check(int status, const char* what) {
  if (status == DBA_OK) return;
  throw std::runtime_error(std::string(what) + " failed (" + std::to_string(status) + "): " +
                           (g_engine ? dba_last_error(g_engine) : dba_last_error(nullptr)));
}

}  // namespace deeparc
