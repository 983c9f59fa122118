// C entry points onto the C++ host mirror (DeepArcManager, solve()), so that the Python test
// suite can drive this library and the reference's own sources (oracle/ref_bridge.cc, same
// function set with the prefix ref_) side by side on the same files.
#include <cstring>
#include <vector>

#include "DeepArcManager.hh"
#include "ba_client.hh"

namespace {
thread_local std::string g_error;
template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const char* m) {
    g_error = m;
  } catch (const std::exception& e) {
    g_error = e.what();
  }
  return -1;
}
}  // namespace

extern "C" {

const char* dam_last_error(void) { return g_error.c_str(); }

void* dam_manager_read(const char* path) {
  DeepArcManager* m = new DeepArcManager();
  if (guarded([&] { m->read(path); }) != 0) {
    delete m;
    return NULL;
  }
  return m;
}
void dam_manager_free(void* h) { delete static_cast<DeepArcManager*>(h); }
int dam_manager_is_shared(void* h) { return static_cast<DeepArcManager*>(h)->isShareExtrinsic() ? 1 : 0; }
void dam_manager_counts(void* h, int64_t* n_obs, int* n_pts, int* n_ext, int* n_intr, int* n_arc, int* n_ring) {
  DeepArcManager* m = static_cast<DeepArcManager*>(h);
  *n_obs = static_cast<int64_t>(m->parameters()->size());
  *n_pts = static_cast<int>(m->point3ds()->size());
  *n_ext = static_cast<int>(m->extrinsics()->size());
  *n_intr = static_cast<int>(m->intrinsics()->size());
  *n_arc = m->arc_size();
  *n_ring = m->ring_size();
}

// flat export through the same gather the solver uses (deeparc::flatten)
int dam_manager_export(void* h, double* obs_xy, int32_t* obs_pt, int32_t* obs_pose_a, int32_t* obs_pose_b,
                       int32_t* obs_intr, double* pts, int32_t* pts_rgb, double* ext_rot, double* ext_trans,
                       double* intr_center, double* intr_focal, double* intr_dist, int32_t* intr_nf,
                       int32_t* intr_nd, uint8_t* ext_const) {
  DeepArcManager* m = static_cast<DeepArcManager*>(h);
  return guarded([&] {
    deeparc::FlatProblem f;
    deeparc::flatten(*m, false, &f);
    auto cp = [](auto* dst, const auto& src) {
      if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
    };
    cp(obs_xy, f.obs_xy);
    cp(obs_pt, f.obs_pt);
    cp(obs_pose_a, f.obs_pose_a);
    cp(obs_pose_b, f.obs_pose_b);
    cp(obs_intr, f.obs_intr);
    cp(pts, f.pts);
    cp(ext_rot, f.ext_rot);
    cp(ext_trans, f.ext_trans);
    cp(intr_center, f.intr_center);
    cp(intr_focal, f.intr_focal);
    cp(intr_dist, f.intr_dist);
    cp(intr_nf, f.intr_nf);
    cp(intr_nd, f.intr_nd);
    cp(ext_const, f.ext_const);
    if (pts_rgb)
      for (size_t i = 0; i < f.point_of.size(); ++i) {
        pts_rgb[3 * i] = f.point_of[i]->r();
        pts_rgb[3 * i + 1] = f.point_of[i]->g();
        pts_rgb[3 * i + 2] = f.point_of[i]->b();
      }
  });
}

int dam_manager_solve(void* h, int max_iteration, int max_second, int freeze_camera) {
  return guarded([&] { solve(*static_cast<DeepArcManager*>(h), max_iteration, max_second, freeze_camera != 0); });
}
void dam_last_summary(dba_summary* s) {
  if (!s) return;
  dba_iteration* buf = s->iterations;
  const int cap = s->iterations_capacity;
  *s = last_solve_summary();
  s->iterations = buf;
  s->iterations_capacity = cap;
  const std::vector<dba_iteration>& its = last_solve_iterations();
  int n = 0;
  for (; n < static_cast<int>(its.size()) && buf && n < cap; ++n) buf[n] = its[n];
  s->num_iterations = n;
}
int dam_manager_filter(void* h, double error_boundary, double* centre, double radius) {
  return guarded([&] { static_cast<DeepArcManager*>(h)->filterPoint3d(error_boundary, centre, radius); });
}
void dam_manager_write(void* h, const char* path) { static_cast<DeepArcManager*>(h)->write(path); }
void dam_manager_write_binary(void* h, const char* path) { static_cast<DeepArcManager*>(h)->writeBinary(path); }
void dam_manager_write_ply(void* h, const char* path) { static_cast<DeepArcManager*>(h)->writePly(path); }
int dam_manager_camera_centers(void* h, double* out, int capacity) {
  std::vector<std::vector<double> > c = static_cast<DeepArcManager*>(h)->getCameraCenter();
  const int n = static_cast<int>(c.size());
  for (int i = 0; i < n && i < capacity; ++i)
    for (int k = 0; k < 3; ++k) out[3 * i + k] = c[i][k];
  return n;
}
void dam_engine_release(void) { deeparc::engine_release(); }
// test hook: the writer's "%.6f" (fast path + snprintf fallback) into out[>= 400]; returns the length
int dam_format_fixed6(double v, char* out) {
  const std::string s = deeparc_format_fixed6(v);
  std::memcpy(out, s.c_str(), s.size() + 1);
  return static_cast<int>(s.size());
}

}  // extern "C"
