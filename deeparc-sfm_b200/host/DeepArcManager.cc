// See DeepArcManager.hh.  Behavioural contract = reference src/DeepArcManager.cc; every
// function names the lines it mirrors.  Quirks that a drop-in must keep are marked QUIRK.
#include "DeepArcManager.hh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "ba_client.hh"
#include "rotation_conv.hh"

namespace {

// Whitespace-separated token reader over the whole file image (the format is plain text:
// DeepArcManager.cc:36-48, :81, :103-118, :133-147, :159).  strtod/strtol on one buffer is
// ~10x faster than iostream extraction, which matters at 5-50 M observation lines.
class Tokens {
 public:
  explicit Tokens(std::string data) : data_(std::move(data)), p_(data_.c_str()) {}
  bool nextDouble(double* v) {
    skip();
    if (!*p_) return false;
    char* end = nullptr;
    *v = std::strtod(p_, &end);
    if (end == p_) return false;
    p_ = end;
    return true;
  }
  // istream >> int semantics: parse an integer prefix of the token
  bool nextInt(int* v) {
    skip();
    if (!*p_) return false;
    char* end = nullptr;
    const long x = std::strtol(p_, &end, 10);
    if (end == p_) return false;
    *v = static_cast<int>(x);
    p_ = end;
    return true;
  }

 private:
  void skip() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r' || *p_ == '\f' || *p_ == '\v') ++p_;
  }
  std::string data_;
  const char* p_;
};

void fmt6(std::string* out, double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.6f", v);
  out->append(buf);
}
void fmtg(std::string* out, double v) {  // default ostream formatting of a double (precision 6)
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", v);
  out->append(buf);
}

}  // namespace

DeepArcManager::DeepArcManager() : arc_size_(0), ring_size_(0), share_extrinsic_(false) {}

DeepArcManager::~DeepArcManager() {
  for (ParameterBlock* b : params_) delete b;
  for (Point3d* p : point3d_) delete p;
  for (Intrinsic* i : intrinsics_) delete i;
  for (Extrinsic* e : extrinsics_) delete e;
  for (Camera* c : camera_) delete c;
  for (auto& row : hemisphere_)
    for (auto& cell : row.second) delete cell.second;
}

bool DeepArcManager::isShareExtrinsic() { return share_extrinsic_; }
std::vector<ParameterBlock*>* DeepArcManager::parameters() { return &params_; }
std::vector<Point3d*>* DeepArcManager::point3ds() { return &point3d_; }

// ring position -> slot in the extrinsic table; ring 0 aliases slot 0 (DeepArcManager.cc:166-171)
int DeepArcManager::ringSlot(int ring_position, int arc_size) {
  return ring_position == 0 ? 0 : ring_position + arc_size - 1;
}

// DeepArcManager.cc:26-74
bool DeepArcManager::read(std::string filename) {
  std::ifstream file(filename, std::ios::binary);
  if (file.fail()) {
    std::cout << "Cannot read " << filename << std::endl;
    throw "Cannot read input file";
  }
  std::ostringstream ss;
  ss << file.rdbuf();
  file.close();
  Tokens tok(ss.str());

  double version = 0.0;
  int n_block = 0, n_intrinsic = 0, n_arc = 0, n_ring = 0, n_point = 0;
  tok.nextDouble(&version);
  tok.nextInt(&n_block);
  tok.nextInt(&n_intrinsic);
  tok.nextInt(&n_arc);
  tok.nextInt(&n_ring);
  tok.nextInt(&n_point);
  share_extrinsic_ = n_ring != 0;
  arc_size_ = n_arc;
  ring_size_ = n_ring;
  const int n_extrinsic = n_ring != 0 ? n_arc + n_ring - 1 : n_arc;

  // observations: pos_arc pos_ring point3d_id x y   (:76-91)
  params_.reserve(params_.size() + static_cast<size_t>(std::max(n_block, 0)));
  for (int i = 0; i < n_block; ++i) {
    int a = 0, r = 0, pid = 0;
    double x = 0.0, y = 0.0;
    tok.nextInt(&a);
    tok.nextInt(&r);
    tok.nextInt(&pid);
    tok.nextDouble(&x);
    tok.nextDouble(&y);
    params_.push_back(new ParameterBlock(a, r, pid, new Point2d(x, y)));
  }
  // intrinsics: cx cy nf f.. nd k..   (:93-122)
  std::vector<Intrinsic*> intrinsics;
  for (int i = 0; i < n_intrinsic; ++i) {
    Intrinsic* in = new Intrinsic();
    in->id(i);
    double cx = 0.0, cy = 0.0, f[2] = {0.0, 0.0}, k[2] = {0.0, 0.0}, hold = 0.0;
    int nf = 0, nd = 0;
    tok.nextDouble(&cx);
    tok.nextDouble(&cy);
    in->center(static_cast<int>(cx), static_cast<int>(cy));  // QUIRK: truncated principal point
    tok.nextInt(&nf);
    for (int j = 0; j < nf; ++j) {
      tok.nextDouble(&hold);
      if (j < 2) f[j] = hold;  // the reference overruns a 2-element buffer here; we drop extras
    }
    in->focal(nf, f);
    tok.nextInt(&nd);
    for (int j = 0; j < nd; ++j) {
      tok.nextDouble(&hold);
      if (j < 2) k[j] = hold;
    }
    in->distrotion(nd, k);
    intrinsics.push_back(in);
  }
  // extrinsics: tx ty tz nrot rot..   (:124-151); nrot 3 angle-axis, 4 quaternion wxyz,
  // 9 COLUMN-MAJOR matrix
  std::vector<Extrinsic*> extrinsics;
  for (int i = 0; i < n_extrinsic; ++i) {
    Extrinsic* ex = new Extrinsic();
    ex->id(i);
    double t[3] = {0, 0, 0}, rot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, aa[3] = {0, 0, 0}, hold = 0.0;
    int nrot = 0;
    tok.nextDouble(&t[0]);
    tok.nextDouble(&t[1]);
    tok.nextDouble(&t[2]);
    ex->translation(t[0], t[1], t[2]);
    tok.nextInt(&nrot);
    for (int j = 0; j < nrot; ++j) {
      tok.nextDouble(&hold);
      if (j < 9) rot[j] = hold;
    }
    if (nrot == 9)
      deeparc::rotation_matrix_to_angle_axis(rot, aa);
    else if (nrot == 4)
      deeparc::quaternion_to_angle_axis(rot, aa);
    ex->rotation(nrot == 3 ? rot : aa);  // any other count: zero rotation (reference: uninitialised)
    extrinsics.push_back(ex);
  }
  // points: x y z r g b (colour parsed as double, stored as int; :153-164)
  std::vector<Point3d*> points;
  points.reserve(static_cast<size_t>(std::max(n_point, 0)));
  for (int i = 0; i < n_point; ++i) {
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (double& x : v) tok.nextDouble(&x);
    points.push_back(new Point3d(v[0], v[1], v[2], static_cast<int>(v[3]), static_cast<int>(v[4]), static_cast<int>(v[5])));
  }

  intrinsics_ = intrinsics;
  extrinsics_ = extrinsics;
  point3d_ = points;
  if (share_extrinsic_)
    buildHemisphere();
  else
    buildCameras();
  linkBlocks(share_extrinsic_ ? arc_size_ : 0);
  return true;
}

// DeepArcManager.cc:173-196 — resolve ids to pointers; out-of-range ids throw std::out_of_range
void DeepArcManager::linkBlocks(int arc_size) {
  for (ParameterBlock* p : params_) {
    p->intrinsic(intrinsics_.at(p->intrinsic_id()));
    p->point3d(point3d_.at(p->point3d_id()));
    if (arc_size != 0) {
      p->arc(extrinsics_.at(p->pos_arc()));
      p->ring(extrinsics_.at(ringSlot(p->pos_ring(), arc_size)));
      p->share_extrinsic(true);
    } else {
      p->extrinsic(extrinsics_.at(p->extrinsic_id()));
      p->share_extrinsic(false);
    }
  }
}

// DeepArcManager.cc:198-218.  QUIRK: extrinsic ids are overwritten with their arc / ring
// POSITION (slot 0 ends up with id 0 either way); write() emits those ids.
void DeepArcManager::buildHemisphere() {
  for (int a = 0; a < arc_size_; ++a) {
    extrinsics_.at(a)->id(a);
    for (int r = 0; r < ring_size_; ++r) {
      const int slot = ringSlot(r, arc_size_);
      extrinsics_.at(slot)->id(r);
      hemisphere_[a][r] = new Camera(intrinsics_.at(a), extrinsics_.at(a), extrinsics_.at(slot));
    }
  }
}

// DeepArcManager.cc:220-240 — one camera per distinct extrinsic id, first intrinsic seen wins
void DeepArcManager::buildCameras() {
  std::map<int, int> intrinsic_of;
  for (ParameterBlock* p : params_) intrinsic_of.insert(std::make_pair(p->extrinsic_id(), p->intrinsic_id()));
  for (const auto& kv : intrinsic_of) camera_.push_back(new Camera(intrinsics_.at(kv.second), extrinsics_.at(kv.first)));
}

// camera centre -R^T t (DeepArcManager.cc:242-251)
std::vector<double> DeepArcManager::centreOf(Extrinsic* pose) {
  double m[9], c[3];
  deeparc::angle_axis_to_rotation_matrix(pose->rotation(), m);
  deeparc::camera_centre(m, pose->translation(), c);
  return std::vector<double>(c, c + 3);
}

// composed pose x_cam = R_arc (R_ring x + t_ring) + t_arc:
// centre = -R_ring^T t_ring - R_ring^T R_arc^T t_arc (DeepArcManager.cc:253-264)
std::vector<double> DeepArcManager::centreOf(Extrinsic* arc, Extrinsic* ring) {
  // evaluated in the association of the reference's expression:
  //   (-R1^T) t1 - ((R1^T R2^T) t2),  R1 = ring, R2 = arc
  double m1[9], m2[9], c1[3];
  deeparc::angle_axis_to_rotation_matrix(ring->rotation(), m1);
  deeparc::angle_axis_to_rotation_matrix(arc->rotation(), m2);
  deeparc::camera_centre(m1, ring->translation(), c1);
  const double* t2 = arc->translation();
  std::vector<double> c(3);
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    for (int j = 0; j < 3; ++j) {
      // (R1^T R2^T)(i, j) = sum_k R1(k, i) R2(j, k); column-major storage m[col * 3 + row]
      double mij = 0.0;
      for (int k = 0; k < 3; ++k) mij += m1[i * 3 + k] * m2[k * 3 + j];
      acc += mij * t2[j];
    }
    c[i] = c1[i] - acc;
  }
  return c;
}

// which pose(s) define camera (arc, ring): same selection as ParameterBlock::get()
std::vector<double> DeepArcManager::centreOfCamera(int arc, int ring, bool* single_pose) {
  Camera* cam = hemisphere_[arc][ring];
  if (ring == 0) {
    *single_pose = true;
    return centreOf(cam->arc());
  }
  if (arc == 0) {
    *single_pose = true;
    return centreOf(cam->ring());
  }
  *single_pose = false;
  return centreOf(cam->arc(), cam->ring());
}

// DeepArcManager.cc:501-518 (empty for non-shared files: ring_size_ == 0)
std::vector<std::vector<double> > DeepArcManager::getCameraCenter() {
  std::vector<std::vector<double> > centres;
  for (int a = 0; a < arc_size_; ++a)
    for (int r = 0; r < ring_size_; ++r) {
      bool single = false;
      centres.push_back(centreOfCamera(a, r, &single));
    }
  return centres;
}

// ASCII PLY: camera centres (green = single pose, magenta = composed), then the points
// (DeepArcManager.cc:266-328).  Numbers use default ostream formatting (%g, 6 digits).
void DeepArcManager::writePly(std::string filename) {
  const int n_cam = share_extrinsic_ ? arc_size_ * ring_size_ : static_cast<int>(camera_.size());
  std::string out;
  out.reserve(64 * (point3d_.size() + static_cast<size_t>(n_cam)) + 256);
  out += "ply\nformat ascii 1.0\nelement vertex " + std::to_string(point3d_.size() + static_cast<size_t>(n_cam)) +
         "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
         "property uchar blue\nend_header\n";
  auto emit_cam = [&out](const std::vector<double>& c, bool green) {
    for (int i = 0; i < 3; ++i) {
      fmtg(&out, c[i]);
      out += ' ';
    }
    out += green ? "0 255 0\n" : "255 0 255\n";
  };
  if (share_extrinsic_) {
    for (int a = 0; a < arc_size_; ++a)
      for (int r = 0; r < ring_size_; ++r) {
        bool single = false;
        const std::vector<double> c = centreOfCamera(a, r, &single);
        emit_cam(c, single);
      }
  } else {
    for (Camera* cam : camera_) emit_cam(centreOf(cam->extrinsic()), true);
  }
  for (Point3d* p : point3d_) {
    const double* x = p->position();
    for (int j = 0; j < 3; ++j) {
      fmtg(&out, x[j]);
      out += ' ';
    }
    out += std::to_string(p->r()) + ' ' + std::to_string(p->g()) + ' ' + std::to_string(p->b()) + '\n';
  }
  std::ofstream of(filename, std::ios::binary);
  of.write(out.data(), static_cast<std::streamsize>(out.size()));
}

// DeepArcManager.cc:331-424.  The per-observation residuals (:332-352) are evaluated by the
// GPU engine at the parameters currently in the scene graph; the pointer-graph surgery that
// follows is the reference's, step for step.
void DeepArcManager::filterPoint3d(double error_boundary, double* hemisphere_center, double hemisphere_radius) {
  if (!params_.empty()) {
    deeparc::FlatProblem flat;
    deeparc::flatten(*this, /*freeze_camera=*/false, &flat);
    dba_problem view = flat.view();
    dba_handle* h = deeparc::engine();
    deeparc::check(dba_problem_set(h, &view), "dba_problem_set");
    std::vector<double> mse(params_.size());
    deeparc::check(dba_filter_mse(h, mse.data()), "dba_filter_mse");
    for (size_t i = 0; i < params_.size(); ++i)
      if (mse[i] < error_boundary) params_[i]->require_remove(true);  // QUIRK: "<" as written (:348)
  }
  auto drop_flagged_blocks = [this]() {
    params_.erase(std::remove_if(params_.begin(), params_.end(),
                                 [](ParameterBlock* b) {
                                   const bool gone = b->require_remove();
                                   if (gone) delete b;
                                   return gone;
                                 }),
                  params_.end());
  };
  drop_flagged_blocks();
  // points left without observations (:368-378)
  point3d_.erase(std::remove_if(point3d_.begin(), point3d_.end(),
                                [](Point3d* p) {
                                  const bool gone = p->empty();
                                  if (gone) delete p;
                                  return gone;
                                }),
                 point3d_.end());
  // points outside the hemisphere: |x - c|^2 > rho / 2, rho being the squared radius (:380-390)
  for (Point3d* p : point3d_) {
    const double* x = p->position();
    double d2 = 0.0;
    for (int i = 0; i < 3; ++i) d2 += (x[i] - hemisphere_center[i]) * (x[i] - hemisphere_center[i]);
    if (d2 > hemisphere_radius / 2) p->require_remove(true);
  }
  point3d_.erase(std::remove_if(point3d_.begin(), point3d_.end(),
                                [](Point3d* p) {
                                  const bool gone = p->require_remove();
                                  if (gone) {
                                    for (ParameterBlock* b : p->total_link()) {
                                      b->require_remove(true);
                                      b->point3d(NULL);
                                    }
                                    delete p;
                                  }
                                  return gone;
                                }),
                 point3d_.end());
  drop_flagged_blocks();
}

// `.deeparc` v0.01 text, fixed 6 decimals, points re-indexed, rotations always angle-axis
// (DeepArcManager.cc:426-499).
void DeepArcManager::write(std::string filename) {
  for (size_t i = 0; i < point3d_.size(); ++i) point3d_[i]->id(static_cast<int>(i));
  std::string out;
  out.reserve(48 * params_.size() + 80 * point3d_.size() + 4096);
  out += "0.010000\n" + std::to_string(params_.size()) + " " + std::to_string(intrinsics_.size()) + " ";
  if (share_extrinsic_)
    out += std::to_string(arc_size_) + " " + std::to_string(ring_size_) + " ";
  else
    out += std::to_string(camera_.size()) + " 0 ";
  out += std::to_string(point3d_.size()) + "\n";
  for (ParameterBlock* b : params_) {
    out += std::to_string(b->intrinsic()->id()) + " ";
    out += std::to_string(share_extrinsic_ ? b->ring()->id() : b->extrinsic()->id()) + " ";
    out += std::to_string(b->point3d()->id()) + " ";
    fmt6(&out, b->point2d()->x());
    out += ' ';
    fmt6(&out, b->point2d()->y());
    out += '\n';
  }
  for (Intrinsic* in : intrinsics_) {
    fmt6(&out, in->center()[0]);
    out += ' ';
    fmt6(&out, in->center()[1]);
    out += ' ' + std::to_string(in->focal_size());
    for (int j = 0; j < in->focal_size(); ++j) {
      out += ' ';
      fmt6(&out, in->focal()[j]);
    }
    out += ' ' + std::to_string(in->distrotion_size());
    for (int j = 0; j < in->distrotion_size(); ++j) {
      out += ' ';
      fmt6(&out, in->distrotion()[j]);
    }
    out += '\n';
  }
  for (Extrinsic* ex : extrinsics_) {
    for (int j = 0; j < 3; ++j) {
      fmt6(&out, ex->translation()[j]);
      out += ' ';
    }
    out += "3 ";
    for (int j = 0; j < 3; ++j) {
      fmt6(&out, ex->rotation()[j]);
      out += j < 2 ? " " : "\n";
    }
  }
  for (Point3d* p : point3d_) {
    for (int j = 0; j < 3; ++j) {
      fmt6(&out, p->position()[j]);
      out += ' ';
    }
    out += std::to_string(p->r()) + ' ' + std::to_string(p->g()) + ' ' + std::to_string(p->b()) + '\n';
  }
  std::ofstream of(filename, std::ios::binary);
  of.write(out.data(), static_cast<std::streamsize>(out.size()));
}
