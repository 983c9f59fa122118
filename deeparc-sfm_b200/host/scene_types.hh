// Host-side data model of a DeepArc scene: the AoS pointer graph that the reference keeps in
// src/Point/Point2d.hh, src/Point/Point3d.hh, src/Camera/{Intrinsic,Extrinsic,Camera}.hh and
// src/ParameterBlock.hh, restated in one header with the same class and accessor names so that
// code written against the reference (its sfm.cc, its DeepArcManager) reads the same here.
// No Ceres / Eigen: the rotation helpers live in rotation_conv.hh.
//
// Contracts that matter to the bundle-adjustment boundary:
//  * raw `double*` accessors — the solver optimises IN PLACE through them (sfm.cc:47-48);
//  * ParameterBlock::get() fixes the parameter-block order the residual indexes
//    (ParameterBlock.hh:68-94) and which extrinsics an observation uses in shared mode;
//  * Intrinsic::center(int,int) truncates the principal point (Intrinsic.hh:24-27);
//  * index aliasing intrinsic_id = pos_arc, extrinsic_id = pos_ring (ParameterBlock.hh:52-55).
#ifndef DEEPARC_B200_SCENE_TYPES_HH_
#define DEEPARC_B200_SCENE_TYPES_HH_

#include <algorithm>
#include <array>
#include <functional>
#include <set>
#include <vector>

class ParameterBlock;

// observed pixel (Point2d.hh)
class Point2d {
 public:
  Point2d(double x, double y) : xy_{x, y} {}
  double x() const { return xy_[0]; }
  double y() const { return xy_[1]; }

 private:
  std::array<double, 2> xy_;
};

// 3-D point with colour, removal flag and back-links to its observations (Point3d.hh)
class Point3d {
 public:
  Point3d(double x, double y, double z, int r = 255, int g = 255, int b = 255)
      : remove_(false), rgb_{r, g, b}, id_(0), position_{x, y, z} {}
  int r() const { return rgb_[0]; }
  int g() const { return rgb_[1]; }
  int b() const { return rgb_[2]; }
  int id() const { return id_; }
  void id(int point3d_id) { id_ = point3d_id; }
  double* position() { return position_.data(); }
  void require_remove(bool remove) { remove_ = remove; }
  bool require_remove() const { return remove_; }
  // Observations of this point.  The reference keeps a std::set<ParameterBlock*>; here it is a sorted
  // vector (same iteration order, one allocation per point instead of one per observation) and
  // total_link() materialises the set the reference's interface returns.
  void link(ParameterBlock* block) {
    auto it = std::lower_bound(blocks_.begin(), blocks_.end(), block, std::less<ParameterBlock*>());
    if (it == blocks_.end() || *it != block) blocks_.insert(it, block);
  }
  void unlink(ParameterBlock* block) {
    auto it = std::lower_bound(blocks_.begin(), blocks_.end(), block, std::less<ParameterBlock*>());
    if (it != blocks_.end() && *it == block) blocks_.erase(it);
  }
  std::set<ParameterBlock*> total_link() const { return std::set<ParameterBlock*>(blocks_.begin(), blocks_.end()); }
  bool empty() const { return blocks_.empty(); }
  // bulk construction (the parallel loader): slots filled in any order, then sorted once
  void link_slots(size_t n) { blocks_.assign(n, nullptr); }
  void link_slot(size_t i, ParameterBlock* block) { blocks_[i] = block; }
  void link_finish() {
    std::sort(blocks_.begin(), blocks_.end(), std::less<ParameterBlock*>());
    blocks_.erase(std::unique(blocks_.begin(), blocks_.end()), blocks_.end());
  }

 private:
  bool remove_;
  std::array<int, 3> rgb_;
  int id_;
  std::array<double, 3> position_;
  std::vector<ParameterBlock*> blocks_;
};

// principal point, 1-2 focal lengths, 0-2 radial distortion coefficients (Intrinsic.hh)
class Intrinsic {
 public:
  Intrinsic() : focal_{0, 0}, center_{0, 0}, distrotion_{0, 0}, focal_size_(0), distrotion_size_(0), id_(0) {}
  double* focal() { return focal_.data(); }
  double* center() { return center_.data(); }
  double* distrotion() { return distrotion_.data(); }  // (sic) reference spelling
  int focal_size() const { return focal_size_; }
  int distrotion_size() const { return distrotion_size_; }
  int id() const { return id_; }
  void id(int v) { id_ = v; }
  void focal(int size, const double* f) {
    focal_size_ = size;
    for (int i = 0; i < size && i < 2; ++i) focal_[i] = f[i];
  }
  void distrotion(int size, const double* k) {
    distrotion_size_ = size;
    for (int i = 0; i < size && i < 2; ++i) distrotion_[i] = k[i];
  }
  // The reference declares center(int, int) and calls it with doubles, so the principal
  // point is truncated toward zero on load (Intrinsic.hh:24-27, DeepArcManager.cc:103-104).
  void center(int cx, int cy) {
    center_[0] = cx;
    center_[1] = cy;
  }

 private:
  std::array<double, 2> focal_, center_, distrotion_;
  int focal_size_, distrotion_size_, id_;
};

// angle-axis rotation + translation (Extrinsic.hh)
class Extrinsic {
 public:
  Extrinsic() : rotation_{0, 0, 0}, translation_{0, 0, 0}, id_(0) {}
  double* rotation() { return rotation_.data(); }
  double* translation() { return translation_.data(); }
  int id() const { return id_; }
  void id(int v) { id_ = v; }
  void rotation(const double* r) {
    for (int i = 0; i < 3; ++i) rotation_[i] = r[i];
  }
  void translation(double x, double y, double z) { translation_ = {x, y, z}; }

 private:
  std::array<double, 3> rotation_, translation_;
  int id_;
};

// intrinsic + one extrinsic, or intrinsic + (arc, ring) extrinsics (Camera.hh)
class Camera {
 public:
  Camera(Intrinsic* intrinsic, Extrinsic* extrinsic)
      : intrinsic_(intrinsic), extrinsic_(extrinsic), arc_(nullptr), ring_(nullptr) {}
  Camera(Intrinsic* intrinsic, Extrinsic* on_arc, Extrinsic* on_ring)
      : intrinsic_(intrinsic), extrinsic_(nullptr), arc_(on_arc), ring_(on_ring) {}
  Intrinsic* intrinsic() { return intrinsic_; }
  Extrinsic* extrinsic() { return extrinsic_; }
  Extrinsic* arc() { return arc_; }
  Extrinsic* ring() { return ring_; }

 private:
  Intrinsic* intrinsic_;
  Extrinsic *extrinsic_, *arc_, *ring_;
};

// One observation: (arc position, ring position, point id, pixel) (ParameterBlock.hh)
class ParameterBlock {
 public:
  ParameterBlock(int position_arc, int position_ring, int point3d_id, Point2d* point2d)
      : share_extrinsic_(false), require_remove_(false), intrinsic_(nullptr), extrinsic_(nullptr), arc_(nullptr),
        ring_(nullptr), point3d_(nullptr), point2d_(point2d), pos_arc_(position_arc), pos_ring_(position_ring),
        intrinsic_id_(position_arc), point3d_id_(point3d_id), extrinsic_id_(position_ring) {}
  ~ParameterBlock() {
    if (point3d_) point3d_->unlink(this);
    delete point2d_;
  }
  ParameterBlock(const ParameterBlock&) = delete;
  ParameterBlock& operator=(const ParameterBlock&) = delete;

  Extrinsic* arc() { return arc_; }
  Extrinsic* ring() { return ring_; }
  Extrinsic* extrinsic() { return extrinsic_; }
  Point2d* point2d() { return point2d_; }
  Point3d* point3d() { return point3d_; }
  Intrinsic* intrinsic() { return intrinsic_; }
  int intrinsic_id() const { return intrinsic_id_; }
  int extrinsic_id() const { return extrinsic_id_; }
  int point3d_id() const { return point3d_id_; }
  int pos_arc() const { return pos_arc_; }
  int pos_ring() const { return pos_ring_; }
  bool share_extrinsic() const { return share_extrinsic_; }
  bool require_remove() const { return require_remove_; }
  // two poses enter the residual only off the arc-0 column and ring-0 row (ParameterBlock.hh:24-28)
  bool compose_extrinsic() const { return share_extrinsic_ && pos_arc_ != 0 && pos_ring_ != 0; }

  void arc(Extrinsic* e) { arc_ = e; }
  void ring(Extrinsic* e) { ring_ = e; }
  void extrinsic(Extrinsic* e) { extrinsic_ = e; }
  void point2d(Point2d* p) { point2d_ = p; }
  void intrinsic(Intrinsic* i) { intrinsic_ = i; }
  void point3d(Point3d* p) {
    if (point3d_) point3d_->unlink(this);
    point3d_ = p;
    if (point3d_) point3d_->link(this);
  }
  // sets the pointer only; the caller links in bulk (Point3d::link_slots / link_slot / link_finish)
  void point3d_unlinked(Point3d* p) { point3d_ = p; }
  void share_extrinsic(bool v) { share_extrinsic_ = v; }
  void require_remove(bool v) { require_remove_ = v; }

  // Parameter-block list in the order the residual indexes it:
  //   [0] point [1] principal point [2] focal [3] distortion [4] rot [5] trans ([6] rot_b [7] trans_b)
  // Shared mode: ring 0 -> the arc pose alone; arc 0 (ring != 0) -> the ring pose alone;
  // otherwise arc pose then ring pose (ParameterBlock.hh:68-94).
  std::vector<double*> get() {
    std::vector<double*> blocks{point3d_->position(), intrinsic_->center(), intrinsic_->focal(),
                                intrinsic_->distrotion()};
    auto push_pose = [&blocks](Extrinsic* e) {
      blocks.push_back(e->rotation());
      blocks.push_back(e->translation());
    };
    if (!share_extrinsic_) {
      push_pose(extrinsic_);
    } else if (pos_ring_ == 0) {
      push_pose(arc_);
    } else if (pos_arc_ == 0) {
      push_pose(ring_);
    } else {
      push_pose(arc_);
      push_pose(ring_);
    }
    return blocks;
  }

 private:
  bool share_extrinsic_, require_remove_;
  Intrinsic* intrinsic_;
  Extrinsic *extrinsic_, *arc_, *ring_;
  Point3d* point3d_;
  Point2d* point2d_;
  int pos_arc_, pos_ring_, intrinsic_id_, point3d_id_, extrinsic_id_;
};

#endif  // DEEPARC_B200_SCENE_TYPES_HH_
