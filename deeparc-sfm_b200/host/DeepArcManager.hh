// Scene I/O + scene-graph manager with the public surface of the reference's DeepArcManager
// (reference src/DeepArcManager.hh:12-22): read / write / writePly / parameters / point3ds /
// filterPoint3d / getCameraCenter / isShareExtrinsic, same names, argument meaning and error
// behaviour (read throws a `const char*` when the file cannot be opened, bad ids raise
// std::out_of_range).  Implementation is new: a single-pass tokenizer instead of iostream
// extraction, no Ceres/Eigen, and the per-observation residuals of filterPoint3d come from the
// GPU engine (dba_filter_mse) — there is no CPU evaluation of the residual in this library.
#ifndef DEEPARC_B200_DEEPARC_MANAGER_HH_
#define DEEPARC_B200_DEEPARC_MANAGER_HH_

#include <map>
#include <string>
#include <vector>

#include "scene_types.hh"

class DeepArcManager {
 public:
  DeepArcManager();
  ~DeepArcManager();
  DeepArcManager(const DeepArcManager&) = delete;
  DeepArcManager& operator=(const DeepArcManager&) = delete;

  // ---- reference surface ------------------------------------------------------------
  bool isShareExtrinsic();
  bool read(std::string filename);
  void writePly(std::string filename);
  std::vector<ParameterBlock*>* parameters();
  std::vector<Point3d*>* point3ds();
  void filterPoint3d(double error_boundary, double* hemishpere_center, double hemisphere_radius);
  void write(std::string filename);
  std::vector<std::vector<double> > getCameraCenter();

  // ---- binary side format (lossless, memcpy-speed; read() recognises it by its magic) ----
  void writeBinary(std::string filename);

  // ---- additions used by the GPU boundary (gather / scatter) ---------------------------
  std::vector<Intrinsic*>* intrinsics() { return &intrinsics_; }
  std::vector<Extrinsic*>* extrinsics() { return &extrinsics_; }
  int arc_size() const { return arc_size_; }
  int ring_size() const { return share_extrinsic_ ? ring_size_ : 0; }

 private:
  int arc_size_, ring_size_;
  bool share_extrinsic_;
  std::map<int, std::map<int, Camera*> > hemisphere_;
  std::vector<Intrinsic*> intrinsics_;
  std::vector<Extrinsic*> extrinsics_;
  std::vector<Camera*> camera_;
  std::vector<ParameterBlock*> params_;
  std::vector<Point3d*> point3d_;

  static int ringSlot(int ring_position, int arc_size);
  bool readText(const char* data, size_t size, bool parallel);  // false: the strict parallel parser gave up
  bool readBinary(const char* data, size_t size);
  void clearScene();
  void linkBlocks(int arc_size);
  void buildHemisphere();
  void buildCameras();
  std::vector<double> centreOf(Extrinsic* pose);
  std::vector<double> centreOf(Extrinsic* arc, Extrinsic* ring);
  std::vector<double> centreOfCamera(int arc, int ring, bool* single_pose);
};

// the writer's "%.6f" formatter (exposed for the byte-parity test)
std::string deeparc_format_fixed6(double v);

#endif  // DEEPARC_B200_DEEPARC_MANAGER_HH_
