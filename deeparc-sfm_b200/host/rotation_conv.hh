// Host rotation conversions used by the loader and the camera-centre / PLY output
// (not on the hot path: O(#extrinsics), once per file).  The reference calls
// ceres::RotationMatrixToAngleAxis / QuaternionToAngleAxis (DeepArcManager.cc:142,144) and
// ceres::AngleAxisToRotationMatrix (Camera/Extrinsic.hh:14); Ceres is not a dependency here,
// so the published formulas of ceres/rotation.h are restated in plain double [Ceres-upstream].
// Conventions: 3x3 matrices are COLUMN-MAJOR (m[col * 3 + row]); quaternions are (w, x, y, z).
#ifndef DEEPARC_B200_ROTATION_CONV_HH_
#define DEEPARC_B200_ROTATION_CONV_HH_

#include <cfloat>
#include <cmath>

namespace deeparc {

inline void quaternion_to_angle_axis(const double q[4], double aa[3]) {
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double k = 2.0;  // zero rotation: first-order limit
  if (s2 > 0.0) {
    const double s = std::sqrt(s2);
    // keep the rotation angle in (-pi, pi]
    const double two_theta = 2.0 * (q[0] < 0.0 ? std::atan2(-s, -q[0]) : std::atan2(s, q[0]));
    k = two_theta / s;
  }
  aa[0] = q[1] * k;
  aa[1] = q[2] * k;
  aa[2] = q[3] * k;
}

inline void rotation_matrix_to_quaternion(const double m[9], double q[4]) {
  auto R = [m](int r, int c) { return m[c * 3 + r]; };
  const double trace = R(0, 0) + R(1, 1) + R(2, 2);
  if (trace >= 0.0) {
    double t = std::sqrt(trace + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R(2, 1) - R(1, 2)) * t;
    q[2] = (R(0, 2) - R(2, 0)) * t;
    q[3] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i + 1] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R(k, j) - R(j, k)) * t;
    q[j + 1] = (R(j, i) + R(i, j)) * t;
    q[k + 1] = (R(k, i) + R(i, k)) * t;
  }
}

inline void rotation_matrix_to_angle_axis(const double m[9], double aa[3]) {
  double q[4];
  rotation_matrix_to_quaternion(m, q);
  quaternion_to_angle_axis(q, aa);
}

inline void angle_axis_to_rotation_matrix(const double aa[3], double m[9]) {
  const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  auto R = [m](int r, int c) -> double& { return m[c * 3 + r]; };
  if (theta2 > DBL_EPSILON) {
    const double theta = std::sqrt(theta2);
    const double wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const double c = std::cos(theta), s = std::sin(theta), omc = 1.0 - c;
    R(0, 0) = c + wx * wx * omc;
    R(1, 0) = wz * s + wx * wy * omc;
    R(2, 0) = -wy * s + wx * wz * omc;
    R(0, 1) = wx * wy * omc - wz * s;
    R(1, 1) = c + wy * wy * omc;
    R(2, 1) = wx * s + wy * wz * omc;
    R(0, 2) = wy * s + wx * wz * omc;
    R(1, 2) = -wx * s + wy * wz * omc;
    R(2, 2) = c + wz * wz * omc;
  } else {  // first-order: I + [aa]x
    R(0, 0) = 1.0;    R(1, 0) = aa[2];  R(2, 0) = -aa[1];
    R(0, 1) = -aa[2]; R(1, 1) = 1.0;    R(2, 1) = aa[0];
    R(0, 2) = aa[1];  R(1, 2) = -aa[0]; R(2, 2) = 1.0;
  }
}

// c = (-R^T) t  (camera centre of x_cam = R x + t); R column-major.  The negation is applied
// to the matrix entries, as the reference's Eigen expression `-R.transpose() * t` does
// (DeepArcManager.cc:248), so a zero translation gives +0, not -0, in the PLY output.
inline void camera_centre(const double m[9], const double t[3], double c[3]) {
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    for (int r = 0; r < 3; ++r) acc += (-m[i * 3 + r]) * t[r];
    c[i] = acc;
  }
}

}  // namespace deeparc

#endif  // DEEPARC_B200_ROTATION_CONV_HH_
