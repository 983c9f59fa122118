// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// Restatement of the four ceres/rotation.h helpers the reference calls
// [Ceres-upstream, version unpinned]:
//   ceres::AngleAxisRotatePoint        <- snavely_reprojection_error.hh:87
//   ceres::RotationMatrixToAngleAxis   <- DeepArcManager.cc:142
//   ceres::QuaternionToAngleAxis       <- DeepArcManager.cc:144
//   ceres::AngleAxisToRotationMatrix   <- Camera/Extrinsic.hh:14
// Matrices are COLUMN-MAJOR 3x3 (upstream's pointer overloads use
// ColumnMajorAdapter3x3); quaternions are (w, x, y, z).
// The formulas and the branch thresholds (theta^2 > DBL_EPSILON for the Rodrigues
// branch, first-order Taylor otherwise) are what parity is sensitive to.
#ifndef ORACLE_CERES_SHIM_ROTATION_H_
#define ORACLE_CERES_SHIM_ROTATION_H_

#include <cmath>
#include <limits>

#include "ceres/jet.h"

namespace ceres {

template <typename T>
inline T DotProduct(const T x[3], const T y[3]) {
  return (x[0] * y[0] + x[1] * y[1] + x[2] * y[2]);
}

template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    // result = pt cos(theta) + (w x pt) sin(theta) + w (w . pt) (1 - cos(theta))
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse,
                    angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    // R ~ I + hat(angle_axis): derivative wrt angle_axis of this branch is -hat(pt).
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1],
                             angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// R is column-major: R[col * 3 + row].
template <typename T>
inline void AngleAxisToRotationMatrix(const T* angle_axis, T* R) {
  static const T kOne = T(1.0);
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T wx = angle_axis[0] / theta;
    const T wy = angle_axis[1] / theta;
    const T wz = angle_axis[2] / theta;
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    R[0 * 3 + 0] = costheta + wx * wx * (kOne - costheta);
    R[0 * 3 + 1] = wz * sintheta + wx * wy * (kOne - costheta);
    R[0 * 3 + 2] = -wy * sintheta + wx * wz * (kOne - costheta);
    R[1 * 3 + 0] = wx * wy * (kOne - costheta) - wz * sintheta;
    R[1 * 3 + 1] = costheta + wy * wy * (kOne - costheta);
    R[1 * 3 + 2] = wx * sintheta + wy * wz * (kOne - costheta);
    R[2 * 3 + 0] = wy * sintheta + wx * wz * (kOne - costheta);
    R[2 * 3 + 1] = -wx * sintheta + wy * wz * (kOne - costheta);
    R[2 * 3 + 2] = costheta + wz * wz * (kOne - costheta);
  } else {
    R[0 * 3 + 0] = kOne;
    R[0 * 3 + 1] = angle_axis[2];
    R[0 * 3 + 2] = -angle_axis[1];
    R[1 * 3 + 0] = -angle_axis[2];
    R[1 * 3 + 1] = kOne;
    R[1 * 3 + 2] = angle_axis[0];
    R[2 * 3 + 0] = angle_axis[1];
    R[2 * 3 + 1] = -angle_axis[0];
    R[2 * 3 + 2] = kOne;
  }
}

template <typename T>
inline void QuaternionToAngleAxis(const T* quaternion, T* angle_axis) {
  const T& q1 = quaternion[1];
  const T& q2 = quaternion[2];
  const T& q3 = quaternion[3];
  const T sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
  if (sin_squared_theta > T(0.0)) {
    const T sin_theta = sqrt(sin_squared_theta);
    const T& cos_theta = quaternion[0];
    // If cos_theta < 0 then theta > pi/2 and the angle 2*theta would exceed pi;
    // use the equivalent rotation with the smaller angle.
    const T two_theta = T(2.0) * ((cos_theta < T(0.0)) ? atan2(-sin_theta, -cos_theta)
                                                       : atan2(sin_theta, cos_theta));
    const T k = two_theta / sin_theta;
    angle_axis[0] = q1 * k;
    angle_axis[1] = q2 * k;
    angle_axis[2] = q3 * k;
  } else {
    const T k(2.0);
    angle_axis[0] = q1 * k;
    angle_axis[1] = q2 * k;
    angle_axis[2] = q3 * k;
  }
}

// R column-major.
template <typename T>
inline void RotationMatrixToQuaternion(const T* R, T* quaternion) {
#define ORACLE_R(r, c) R[(c)*3 + (r)]
  const T trace = ORACLE_R(0, 0) + ORACLE_R(1, 1) + ORACLE_R(2, 2);
  if (trace >= T(0.0)) {
    T t = sqrt(trace + T(1.0));
    quaternion[0] = T(0.5) * t;
    t = T(0.5) / t;
    quaternion[1] = (ORACLE_R(2, 1) - ORACLE_R(1, 2)) * t;
    quaternion[2] = (ORACLE_R(0, 2) - ORACLE_R(2, 0)) * t;
    quaternion[3] = (ORACLE_R(1, 0) - ORACLE_R(0, 1)) * t;
  } else {
    int i = 0;
    if (ORACLE_R(1, 1) > ORACLE_R(0, 0)) i = 1;
    if (ORACLE_R(2, 2) > ORACLE_R(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    T t = sqrt(ORACLE_R(i, i) - ORACLE_R(j, j) - ORACLE_R(k, k) + T(1.0));
    quaternion[i + 1] = T(0.5) * t;
    t = T(0.5) / t;
    quaternion[0] = (ORACLE_R(k, j) - ORACLE_R(j, k)) * t;
    quaternion[j + 1] = (ORACLE_R(j, i) + ORACLE_R(i, j)) * t;
    quaternion[k + 1] = (ORACLE_R(k, i) + ORACLE_R(i, k)) * t;
  }
#undef ORACLE_R
}

template <typename T>
inline void RotationMatrixToAngleAxis(const T* R, T* angle_axis) {
  T quaternion[4];
  RotationMatrixToQuaternion(R, quaternion);
  QuaternionToAngleAxis(quaternion, angle_axis);
}

}  // namespace ceres

#endif  // ORACLE_CERES_SHIM_ROTATION_H_
