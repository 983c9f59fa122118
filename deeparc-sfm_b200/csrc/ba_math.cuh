// Device math of the Snavely reprojection residual and its ANALYTIC Jacobians (fp64).
//
// Replaces, per observation, what the reference obtains by autodiff of
// SnavelyReprojectionError::operator() (reference src/snavely_reprojection_error.hh:93-118)
// through ceres::DynamicAutoDiffCostFunction (:11-14, 121-141):
//   rotatePoint  (:80-91)  -> PoseRow (per-extrinsic precomputation) + transform()
//   projectPoint (:38-78)  -> project()
// Design: everything transcendental (sqrt/sin/cos of the angle-axis vector) is hoisted out
// of the O(n_obs) kernels into a per-extrinsic table ("pose rows", a few thousand entries,
// L1/L2 resident), so the observation kernels are pure multiply-add streams.
//
// Rotation model [Ceres-upstream ceres/rotation.h AngleAxisRotatePoint]:
//   theta^2 >  DBL_EPSILON : R X = X cos + (w x X) sin + w (w.X)(1 - cos),  w = omega/theta
//   theta^2 <= DBL_EPSILON : R X = X + omega x X   (autodiff derivative of THIS branch: -[X]x)
// written here as  R X = X + a (omega x X) + b omega x (omega x X)  with a = sin/theta,
// b = (1-cos)/theta^2 (or a = 1, b = 0 in the small-angle branch, which reproduces the
// branch value AND its derivative).  Then, exactly,
//   d(R X)/d omega_k = a (e_k x X) + b (e_k x (omega x X) + omega x (e_k x X))
//                      + omega_k (a1 (omega x X) + b1 omega x (omega x X)),
//   a1 = (theta cos - sin)/theta^3,  b1 = (theta sin - 2(1-cos))/theta^4   (series near 0).
#pragma once
#include <cfloat>
#include <cmath>

namespace dba {

// One row per extrinsic; 20 doubles = 160 B (5 x 32 B sectors).
struct PoseRow {
  double R[9];   // row-major rotation matrix
  double t[3];
  double w[3];   // angle-axis vector
  double a, b, a1, b1;
  double free_;  // 1.0 if the pose is optimised, 0.0 if constant (gauge / freeze)
};

// One row per intrinsic; 8 doubles.
struct IntrRow {
  double fx, fy, cx, cy, k0, k1;
  double nf;   // 1 or 2 (as double to keep the row homogeneous)
  double nd;   // 0, 1 or 2: distortion coefficients that exist as parameters
};

__host__ __device__ inline void make_pose_row(const double* w, const double* t, double free_flag, PoseRow* out) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double a, b, a1, b1;
  double* R = out->R;
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    const double s = sin(theta);
    const double c = cos(theta);
    const double sh = sin(0.5 * theta);
    const double omc = 2.0 * sh * sh;  // 1 - cos(theta) without cancellation
    const double wx = w[0] / theta, wy = w[1] / theta, wz = w[2] / theta;
    R[0] = c + wx * wx * omc;
    R[1] = wx * wy * omc - wz * s;
    R[2] = wy * s + wx * wz * omc;
    R[3] = wz * s + wx * wy * omc;
    R[4] = c + wy * wy * omc;
    R[5] = -wx * s + wy * wz * omc;
    R[6] = -wy * s + wx * wz * omc;
    R[7] = wx * s + wy * wz * omc;
    R[8] = c + wz * wz * omc;
    a = s / theta;
    b = omc / theta2;
    if (theta2 < 1e-2) {
      const double x = theta2;
      a1 = -1.0 / 3.0 + x * (1.0 / 30.0 + x * (-1.0 / 840.0 + x * (1.0 / 45360.0 - x * (1.0 / 3991680.0))));
      b1 = -1.0 / 12.0 + x * (1.0 / 180.0 + x * (-1.0 / 6720.0 + x * (1.0 / 453600.0 - x * (1.0 / 47900160.0))));
    } else {
      a1 = (theta * c - s) / (theta2 * theta);
      b1 = (theta * s - 2.0 * omc) / (theta2 * theta2);
    }
  } else {
    R[0] = 1.0;   R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2];  R[4] = 1.0;   R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0];  R[8] = 1.0;
    a = 1.0; b = 0.0; a1 = 0.0; b1 = 0.0;
  }
  out->t[0] = t[0]; out->t[1] = t[1]; out->t[2] = t[2];
  out->w[0] = w[0]; out->w[1] = w[1]; out->w[2] = w[2];
  out->a = a; out->b = b; out->a1 = a1; out->b1 = b1;
  out->free_ = free_flag;
}

// out = R X + t
__host__ __device__ __forceinline__ void transform(const PoseRow& P, const double X[3], double out[3]) {
  out[0] = P.R[0] * X[0] + P.R[1] * X[1] + P.R[2] * X[2] + P.t[0];
  out[1] = P.R[3] * X[0] + P.R[4] * X[1] + P.R[5] * X[2] + P.t[1];
  out[2] = P.R[6] * X[0] + P.R[7] * X[1] + P.R[8] * X[2] + P.t[2];
}

// D[r][k] = d (R X)_r / d omega_k
__host__ __device__ __forceinline__ void rotation_derivative(const PoseRow& P, const double X[3], double D[3][3]) {
  const double wx = P.w[0], wy = P.w[1], wz = P.w[2];
  // c1 = w x X, c2 = w x c1
  const double c1x = wy * X[2] - wz * X[1], c1y = wz * X[0] - wx * X[2], c1z = wx * X[1] - wy * X[0];
  const double c2x = wy * c1z - wz * c1y, c2y = wz * c1x - wx * c1z, c2z = wx * c1y - wy * c1x;
  const double mx = P.a1 * c1x + P.b1 * c2x, my = P.a1 * c1y + P.b1 * c2y, mz = P.a1 * c1z + P.b1 * c2z;
  const double a = P.a, b = P.b;
  // k = 0: e0 x X = (0,-X2,X1); e0 x c1 = (0,-c1z,c1y); w x (e0 x X) = (wy X1 + wz X2, -wx X1, -wx X2)
  D[0][0] = b * (wy * X[1] + wz * X[2]) + wx * mx;
  D[1][0] = a * (-X[2]) + b * (-c1z - wx * X[1]) + wx * my;
  D[2][0] = a * (X[1]) + b * (c1y - wx * X[2]) + wx * mz;
  // k = 1: e1 x X = (X2,0,-X0); e1 x c1 = (c1z,0,-c1x); w x (e1 x X) = (-wy X0, wx X0 + wz X2, -wy X2)
  D[0][1] = a * (X[2]) + b * (c1z - wy * X[0]) + wy * mx;
  D[1][1] = b * (wx * X[0] + wz * X[2]) + wy * my;
  D[2][1] = a * (-X[0]) + b * (-c1x - wy * X[2]) + wy * mz;
  // k = 2: e2 x X = (-X1,X0,0); e2 x c1 = (-c1y,c1x,0); w x (e2 x X) = (-wz X0, -wz X1, wx X0 + wy X1)
  D[0][2] = a * (-X[1]) + b * (-c1y - wz * X[0]) + wz * mx;
  D[1][2] = a * (X[0]) + b * (c1x - wz * X[1]) + wz * my;
  D[2][2] = b * (wx * X[0] + wy * X[1]) + wz * mz;
}

struct Projection {
  double r0, r1;      // residual
  double G[2][3];     // d r / d p (camera-frame point)
  double df[2];       // d r / d f   (nf = 1: both rows; nf = 2: (d r0/d fx, d r1/d fy))
  double dk0[2], dk1[2];
};

// projectPoint (snavely_reprojection_error.hh:38-78): no sign flip, fy = nf==2 ? f[1] : f[0],
// distortion 1 + r2 (k0 + k1 r2) with unused coefficients stored as exact zeros.
template <bool WITH_JAC>
__host__ __device__ __forceinline__ void project(const IntrRow& I, const double p[3], double ox, double oy, Projection& out) {
  const double u = p[0] / p[2];
  const double v = p[1] / p[2];
  const double rr = u * u + v * v;
  const double d = 1.0 + rr * (I.k0 + I.k1 * rr);
  const double fxd = I.fx * d, fyd = I.fy * d;
  out.r0 = fxd * u + I.cx - ox;
  out.r1 = fyd * v + I.cy - oy;
  if (WITH_JAC) {
    const double iz = 1.0 / p[2];
    const double kap = I.k0 + 2.0 * I.k1 * rr;  // d(d)/d(rr)
    const double r0u = I.fx * (d + 2.0 * u * u * kap), r0v = I.fx * (2.0 * u * v * kap);
    const double r1u = I.fy * (2.0 * u * v * kap), r1v = I.fy * (d + 2.0 * v * v * kap);
    out.G[0][0] = r0u * iz;
    out.G[0][1] = r0v * iz;
    out.G[0][2] = -(r0u * u + r0v * v) * iz;
    out.G[1][0] = r1u * iz;
    out.G[1][1] = r1v * iz;
    out.G[1][2] = -(r1u * u + r1v * v) * iz;
    // nf = 1: d r / d f.  nf = 2: the diagonal of the 2x2 focal Jacobian (d r0/d fx, d r1/d fy); its
    // off-diagonal entries are identically zero.  (Free intrinsics need nf = 1, so the solver only ever
    // sees the first meaning; dba_eval reports both.)
    out.df[0] = d * u;
    out.df[1] = d * v;
    // coefficients that are not parameters of this intrinsic (nd < 1, nd < 2) have no column
    const double m0 = I.nd >= 1.0 ? rr : 0.0, m1 = I.nd >= 2.0 ? rr * rr : 0.0;
    out.dk0[0] = I.fx * u * m0;
    out.dk0[1] = I.fy * v * m0;
    out.dk1[0] = I.fx * u * m1;
    out.dk1[1] = I.fy * v * m1;
  }
}

// C(2x3) = A(2x3) * B(3x3), B given as B[r][c]
__host__ __device__ __forceinline__ void mul23_33(const double A[2][3], const double B[3][3], double C[2][3]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}
// C(2x3) = A(2x3) * R, R row-major 9
__host__ __device__ __forceinline__ void mul23_R(const double A[2][3], const double* R, double C[2][3]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * R[j] + A[i][1] * R[3 + j] + A[i][2] * R[6 + j];
}

// Everything the Jacobian kernel stores for one observation, UNSCALED:
//   r, Jp = d r/d X, JA = d r/d (w_a, t_a [, f, k0, k1]), JB = d r/d (w_b, t_b) (two-pose only).
struct ObsJacobian {
  double r0, r1;
  double Jp[2][3];
  double JwA[2][3], JtA[2][3];
  double JwB[2][3], JtB[2][3];
  double df[2], dk0[2], dk1[2];
};

// p = R_a (R_b X + t_b) + t_a (B != nullptr) or p = R_a X + t_a; chain rule through project():
//   d r/d t_a = G;  d r/d w_a = G D(w_a; mid);  d r/d mid = G R_a =: GA
//   d r/d t_b = GA; d r/d w_b = GA D(w_b; X);   d r/d X = GA R_b   (or GA when there is no B)
__host__ __device__ __forceinline__ void observation_jacobian(const PoseRow& A, const PoseRow* B, const IntrRow& I,
                                                              const double X[3], double ox, double oy, bool want_cam,
                                                              ObsJacobian& out) {
  double mid[3], cam[3];
  if (B) {
    transform(*B, X, mid);
  } else {
    mid[0] = X[0];
    mid[1] = X[1];
    mid[2] = X[2];
  }
  transform(A, mid, cam);
  Projection pr;
  project<true>(I, cam, ox, oy, pr);
  out.r0 = pr.r0;
  out.r1 = pr.r1;
  double GA[2][3];
  mul23_R(pr.G, A.R, GA);
  if (B) {
    mul23_R(GA, B->R, out.Jp);
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) out.Jp[i][k] = GA[i][k];
  }
  if (!want_cam) return;
  double D[3][3];
  rotation_derivative(A, mid, D);
  mul23_33(pr.G, D, out.JwA);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) out.JtA[i][k] = pr.G[i][k];
  out.df[0] = pr.df[0];
  out.df[1] = pr.df[1];
  out.dk0[0] = pr.dk0[0];
  out.dk0[1] = pr.dk0[1];
  out.dk1[0] = pr.dk1[0];
  out.dk1[1] = pr.dk1[1];
  if (B) {
    rotation_derivative(*B, X, D);
    mul23_33(GA, D, out.JwB);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) out.JtB[i][k] = GA[i][k];
  }
}

// inverse of a symmetric positive definite 3x3 given by its 6 unique entries
// (c00 c01 c02 c11 c12 c22); returns false if not positive definite.
__host__ __device__ inline bool inv_sym3(const double c[6], double inv[6]) {
  const double m00 = c[3] * c[5] - c[4] * c[4];
  const double m01 = c[2] * c[4] - c[1] * c[5];
  const double m02 = c[1] * c[4] - c[2] * c[3];
  const double det = c[0] * m00 + c[1] * m01 + c[2] * m02;
  if (!(det > 0.0) || !(c[0] > 0.0) || !(c[0] * c[3] - c[1] * c[1] > 0.0)) return false;
  const double id = 1.0 / det;
  inv[0] = m00 * id;
  inv[1] = m01 * id;
  inv[2] = m02 * id;
  inv[3] = (c[0] * c[5] - c[2] * c[2]) * id;
  inv[4] = (c[1] * c[2] - c[0] * c[4]) * id;
  inv[5] = (c[0] * c[3] - c[1] * c[1]) * id;
  return true;
}

}  // namespace dba
