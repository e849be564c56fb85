// ceres/rotation.h — the rotation helpers the reference's host code uses, with the Ceres 1.7.0 conventions
// (SURVEY.md App. A1): gc.cpp:31,45 call AngleAxisToRotationMatrix / RotationMatrixToAngleAxis on COLUMN-MAJOR
// 3x3 arrays; po_problem.h:38,47-51,59 use the quaternion / rotate-point functions inside its cost functor.
// Host-side only: pose packing and unpacking around the solve.  The device kernels carry their own versions.
#ifndef SLSLAM_B200_CERES_ROTATION_SHIM_H_
#define SLSLAM_B200_CERES_ROTATION_SHIM_H_

#include <cmath>
#include <limits>

namespace ceres {

// R = exp([w]x), column-major: R[i + 3 j] = R_ij.  First-order form I + [w]x when |w|^2 is exactly 0 or tiny.
template <typename T>
inline void AngleAxisToRotationMatrix(const T* w, T* R) {
  const T th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > T(std::numeric_limits<double>::epsilon())) {
    const T th = std::sqrt(th2);
    const T x = w[0] / th, y = w[1] / th, z = w[2] / th;
    const T c = std::cos(th), s = std::sin(th), v = T(1.0) - c;
    R[0] = c + x * x * v;      R[3] = x * y * v - z * s;  R[6] = x * z * v + y * s;
    R[1] = y * x * v + z * s;  R[4] = c + y * y * v;      R[7] = y * z * v - x * s;
    R[2] = z * x * v - y * s;  R[5] = z * y * v + x * s;  R[8] = c + z * z * v;
  } else {
    R[0] = T(1.0); R[3] = -w[2];  R[6] = w[1];
    R[1] = w[2];   R[4] = T(1.0); R[7] = -w[0];
    R[2] = -w[1];  R[5] = w[0];   R[8] = T(1.0);
  }
}

// log map of a column-major rotation matrix; |angle| <= pi.
template <typename T>
inline void RotationMatrixToAngleAxis(const T* R, T* w) {
  // twice the axis times sin(theta)
  w[0] = R[5] - R[7];
  w[1] = R[6] - R[2];
  w[2] = R[1] - R[3];
  T c = (R[0] + R[4] + R[8] - T(1.0)) * T(0.5);
  c = c < T(-1.0) ? T(-1.0) : (c > T(1.0) ? T(1.0) : c);
  T s = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) * T(0.5);
  s = s > T(1.0) ? T(1.0) : s;
  const T th = std::atan2(s, c);
  const T kTiny = T(1e-12);
  if (s > kTiny) {
    const T k = th / (T(2.0) * s);
    w[0] *= k; w[1] *= k; w[2] *= k;
    return;
  }
  if (c > T(0.0)) {   // near the identity: sin(theta) ~ theta
    w[0] *= T(0.5); w[1] *= T(0.5); w[2] *= T(0.5);
    return;
  }
  // near pi: the skew part vanishes; recover the axis from the diagonal, signs from the skew part
  const T inv = T(1.0) / (T(1.0) - c);
  for (int i = 0; i < 3; ++i) {
    T a = (R[4 * i] - c) * inv;
    a = a < T(0.0) ? T(0.0) : a;
    T v = th * std::sqrt(a);
    if (w[i] < T(0.0)) v = -v;
    w[i] = v;
  }
}

template <typename T>
inline void AngleAxisRotatePoint(const T* w, const T* p, T* out) {
  const T th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > T(0.0)) {
    const T th = std::sqrt(th2);
    const T k[3] = {w[0] / th, w[1] / th, w[2] / th};
    const T c = std::cos(th), s = std::sin(th);
    const T x[3] = {k[1] * p[2] - k[2] * p[1], k[2] * p[0] - k[0] * p[2], k[0] * p[1] - k[1] * p[0]};
    const T kp = (k[0] * p[0] + k[1] * p[1] + k[2] * p[2]) * (T(1.0) - c);
    const T r0 = p[0] * c + x[0] * s + k[0] * kp, r1 = p[1] * c + x[1] * s + k[1] * kp, r2 = p[2] * c + x[2] * s + k[2] * kp;
    out[0] = r0; out[1] = r1; out[2] = r2;
  } else {
    const T r0 = p[0] + (w[1] * p[2] - w[2] * p[1]), r1 = p[1] + (w[2] * p[0] - w[0] * p[2]), r2 = p[2] + (w[0] * p[1] - w[1] * p[0]);
    out[0] = r0; out[1] = r1; out[2] = r2;
  }
}

// quaternions are (scalar, vector)
template <typename T>
inline void AngleAxisToQuaternion(const T* w, T* q) {
  const T th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > T(0.0)) {
    const T th = std::sqrt(th2), h = th * T(0.5), k = std::sin(h) / th;
    q[0] = std::cos(h); q[1] = w[0] * k; q[2] = w[1] * k; q[3] = w[2] * k;
  } else {
    q[0] = T(1.0); q[1] = w[0] * T(0.5); q[2] = w[1] * T(0.5); q[3] = w[2] * T(0.5);
  }
}

template <typename T>
inline void QuaternionToAngleAxis(const T* q, T* w) {
  const T s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > T(0.0)) {
    const T s = std::sqrt(s2);
    const T two_th = T(2.0) * (q[0] < T(0.0) ? std::atan2(-s, -q[0]) : std::atan2(s, q[0]));
    const T k = two_th / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    w[0] = q[1] * T(2.0); w[1] = q[2] * T(2.0); w[2] = q[3] * T(2.0);
  }
}

template <typename T>
inline void QuaternionProduct(const T* z, const T* w, T* zw) {
  const T a = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  const T b = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  const T c = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  const T d = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
  zw[0] = a; zw[1] = b; zw[2] = c; zw[3] = d;
}

}  // namespace ceres

#endif  // SLSLAM_B200_CERES_ROTATION_SHIM_H_
