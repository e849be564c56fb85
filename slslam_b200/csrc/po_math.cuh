// Device math for the pose-graph path: the relative-pose residual of PoseConstraintError::operator()
// (reference src/po_problem.h:73-105, built from gc_T_inv / gc_w_20 / gc_T_20 at :27-64) and its Jacobians with
// respect to the two 6-dof poses.  The reference obtains the Jacobians from Ceres dual numbers
// (AutoDiffCostFunction<PoseConstraintError,6,6,6>, src/po_problem.cpp:45-52); the device does the same with a
// small forward-mode dual type so that the branch structure of the rotation helpers (theta^2 > 0 tests on the
// scalar part, the c < 0 branch of the quaternion logarithm) is followed exactly.  Edges are few
// (hundreds), so this kernel is not the cost centre of the PO solve; the dense factorisation is.
#pragma once
#include <cuda_runtime.h>

namespace slslam {

// value + N partials; ND = 0 degenerates to a plain double (used by the residual-only pass)
template <int ND>
struct Dual {
  double a;
  double v[ND > 0 ? ND : 1];
};

template <int ND> __device__ __forceinline__ Dual<ND> dconst(double s) {
  Dual<ND> r; r.a = s;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = 0.0;
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> dvar(double s, int k) {
  Dual<ND> r = dconst<ND>(s);
#pragma unroll
  for (int i = 0; i < ND; ++i) if (i == k) r.v[i] = 1.0;
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> operator+(const Dual<ND>& f, const Dual<ND>& g) {
  Dual<ND> r; r.a = f.a + g.a;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = f.v[i] + g.v[i];
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& f, const Dual<ND>& g) {
  Dual<ND> r; r.a = f.a - g.a;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = f.v[i] - g.v[i];
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& f) {
  Dual<ND> r; r.a = -f.a;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = -f.v[i];
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& f, const Dual<ND>& g) {
  Dual<ND> r; r.a = f.a * g.a;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& f, double s) {
  Dual<ND> r; r.a = f.a * s;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = f.v[i] * s;
  return r;
}
// h = f / g, dh = (df - h dg) / g with one reciprocal (the form ceres::Jet uses)
template <int ND> __device__ __forceinline__ Dual<ND> operator/(const Dual<ND>& f, const Dual<ND>& g) {
  Dual<ND> r; const double gi = 1.0 / g.a; r.a = f.a * gi;
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = (f.v[i] - r.a * g.v[i]) * gi;
  return r;
}
template <int ND> __device__ __forceinline__ Dual<ND> dsqrt(const Dual<ND>& f) {
  Dual<ND> r; r.a = sqrt(f.a); const double t = 1.0 / (2.0 * r.a);
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = t * f.v[i];
  return r;
}
template <int ND> __device__ __forceinline__ void dsincos(const Dual<ND>& f, Dual<ND>& s, Dual<ND>& c) {
  double sv, cv; sincos(f.a, &sv, &cv);
  s.a = sv; c.a = cv;
#pragma unroll
  for (int i = 0; i < ND; ++i) { s.v[i] = cv * f.v[i]; c.v[i] = -sv * f.v[i]; }
}
// atan2(g, f): d = (f dg - g df) / (f^2 + g^2)
template <int ND> __device__ __forceinline__ Dual<ND> datan2(const Dual<ND>& g, const Dual<ND>& f) {
  Dual<ND> r; r.a = atan2(g.a, f.a); const double t = 1.0 / (f.a * f.a + g.a * g.a);
#pragma unroll
  for (int i = 0; i < ND; ++i) r.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]);
  return r;
}

// ---- ceres/rotation.h semantics, written for the dual type ----
template <int ND>
__device__ void rotate_point(const Dual<ND>* w, const Dual<ND>* p, Dual<ND>* out) {
  const Dual<ND> th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2.a > 0.0) {
    const Dual<ND> th = dsqrt(th2);
    Dual<ND> s, c;
    dsincos(th, s, c);
    const Dual<ND> k0 = w[0] / th, k1 = w[1] / th, k2 = w[2] / th;
    const Dual<ND> x0 = k1 * p[2] - k2 * p[1], x1 = k2 * p[0] - k0 * p[2], x2 = k0 * p[1] - k1 * p[0];
    const Dual<ND> kp = k0 * p[0] + k1 * p[1] + k2 * p[2];
    const Dual<ND> omc = dconst<ND>(1.0) - c;
    out[0] = p[0] * c + x0 * s + k0 * omc * kp;
    out[1] = p[1] * c + x1 * s + k1 * omc * kp;
    out[2] = p[2] * c + x2 * s + k2 * omc * kp;
  } else {
    // first-order branch p + w x p (an exactly-zero rotation: the anchor pose of the graph)
    out[0] = p[0] + (w[1] * p[2] - w[2] * p[1]);
    out[1] = p[1] + (w[2] * p[0] - w[0] * p[2]);
    out[2] = p[2] + (w[0] * p[1] - w[1] * p[0]);
  }
}

template <int ND>
__device__ void aa_to_quat(const Dual<ND>* w, Dual<ND>* q) {
  const Dual<ND> th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2.a > 0.0) {
    const Dual<ND> th = dsqrt(th2);
    Dual<ND> s, c;
    dsincos(th * 0.5, s, c);
    const Dual<ND> k = s / th;
    q[0] = c; q[1] = w[0] * k; q[2] = w[1] * k; q[3] = w[2] * k;
  } else {
    q[0] = dconst<ND>(1.0); q[1] = w[0] * 0.5; q[2] = w[1] * 0.5; q[3] = w[2] * 0.5;
  }
}

template <int ND>
__device__ void quat_to_aa(const Dual<ND>* q, Dual<ND>* w) {
  const Dual<ND> s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2.a > 0.0) {
    const Dual<ND> s = dsqrt(s2);
    // |angle| <= pi: use atan2(-s, -c) when the scalar part is negative
    const Dual<ND> half = (q[0].a < 0.0) ? datan2(-s, -q[0]) : datan2(s, q[0]);
    const Dual<ND> k = (half * 2.0) / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    w[0] = q[1] * 2.0; w[1] = q[2] * 2.0; w[2] = q[3] * 2.0;
  }
}

// T20 = T21 o T10 on (angle-axis, translation) 6-vectors: rotations through quaternions, t20 = R21 t10 + t21
// (reference src/po_problem.h:42-64).
template <int ND>
__device__ void pose_compose(const Dual<ND>* T21, const Dual<ND>* T10, Dual<ND>* T20) {
  Dual<ND> qa[4], qb[4], qc[4];
  aa_to_quat(T21, qa);
  aa_to_quat(T10, qb);
  qc[0] = qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2] - qa[3] * qb[3];
  qc[1] = qa[0] * qb[1] + qa[1] * qb[0] + qa[2] * qb[3] - qa[3] * qb[2];
  qc[2] = qa[0] * qb[2] - qa[1] * qb[3] + qa[2] * qb[0] + qa[3] * qb[1];
  qc[3] = qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1] + qa[3] * qb[0];
  quat_to_aa(qc, T20);
  rotate_point(T21, T10 + 3, T20 + 3);
  T20[3] = T20[3] + T21[3]; T20[4] = T20[4] + T21[4]; T20[5] = T20[5] + T21[5];
}

// inverse pose: (-w, Rot(-w)(-t))  (reference src/po_problem.h:27-39)
template <int ND>
__device__ void pose_inverse(const Dual<ND>* P, Dual<ND>* Pi) {
  Pi[0] = -P[0]; Pi[1] = -P[1]; Pi[2] = -P[2];
  Dual<ND> nt[3] = {-P[3], -P[4], -P[5]};
  rotate_point(Pi, nt, Pi + 3);
}

// residual = T2^-1 o (C o T1) as (angle-axis, translation); identity information matrix
// (reference src/po_problem.h:73-105).
template <int ND>
__device__ void pose_constraint_residual(const Dual<ND>* T1, const Dual<ND>* T2, const double* cons, Dual<ND>* res) {
  Dual<ND> C[6], Tc[6], T2i[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) C[i] = dconst<ND>(cons[i]);
  pose_compose(C, T1, Tc);
  pose_inverse(T2, T2i);
  pose_compose(T2i, Tc, res);
}

}  // namespace slslam
