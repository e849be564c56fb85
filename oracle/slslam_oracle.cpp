// ORACLE — TEST INFRASTRUCTURE ONLY.  A CPU restatement of the reference's LBA / PO hot path:
//   cost functors           reference src/lba_problem.h:46-118, src/po_problem.h:27-105
//   problem semantics       reference src/lba_problem.cpp:54-132, src/po_problem.cpp:40-77
//   solver                  Ceres 1.7.0 (README:8) — NOT in /root/reference, not installed, cannot be
//                           built here; restated from its published algorithm.
// PARITY UNPINNED: the reference ships no tests/golden vectors for this path and cannot be compiled in
// this image, so this oracle is pinned only by (a) the geometric known-answer tests, (b) an independent
// numpy/torch-autograd restatement (tests/golden/make_golden.py) and (c) scipy.optimize.least_squares
// minima — never by a run of the reference binary.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this
// library.  The product path (slslam_b200/csrc) never links or calls it.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "jet.h"

namespace {

// ---------------------------------------------------------------------------------------------
// ceres/rotation.h (1.7.0) restated — SURVEY.md Appendix A1.
// ---------------------------------------------------------------------------------------------
template <typename T>
inline void AngleAxisRotatePoint(const T w[3], const T pt[3], T result[3]) {
  // used at reference lba_problem.h:75-76 and po_problem.h:38,59
  const T theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > 0.0) {
    const T theta = sqrt(theta2);
    const T k[3] = {w[0] / theta, w[1] / theta, w[2] / theta};
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T kxp[3] = {k[1] * pt[2] - k[2] * pt[1], k[2] * pt[0] - k[0] * pt[2], k[0] * pt[1] - k[1] * pt[0]};
    const T kdp = k[0] * pt[0] + k[1] * pt[1] + k[2] * pt[2];
    for (int i = 0; i < 3; ++i)
      result[i] = pt[i] * costheta + kxp[i] * sintheta + k[i] * (T(1.0) - costheta) * kdp;
  } else {
    // first-order Taylor branch R = I + [w]x : hit by the newest keyframe, which is exactly identity
    // (reference slam.cpp:1322).
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    for (int i = 0; i < 3; ++i) result[i] = pt[i] + wxp[i];
  }
}

template <typename T>
inline void AngleAxisToQuaternion(const T w[3], T q[4]) {
  const T theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > 0.0) {
    const T theta = sqrt(theta2);
    const T half = theta * T(0.5);
    const T k = sin(half) / theta;
    q[0] = cos(half); q[1] = w[0] * k; q[2] = w[1] * k; q[3] = w[2] * k;
  } else {
    const T k(0.5);
    q[0] = T(1.0); q[1] = w[0] * k; q[2] = w[1] * k; q[3] = w[2] * k;
  }
}

template <typename T>
inline void QuaternionToAngleAxis(const T q[4], T w[3]) {
  const T s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const T s = sqrt(s2);
    const T& c = q[0];
    // keep |angle| <= pi: atan2(-s,-c) when c < 0
    const T two_theta = T(2.0) * ((c < 0.0) ? atan2(-s, -c) : atan2(s, c));
    const T k = two_theta / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    const T k(2.0);
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  }
}

template <typename T>
inline void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}

// ---------------------------------------------------------------------------------------------
// LineReprojectionError::operator()  — reference src/lba_problem.h:46-118
// ---------------------------------------------------------------------------------------------
template <typename T>
inline void line_reprojection_error(const T* camera, const T* line, const double* ob, double baseline,
                                    T* residuals) {
  const T a = line[0], b = line[1], g = line[2], t = line[3];
  const T s1 = sin(a), c1 = cos(a), s2 = sin(b), c2 = cos(b), s3 = sin(g), c3 = cos(g);
  const T d = cos(t) / sin(t);                                      // :63
  T cp[3], dv[3];
  cp[0] = -(c1 * s2 * c3 + s1 * s3) * d;                             // :66-68
  cp[1] = -(c1 * s2 * s3 - s1 * c3) * d;
  cp[2] = -(c1 * c2) * d;
  dv[0] = s1 * s2 * c3 - c1 * s3;                                    // :70-72
  dv[1] = s1 * s2 * s3 + c1 * c3;
  dv[2] = s1 * c2;
  T pc[3], dc[3];
  AngleAxisRotatePoint(camera, cp, pc);                              // :75-76
  AngleAxisRotatePoint(camera, dv, dc);
  pc[0] += camera[3]; pc[1] += camera[4]; pc[2] += camera[5];        // :81-83
  T n[3];
  n[0] = pc[1] * dc[2] - pc[2] * dc[1];
  n[1] = pc[2] * dc[0] - pc[0] * dc[2];
  n[2] = pc[0] * dc[1] - pc[1] * dc[0];
  T sql = sqrt(n[0] * n[0] + n[1] * n[1]);                           // :90
  n[0] /= sql; n[1] /= sql; n[2] /= sql;
  residuals[0] = -(T(ob[0]) * n[0] + T(ob[1]) * n[1] + n[2]);        // :95-96
  residuals[1] = -(T(ob[2]) * n[0] + T(ob[3]) * n[1] + n[2]);
  pc[0] -= T(baseline);                                              // :101-103 (0.12 literal there)
  n[0] = pc[1] * dc[2] - pc[2] * dc[1];
  n[1] = pc[2] * dc[0] - pc[0] * dc[2];
  n[2] = pc[0] * dc[1] - pc[1] * dc[0];
  sql = sqrt(n[0] * n[0] + n[1] * n[1]);
  n[0] /= sql; n[1] /= sql; n[2] /= sql;
  residuals[2] = -(T(ob[4]) * n[0] + T(ob[5]) * n[1] + n[2]);        // :114-115
  residuals[3] = -(T(ob[6]) * n[0] + T(ob[7]) * n[1] + n[2]);
}

// ---------------------------------------------------------------------------------------------
// gc_T_inv / gc_w_20 / gc_T_20 / PoseConstraintError — reference src/po_problem.h:27-105
// ---------------------------------------------------------------------------------------------
template <typename T>
inline void po_T_inv(const T P[6], T Pi[6]) {                        // :27-39
  Pi[0] = -P[0]; Pi[1] = -P[1]; Pi[2] = -P[2];
  T v[3] = {-P[3], -P[4], -P[5]};
  AngleAxisRotatePoint(Pi, v, Pi + 3);
}
template <typename T>
inline void po_w_20(const T w21[3], const T w10[3], T w20[3]) {      // :42-52
  T q21[4], q10[4], q20[4];
  AngleAxisToQuaternion(w21, q21);
  AngleAxisToQuaternion(w10, q10);
  QuaternionProduct(q21, q10, q20);
  QuaternionToAngleAxis(q20, w20);
}
template <typename T>
inline void po_T_20(const T T21[6], const T T10[6], T T20[6]) {      // :55-64
  po_w_20(T21, T10, T20);
  AngleAxisRotatePoint(T21, T10 + 3, T20 + 3);
  T20[3] += T21[3]; T20[4] += T21[4]; T20[5] += T21[5];
}
template <typename T>
inline void pose_constraint_error(const T* pose1, const T* pose2, const double* c, T* residuals) {  // :73-105
  T T1[6], T2[6], C[6], Tc[6], Te[6], T2i[6];
  for (int i = 0; i < 6; ++i) { T1[i] = pose1[i]; T2[i] = pose2[i]; C[i] = T(c[i]); }
  po_T_20(C, T1, Tc);
  po_T_inv(T2, T2i);
  po_T_20(T2i, Tc, Te);
  for (int i = 0; i < 6; ++i) residuals[i] = Te[i];
}

// ---------------------------------------------------------------------------------------------
// Loss: ceres::HuberLoss(a) + Corrector with rho'' <= 0.
//   s <= a^2 : rho = s, rho' = 1 ; else rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s).
//   block cost = rho/2 ; residual and Jacobian rows are scaled by sqrt(rho').
// ---------------------------------------------------------------------------------------------
inline void huber(double s, double a, bool robust, double* rho, double* sqrt_rho1) {
  if (!robust || s <= a * a) { *rho = s; *sqrt_rho1 = 1.0; return; }
  const double r = std::sqrt(s);
  *rho = 2.0 * a * r - a * a;
  *sqrt_rho1 = std::sqrt(a / r);
}

// dense Cholesky A = L L^T on the lower triangle, row-major n x n; returns false if not PD.
bool cholesky_lower(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double* Aj = &A[(size_t)j * n];
    double d = Aj[j];
    for (int k = 0; k < j; ++k) d -= Aj[k] * Aj[k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double ljj = std::sqrt(d);
    Aj[j] = ljj;
    const double inv = 1.0 / ljj;
    for (int i = j + 1; i < n; ++i) {
      double* Ai = &A[(size_t)i * n];
      double s = Ai[j];
      for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];
      Ai[j] = s * inv;
    }
  }
  return true;
}
void cholesky_solve(const std::vector<double>& L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    const double* Li = &L[(size_t)i * n];
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= Li[k] * b[k];
    b[i] = s / Li[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

// ---------------------------------------------------------------------------------------------
// Ceres 1.7.0 TrustRegionMinimizer + LevenbergMarquardtStrategy, restated.
// ---------------------------------------------------------------------------------------------
enum { TERM_NO_CONVERGENCE = 0, TERM_GRADIENT = 1, TERM_FUNCTION = 2, TERM_PARAMETER = 3, TERM_FAILURE = 4 };
enum { TRACE_W = 8 };  // cost, trial_cost, model_change, radius, step_norm, accepted, grad_max, rel_decrease

struct LMOptions {
  int max_iterations = 10;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  double min_relative_decrease = 1e-3;
  double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
  double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  int max_consecutive_invalid = 5;
};
struct LMSummary {
  double initial_cost = 0, final_cost = 0, fixed_cost = 0, grad_max = 0;
  int successful = 0, unsuccessful = 0, termination = TERM_NO_CONVERGENCE, iterations = 0;
};

// P must provide: n(), x0(x), cost(x) [reduced], linearize(x) -> cost, gradient(g) [unscaled J],
// col_sqnorm(scale,out), solve(scale,D2,y) -> bool, model_change(scale,step), fixed_cost(), store(x).
template <class P>
void levenberg_marquardt(P& prob, const LMOptions& opt, LMSummary* sum, double* trace) {
  const int n = prob.n();
  sum->fixed_cost = prob.fixed_cost();
  std::vector<double> x(n), xt(n), scale(n, 1.0), diag(n), D2(n), y(n), step(n), g(n), delta(n);
  prob.x0(x.data());
  double cost = (n > 0) ? prob.linearize(x.data()) : 0.0;
  sum->initial_cost = cost + sum->fixed_cost;
  sum->final_cost = sum->initial_cost;
  if (n == 0) { sum->termination = TERM_GRADIENT; return; }
  prob.col_sqnorm(scale.data(), diag.data());  // scale==1 here: unscaled column norms
  for (int j = 0; j < n; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(diag[j]));
  prob.gradient(g.data());
  double gmax = 0.0;
  for (int j = 0; j < n; ++j) gmax = std::max(gmax, std::fabs(g[j]));
  sum->grad_max = gmax;
  const double gtol = opt.gradient_tolerance * std::max(gmax, std::numeric_limits<double>::epsilon());
  if (gmax <= gtol) { sum->termination = TERM_GRADIENT; return; }
  double x_norm = 0.0;
  for (int j = 0; j < n; ++j) x_norm += x[j] * x[j];
  x_norm = std::sqrt(x_norm);

  double radius = opt.initial_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid = 0;
  sum->termination = TERM_NO_CONVERGENCE;
  for (int it = 0; it < opt.max_iterations; ++it) {
    sum->iterations = it + 1;
    double* tr = trace ? trace + (size_t)it * TRACE_W : nullptr;
    if (tr) { for (int k = 0; k < TRACE_W; ++k) tr[k] = 0.0; tr[0] = cost; tr[3] = radius; tr[6] = gmax; }
    if (!reuse_diagonal) {
      prob.col_sqnorm(scale.data(), diag.data());
      for (int j = 0; j < n; ++j) diag[j] = std::min(std::max(diag[j], opt.min_lm_diagonal), opt.max_lm_diagonal);
    }
    for (int j = 0; j < n; ++j) D2[j] = diag[j] / radius;
    bool ok = prob.solve(scale.data(), D2.data(), y.data());
    double model = 0.0;
    if (ok) {
      for (int j = 0; j < n; ++j) { step[j] = -y[j]; if (!std::isfinite(step[j])) ok = false; }
      if (ok) model = prob.model_change(scale.data(), step.data());
    }
    if (tr) tr[2] = model;
    if (!ok || model < 0.0) {
      // invalid step: LevenbergMarquardtStrategy::StepIsInvalid
      ++sum->unsuccessful;
      if (tr) tr[5] = -1.0;
      if (++invalid >= opt.max_consecutive_invalid) { sum->termination = TERM_FAILURE; break; }
      radius *= 0.5; reuse_diagonal = true;
      if (radius < opt.min_radius) { sum->termination = TERM_PARAMETER; break; }
      continue;
    }
    invalid = 0;
    double step_norm = 0.0;
    for (int j = 0; j < n; ++j) { delta[j] = step[j] * scale[j]; xt[j] = x[j] + delta[j]; step_norm += delta[j] * delta[j]; }
    step_norm = std::sqrt(step_norm);
    const double new_cost = prob.cost(xt.data());
    if (tr) { tr[1] = new_cost; tr[4] = step_norm; }
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      sum->termination = TERM_PARAMETER; break;
    }
    const double cost_change = cost - new_cost;
    if (std::fabs(cost_change) < opt.function_tolerance * cost) {
      // 1.7.0 returns here without adopting the trial point.
      sum->termination = TERM_FUNCTION; break;
    }
    const double rel = cost_change / model;
    if (tr) tr[7] = rel;
    if (rel > opt.min_relative_decrease) {
      ++sum->successful;
      if (tr) tr[5] = 1.0;
      x = xt;
      x_norm = 0.0;
      for (int j = 0; j < n; ++j) x_norm += x[j] * x[j];
      x_norm = std::sqrt(x_norm);
      cost = prob.linearize(x.data());
      prob.gradient(g.data());
      gmax = 0.0;
      for (int j = 0; j < n; ++j) gmax = std::max(gmax, std::fabs(g[j]));
      sum->grad_max = gmax;
      // LevenbergMarquardtStrategy::StepAccepted
      radius = std::min(opt.max_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3)));
      decrease_factor = 2.0; reuse_diagonal = false;
      if (gmax <= gtol) { sum->termination = TERM_GRADIENT; break; }
    } else {
      ++sum->unsuccessful;
      // StepRejected
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
    if (radius < opt.min_radius) { sum->termination = TERM_PARAMETER; break; }
  }
  sum->final_cost = cost + sum->fixed_cost;
  prob.store(x.data());
}

// ---------------------------------------------------------------------------------------------
// LBA program: what LBAProblem::build hands to Ceres (reference src/lba_problem.cpp:54-93)
// ---------------------------------------------------------------------------------------------
struct LBAProgram {
  int C, L, N;
  const int *cam_idx, *line_idx;
  const double* obs;
  bool robust; double huber_a, baseline;
  int solver;  // 0: normal equations on the full reduced vector (what SPARSE_NORMAL_CHOLESKY solves,
               //    lba_problem.cpp:96-101 always ends there); 1: Schur on the line blocks (same step).
  double* params;                       // [6C + 4L], updated in place by store()
  std::vector<int> cam_slot, line_slot; // reduced block index or -1 (constant / unused)
  std::vector<char> active;             // residual block has >= 1 free parameter block
  int Cf = 0, Lf = 0;
  double fixed_cost_ = 0.0;
  // linearisation at the last linearize(): Huber-scaled, Jacobi-unscaled
  std::vector<double> r, Jc, Jl;        // [4N], [24N] row-major 4x6, [16N] row-major 4x4

  void setup(const int* fixed_idx) {
    std::vector<char> cam_used(C, 0), line_used(L, 0), cam_const(C, 0), line_const(L, 0);
    for (int i = 0; i < N; ++i) {
      cam_used[cam_idx[i]] = 1; line_used[line_idx[i]] = 1;
      if (fixed_idx[2 * i]) cam_const[cam_idx[i]] = 1;         // sticky per block: lba_problem.cpp:88-91
      if (fixed_idx[2 * i + 1]) line_const[line_idx[i]] = 1;
    }
    cam_slot.assign(C, -1); line_slot.assign(L, -1);
    for (int c = 0; c < C; ++c) if (cam_used[c] && !cam_const[c]) cam_slot[c] = Cf++;
    for (int l = 0; l < L; ++l) if (line_used[l] && !line_const[l]) line_slot[l] = Lf++;
    active.assign(N, 0);
    for (int i = 0; i < N; ++i) active[i] = (cam_slot[cam_idx[i]] >= 0 || line_slot[line_idx[i]] >= 0);
    r.resize(4 * (size_t)N); Jc.resize(24 * (size_t)N); Jl.resize(16 * (size_t)N);
    // residual blocks whose blocks are all constant leave the program; their cost stays in the summary
    fixed_cost_ = 0.0;
    for (int i = 0; i < N; ++i) if (!active[i]) {
      double res[4];
      line_reprojection_error<double>(params + 6 * cam_idx[i], params + 6 * C + 4 * line_idx[i], obs + 8 * i, baseline, res);
      const double s = res[0] * res[0] + res[1] * res[1] + res[2] * res[2] + res[3] * res[3];
      double rho, w; huber(s, huber_a, robust, &rho, &w);
      fixed_cost_ += 0.5 * rho;
    }
  }
  int n() const { return 6 * Cf + 4 * Lf; }
  double fixed_cost() const { return fixed_cost_; }
  void x0(double* x) const {
    for (int c = 0; c < C; ++c) if (cam_slot[c] >= 0) std::memcpy(x + 6 * cam_slot[c], params + 6 * c, 6 * sizeof(double));
    for (int l = 0; l < L; ++l) if (line_slot[l] >= 0) std::memcpy(x + 6 * Cf + 4 * line_slot[l], params + 6 * C + 4 * l, 4 * sizeof(double));
  }
  void store(const double* x) {
    for (int c = 0; c < C; ++c) if (cam_slot[c] >= 0) std::memcpy(params + 6 * c, x + 6 * cam_slot[c], 6 * sizeof(double));
    for (int l = 0; l < L; ++l) if (line_slot[l] >= 0) std::memcpy(params + 6 * C + 4 * l, x + 6 * Cf + 4 * line_slot[l], 4 * sizeof(double));
  }
  const double* cam_ptr(const double* x, int c) const { return cam_slot[c] >= 0 ? x + 6 * cam_slot[c] : params + 6 * c; }
  const double* line_ptr(const double* x, int l) const { return line_slot[l] >= 0 ? x + 6 * Cf + 4 * line_slot[l] : params + 6 * C + 4 * l; }

  double cost(const double* x) const {
    double total = 0.0;
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      double res[4];
      line_reprojection_error<double>(cam_ptr(x, cam_idx[i]), line_ptr(x, line_idx[i]), obs + 8 * i, baseline, res);
      const double s = res[0] * res[0] + res[1] * res[1] + res[2] * res[2] + res[3] * res[3];
      double rho, w; huber(s, huber_a, robust, &rho, &w);
      total += 0.5 * rho;
    }
    return total;
  }
  double linearize(const double* x) {
    typedef Jet<10> J;
    double total = 0.0;
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const double* cp = cam_ptr(x, cam_idx[i]);
      const double* lp = line_ptr(x, line_idx[i]);
      J cam[6], line[4], res[4];
      for (int k = 0; k < 6; ++k) cam[k] = J(cp[k], k);
      for (int k = 0; k < 4; ++k) line[k] = J(lp[k], 6 + k);
      line_reprojection_error<J>(cam, line, obs + 8 * i, baseline, res);
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += res[k].a * res[k].a;
      double rho, w; huber(s, huber_a, robust, &rho, &w);
      total += 0.5 * rho;
      const bool cfree = cam_slot[cam_idx[i]] >= 0, lfree = line_slot[line_idx[i]] >= 0;
      for (int k = 0; k < 4; ++k) {
        r[4 * (size_t)i + k] = w * res[k].a;
        for (int j = 0; j < 6; ++j) Jc[24 * (size_t)i + 6 * k + j] = cfree ? w * res[k].v[j] : 0.0;
        for (int j = 0; j < 4; ++j) Jl[16 * (size_t)i + 4 * k + j] = lfree ? w * res[k].v[6 + j] : 0.0;
      }
    }
    return total;
  }
  void gradient(double* g) const {
    std::fill(g, g + n(), 0.0);
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const int cs = cam_slot[cam_idx[i]], ls = line_slot[line_idx[i]];
      for (int k = 0; k < 4; ++k) {
        const double rk = r[4 * (size_t)i + k];
        if (cs >= 0) for (int j = 0; j < 6; ++j) g[6 * cs + j] += Jc[24 * (size_t)i + 6 * k + j] * rk;
        if (ls >= 0) for (int j = 0; j < 4; ++j) g[6 * Cf + 4 * ls + j] += Jl[16 * (size_t)i + 4 * k + j] * rk;
      }
    }
  }
  void col_sqnorm(const double* scale, double* out) const {
    std::fill(out, out + n(), 0.0);
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const int cs = cam_slot[cam_idx[i]], ls = line_slot[line_idx[i]];
      for (int k = 0; k < 4; ++k) {
        if (cs >= 0) for (int j = 0; j < 6; ++j) { const double v = Jc[24 * (size_t)i + 6 * k + j] * scale[6 * cs + j]; out[6 * cs + j] += v * v; }
        if (ls >= 0) for (int j = 0; j < 4; ++j) { const double v = Jl[16 * (size_t)i + 4 * k + j] * scale[6 * Cf + 4 * ls + j]; out[6 * Cf + 4 * ls + j] += v * v; }
      }
    }
  }
  double model_change(const double* scale, const double* step) const {
    // -(m . (r + m/2)), m = J_scaled * step     (TrustRegionMinimizer)
    double acc = 0.0;
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const int cs = cam_slot[cam_idx[i]], ls = line_slot[line_idx[i]];
      for (int k = 0; k < 4; ++k) {
        double m = 0.0;
        if (cs >= 0) for (int j = 0; j < 6; ++j) m += Jc[24 * (size_t)i + 6 * k + j] * scale[6 * cs + j] * step[6 * cs + j];
        if (ls >= 0) for (int j = 0; j < 4; ++j) m += Jl[16 * (size_t)i + 4 * k + j] * scale[6 * Cf + 4 * ls + j] * step[6 * Cf + 4 * ls + j];
        acc += m * (r[4 * (size_t)i + k] + 0.5 * m);
      }
    }
    return -acc;
  }
  bool solve(const double* scale, const double* D2, double* y) const { return solver == 0 ? solve_full(scale, D2, y) : solve_schur(scale, D2, y); }

  // (J^T J + D^2) y = J^T r on the whole reduced vector, dense Cholesky.
  bool solve_full(const double* scale, const double* D2, double* y) const {
    const int nn = n();
    std::vector<double> H((size_t)nn * nn, 0.0);
    std::fill(y, y + nn, 0.0);
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const int cs = cam_slot[cam_idx[i]], ls = line_slot[line_idx[i]];
      int idx[10]; double row[10]; int m;
      for (int k = 0; k < 4; ++k) {
        m = 0;
        if (cs >= 0) for (int j = 0; j < 6; ++j) { idx[m] = 6 * cs + j; row[m++] = Jc[24 * (size_t)i + 6 * k + j] * scale[6 * cs + j]; }
        if (ls >= 0) for (int j = 0; j < 4; ++j) { idx[m] = 6 * Cf + 4 * ls + j; row[m++] = Jl[16 * (size_t)i + 4 * k + j] * scale[6 * Cf + 4 * ls + j]; }
        const double rk = r[4 * (size_t)i + k];
        for (int p = 0; p < m; ++p) {
          y[idx[p]] += row[p] * rk;
          for (int q = 0; q <= p; ++q) {
            const int a = std::max(idx[p], idx[q]), b = std::min(idx[p], idx[q]);
            H[(size_t)a * nn + b] += row[p] * row[q];
          }
        }
      }
    }
    for (int j = 0; j < nn; ++j) H[(size_t)j * nn + j] += D2[j];
    if (!cholesky_lower(H, nn)) return false;
    cholesky_solve(H, nn, y);
    return true;
  }

  // Schur complement on the line blocks: mathematically the same step.
  bool solve_schur(const double* scale, const double* D2, double* y) const {
    const int nc = 6 * Cf;
    std::vector<double> S((size_t)nc * nc, 0.0), bc(nc, 0.0);
    std::vector<double> Hll((size_t)Lf * 16, 0.0), gl((size_t)Lf * 4, 0.0);
    std::vector<std::vector<int> > by_line(Lf);
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      const int cs = cam_slot[cam_idx[i]], ls = line_slot[line_idx[i]];
      double A[4][6], B[4][4];
      for (int k = 0; k < 4; ++k) {
        for (int j = 0; j < 6; ++j) A[k][j] = cs >= 0 ? Jc[24 * (size_t)i + 6 * k + j] * scale[6 * cs + j] : 0.0;
        for (int j = 0; j < 4; ++j) B[k][j] = ls >= 0 ? Jl[16 * (size_t)i + 4 * k + j] * scale[6 * Cf + 4 * ls + j] : 0.0;
      }
      if (cs >= 0) for (int p = 0; p < 6; ++p) {
        for (int k = 0; k < 4; ++k) bc[6 * cs + p] += A[k][p] * r[4 * (size_t)i + k];
        for (int q = 0; q < 6; ++q) { double s = 0; for (int k = 0; k < 4; ++k) s += A[k][p] * A[k][q]; S[(size_t)(6 * cs + p) * nc + 6 * cs + q] += s; }
      }
      if (ls >= 0) {
        by_line[ls].push_back(i);
        for (int p = 0; p < 4; ++p) {
          for (int k = 0; k < 4; ++k) gl[4 * (size_t)ls + p] += B[k][p] * r[4 * (size_t)i + k];
          for (int q = 0; q < 4; ++q) { double s = 0; for (int k = 0; k < 4; ++k) s += B[k][p] * B[k][q]; Hll[16 * (size_t)ls + 4 * p + q] += s; }
        }
      }
    }
    for (int j = 0; j < nc; ++j) S[(size_t)j * nc + j] += D2[j];
    // eliminate each line: Hll + D_l^2 = L L^T ; Z_i = (Jc_i^T Jl_i) L^-T ; u = L^-1 g_l
    std::vector<double> Ls((size_t)Lf * 16), us((size_t)Lf * 4);
    std::vector<std::vector<double> > Zs(Lf);
    for (int l = 0; l < Lf; ++l) {
      std::vector<double> M(Hll.begin() + 16 * (size_t)l, Hll.begin() + 16 * (size_t)l + 16);
      for (int p = 0; p < 4; ++p) M[4 * p + p] += D2[nc + 4 * l + p];
      if (!cholesky_lower(M, 4)) return false;
      double u[4];
      for (int p = 0; p < 4; ++p) { double s = gl[4 * (size_t)l + p]; for (int k = 0; k < p; ++k) s -= M[4 * p + k] * u[k]; u[p] = s / M[4 * p + p]; }
      std::copy(M.begin(), M.end(), Ls.begin() + 16 * (size_t)l);
      std::copy(u, u + 4, us.begin() + 4 * (size_t)l);
      const std::vector<int>& ob = by_line[l];
      std::vector<double>& Z = Zs[l];
      Z.assign(ob.size() * 24, 0.0);
      for (size_t a = 0; a < ob.size(); ++a) {
        const int i = ob[a]; const int cs = cam_slot[cam_idx[i]];
        if (cs < 0) continue;
        for (int p = 0; p < 6; ++p) {
          double W[4];
          for (int q = 0; q < 4; ++q) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += Jc[24 * (size_t)i + 6 * k + p] * scale[6 * cs + p] * Jl[16 * (size_t)i + 4 * k + q] * scale[nc + 4 * l + q];
            W[q] = s;
          }
          // row of Z: solve z L^T = W  ->  z_q = (W_q - sum_{k<q} z_k L[q][k]) / L[q][q]
          for (int q = 0; q < 4; ++q) { double s = W[q]; for (int k = 0; k < q; ++k) s -= Z[a * 24 + 4 * p + k] * M[4 * q + k]; Z[a * 24 + 4 * p + q] = s / M[4 * q + q]; }
        }
      }
      for (size_t a = 0; a < ob.size(); ++a) {
        const int ca = cam_slot[cam_idx[ob[a]]]; if (ca < 0) continue;
        for (int p = 0; p < 6; ++p) { double s = 0; for (int k = 0; k < 4; ++k) s += Z[a * 24 + 4 * p + k] * u[k]; bc[6 * ca + p] -= s; }
        for (size_t b = 0; b < ob.size(); ++b) {
          const int cb = cam_slot[cam_idx[ob[b]]]; if (cb < 0) continue;
          for (int p = 0; p < 6; ++p) for (int q = 0; q < 6; ++q) {
            double s = 0; for (int k = 0; k < 4; ++k) s += Z[a * 24 + 4 * p + k] * Z[b * 24 + 4 * q + k];
            S[(size_t)(6 * ca + p) * nc + 6 * cb + q] -= s;
          }
        }
      }
    }
    if (nc > 0) {
      if (!cholesky_lower(S, nc)) return false;
      cholesky_solve(S, nc, bc.data());
    }
    std::copy(bc.begin(), bc.end(), y);
    // back-substitute: y_l = L^-T (u - sum_i Z_i^T y_c(i))
    for (int l = 0; l < Lf; ++l) {
      const double* M = &Ls[16 * (size_t)l];
      double v[4] = {us[4 * (size_t)l], us[4 * (size_t)l + 1], us[4 * (size_t)l + 2], us[4 * (size_t)l + 3]};
      const std::vector<int>& ob = by_line[l];
      for (size_t a = 0; a < ob.size(); ++a) {
        const int ca = cam_slot[cam_idx[ob[a]]]; if (ca < 0) continue;
        for (int q = 0; q < 4; ++q) for (int p = 0; p < 6; ++p) v[q] -= Zs[l][a * 24 + 4 * p + q] * bc[6 * ca + p];
      }
      for (int p = 3; p >= 0; --p) { double s = v[p]; for (int k = p + 1; k < 4; ++k) s -= M[4 * k + p] * v[k]; v[p] = s / M[4 * p + p]; }
      for (int p = 0; p < 4; ++p) y[nc + 4 * l + p] = v[p];
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// Sparse Cholesky: what SPARSE_NORMAL_CHOLESKY (reference src/po_problem.cpp:68) ends in.  Ceres 1.7.0 hands the
// normal equations to CHOLMOD or CXSparse -- neither is in /root/reference nor installed --: a fill-reducing ordering,
// a symbolic analysis through the elimination tree, an up-looking numeric factorisation and two triangular solves.
// Restated here from the published algorithm (T. Davis, "Direct Methods for Sparse Linear Systems", ch. 4: etree,
// ereach, up-looking cs_chol); the ordering is an exact greedy minimum-degree on the 6x6-block graph instead of AMD
// (an ordering changes fill and rounding, not the solution).
// ---------------------------------------------------------------------------------------------
struct SparseSym {                 // upper triangle of the permuted matrix C = P A P^T, compressed columns
  int n = 0;
  std::vector<int> Ap, Ai;
  std::vector<double> Ax;
};
struct SparseChol {
  int n = 0;
  std::vector<int> parent, Lp, Li;
  std::vector<double> Lx;
  long long flops = 0;
  // elimination tree of C (upper part given): parent[i] = min { j > i : L(j,i) != 0 }
  void analyze(const SparseSym& C) {
    n = C.n;
    parent.assign(n, -1);
    std::vector<int> ancestor(n, -1);
    for (int k = 0; k < n; ++k) {
      for (int p = C.Ap[k]; p < C.Ap[k + 1]; ++p) {
        int i = C.Ai[p];
        while (i != -1 && i < k) {
          const int inext = ancestor[i];
          ancestor[i] = k;
          if (inext == -1) parent[i] = k;
          i = inext;
        }
      }
    }
    // column counts by walking every row's reach once (row subtree of the elimination tree)
    std::vector<int> cnt(n, 1), mark(n, -1);
    for (int k = 0; k < n; ++k) {
      mark[k] = k;
      for (int p = C.Ap[k]; p < C.Ap[k + 1]; ++p) {
        int i = C.Ai[p];
        while (i < k && mark[i] != k) { ++cnt[i]; mark[i] = k; i = parent[i]; }
      }
    }
    Lp.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) Lp[j + 1] = Lp[j] + cnt[j];
    Li.assign(Lp[n], 0); Lx.assign(Lp[n], 0.0);
  }
  // up-looking numeric factorisation: row k of L from a sparse triangular solve with the rows above
  bool factor(const SparseSym& C) {
    std::vector<int> c(Lp.begin(), Lp.end() - 1), stack(n), w(n, -1), pattern(n);
    std::vector<double> x(n, 0.0);
    flops = 0;
    for (int k = 0; k < n; ++k) {
      // ereach: nonzero pattern of row k of L, in topological order on top of `pattern`
      int top = n;
      w[k] = k;
      for (int p = C.Ap[k]; p < C.Ap[k + 1]; ++p) {
        int i = C.Ai[p];
        if (i > k) continue;
        x[i] = C.Ax[p];
        int len = 0;
        for (; w[i] != k; i = parent[i]) { stack[len++] = i; w[i] = k; }
        while (len > 0) pattern[--top] = stack[--len];
      }
      double d = x[k];
      x[k] = 0.0;
      for (; top < n; ++top) {
        const int i = pattern[top];
        const double lki = x[i] / Lx[Lp[i]];
        x[i] = 0.0;
        for (int p = Lp[i] + 1; p < c[i]; ++p) x[Li[p]] -= Lx[p] * lki;
        flops += 2 * (long long)(c[i] - Lp[i]);
        d -= lki * lki;
        const int q = c[i]++;
        Li[q] = k; Lx[q] = lki;
      }
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      const int q = c[k]++;
      Li[q] = k; Lx[q] = std::sqrt(d);
    }
    return true;
  }
  void solve(double* b) const {          // L L^T x = b in place
    for (int j = 0; j < n; ++j) {
      b[j] /= Lx[Lp[j]];
      for (int p = Lp[j] + 1; p < Lp[j + 1]; ++p) b[Li[p]] -= Lx[p] * b[j];
    }
    for (int j = n - 1; j >= 0; --j) {
      for (int p = Lp[j] + 1; p < Lp[j + 1]; ++p) b[j] -= Lx[p] * b[Li[p]];
      b[j] /= Lx[Lp[j]];
    }
  }
};

// Exact greedy minimum-degree elimination order of an undirected graph (adjacency without self loops); ties by index.
std::vector<int> minimum_degree_order(int nv, std::vector<std::vector<int> > adj) {
  std::vector<std::vector<char> > has(nv, std::vector<char>(nv, 0));
  for (int v = 0; v < nv; ++v) for (int u : adj[v]) has[v][u] = 1;
  std::vector<char> gone(nv, 0);
  std::vector<int> order;
  order.reserve(nv);
  for (int step = 0; step < nv; ++step) {
    int best = -1, bestdeg = 1 << 30;
    for (int v = 0; v < nv; ++v) {
      if (gone[v]) continue;
      int deg = 0;
      for (int u : adj[v]) if (!gone[u]) ++deg;
      if (deg < bestdeg) { bestdeg = deg; best = v; }
    }
    gone[best] = 1;
    order.push_back(best);
    std::vector<int> nb;
    for (int u : adj[best]) if (!gone[u]) nb.push_back(u);
    for (size_t a = 0; a < nb.size(); ++a)
      for (size_t b = a + 1; b < nb.size(); ++b)
        if (!has[nb[a]][nb[b]]) { has[nb[a]][nb[b]] = has[nb[b]][nb[a]] = 1; adj[nb[a]].push_back(nb[b]); adj[nb[b]].push_back(nb[a]); }
  }
  return order;
}

// ---------------------------------------------------------------------------------------------
// PO program: what POProblem::build hands to Ceres (reference src/po_problem.cpp:40-65)
// ---------------------------------------------------------------------------------------------
struct POProgram {
  int K, E;
  const int *idx1, *idx2;
  const double* cons;
  double* params;                 // [6K]
  std::vector<int> slot;          // reduced block index or -1
  int Kf = 0;
  double fixed_cost_ = 0.0;
  std::vector<char> active;
  std::vector<double> r, J1, J2;  // [6E], [36E], [36E]
  int solver = 1;                 // 1: sparse Cholesky (what the reference selects, po_problem.cpp:68); 0: dense (cross-check)
  std::vector<int> pos;           // [Kf] position of every reduced block in the elimination order
  SparseSym C;                    // pattern of the upper triangle of P H P^T (values refilled by every solve)
  mutable SparseChol chol;
  std::vector<int> blk_col_start; // scratch of the assembly: first entry of every scalar column
  long long last_factor_flops() const { return chol.flops; }

  void setup_sparse() {
    // block graph of the free poses, minimum-degree order, scalar pattern (6x6 blocks, upper triangle, sorted rows)
    std::vector<std::vector<int> > adj(Kf);
    for (int e = 0; e < E; ++e) {
      const int a = slot[idx1[e]], b = slot[idx2[e]];
      if (a >= 0 && b >= 0 && a != b) { adj[a].push_back(b); adj[b].push_back(a); }
    }
    for (int v = 0; v < Kf; ++v) { std::sort(adj[v].begin(), adj[v].end()); adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end()); }
    const std::vector<int> order = minimum_degree_order(Kf, adj);
    pos.assign(Kf, 0);
    for (int k = 0; k < Kf; ++k) pos[order[k]] = k;
    const int nn = 6 * Kf;
    C.n = nn; C.Ap.assign(nn + 1, 0); C.Ai.clear();
    for (int cb = 0; cb < Kf; ++cb) {
      const int v = order[cb];
      std::vector<int> rows;                       // block rows above the diagonal in column cb
      for (int u : adj[v]) if (pos[u] < cb) rows.push_back(pos[u]);
      std::sort(rows.begin(), rows.end());
      for (int q = 0; q < 6; ++q) {
        for (int rb : rows) for (int p = 0; p < 6; ++p) C.Ai.push_back(6 * rb + p);
        for (int p = 0; p <= q; ++p) C.Ai.push_back(6 * cb + p);
        C.Ap[6 * cb + q + 1] = (int)C.Ai.size();
      }
    }
    C.Ax.assign(C.Ai.size(), 0.0);
    chol.analyze(C);
  }
  double* entry(int row, int col) {              // row <= col, permuted scalar indices; the pattern holds every block fully
    const int* b = &C.Ai[C.Ap[col]];
    const int* e = &C.Ai[C.Ap[col + 1]];
    const int* it = std::lower_bound(b, e, row);
    return &C.Ax[it - &C.Ai[0]];
  }
  bool solve_sparse(const double* scale, const double* D2, double* y) {
    const int nn = n();
    std::fill(C.Ax.begin(), C.Ax.end(), 0.0);
    std::vector<double> b(nn, 0.0);
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const int s1 = slot[idx1[e]], s2 = slot[idx2[e]];
      for (int k = 0; k < 6; ++k) {
        int idx[12]; double row[12]; int m = 0;
        if (s1 >= 0) for (int j = 0; j < 6; ++j) { idx[m] = 6 * pos[s1] + j; row[m++] = J1[36 * (size_t)e + 6 * k + j] * scale[6 * s1 + j]; }
        if (s2 >= 0 && idx1[e] != idx2[e]) for (int j = 0; j < 6; ++j) { idx[m] = 6 * pos[s2] + j; row[m++] = J2[36 * (size_t)e + 6 * k + j] * scale[6 * s2 + j]; }
        const double rk = r[6 * (size_t)e + k];
        for (int p = 0; p < m; ++p) {
          b[idx[p]] += row[p] * rk;
          for (int q = 0; q <= p; ++q) *entry(std::min(idx[p], idx[q]), std::max(idx[p], idx[q])) += row[p] * row[q];
        }
      }
    }
    for (int s0 = 0; s0 < Kf; ++s0) for (int j = 0; j < 6; ++j) *entry(6 * pos[s0] + j, 6 * pos[s0] + j) += D2[6 * s0 + j];
    if (!chol.factor(C)) return false;
    chol.solve(b.data());
    for (int s0 = 0; s0 < Kf; ++s0) for (int j = 0; j < 6; ++j) y[6 * s0 + j] = b[6 * pos[s0] + j];
    return true;
  }

  void setup() {
    std::vector<char> used(K, 0);
    for (int e = 0; e < E; ++e) { used[idx1[e]] = 1; used[idx2[e]] = 1; }
    slot.assign(K, -1);
    const int konst = E > 0 ? idx1[0] : -1;       // po_problem.cpp:62-63: pose1 of edge 0 is constant
    for (int k = 0; k < K; ++k) if (used[k] && k != konst) slot[k] = Kf++;
    active.assign(E, 0);
    fixed_cost_ = 0.0;
    for (int e = 0; e < E; ++e) {
      active[e] = slot[idx1[e]] >= 0 || slot[idx2[e]] >= 0;
      if (!active[e]) {
        double res[6];
        pose_constraint_error<double>(params + 6 * idx1[e], params + 6 * idx2[e], cons + 6 * e, res);
        for (int k = 0; k < 6; ++k) fixed_cost_ += 0.5 * res[k] * res[k];
      }
    }
    r.resize(6 * (size_t)E); J1.resize(36 * (size_t)E); J2.resize(36 * (size_t)E);
    if (solver == 1 && Kf > 0) setup_sparse();
  }
  int n() const { return 6 * Kf; }
  double fixed_cost() const { return fixed_cost_; }
  void x0(double* x) const { for (int k = 0; k < K; ++k) if (slot[k] >= 0) std::memcpy(x + 6 * slot[k], params + 6 * k, 48); }
  void store(const double* x) { for (int k = 0; k < K; ++k) if (slot[k] >= 0) std::memcpy(params + 6 * k, x + 6 * slot[k], 48); }
  const double* pose_ptr(const double* x, int k) const { return slot[k] >= 0 ? x + 6 * slot[k] : params + 6 * k; }
  double cost(const double* x) const {
    double total = 0.0;
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      double res[6];
      pose_constraint_error<double>(pose_ptr(x, idx1[e]), pose_ptr(x, idx2[e]), cons + 6 * e, res);
      for (int k = 0; k < 6; ++k) total += 0.5 * res[k] * res[k];
    }
    return total;
  }
  double linearize(const double* x) {
    typedef Jet<12> J;
    double total = 0.0;
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const double* p1 = pose_ptr(x, idx1[e]); const double* p2 = pose_ptr(x, idx2[e]);
      J a[6], b[6], res[6];
      if (idx1[e] == idx2[e]) {
        // degenerate self edge: both arguments alias one block; derivative directions coincide
        for (int k = 0; k < 6; ++k) { a[k] = J(p1[k], k); b[k] = J(p2[k], k); }
      } else {
        for (int k = 0; k < 6; ++k) { a[k] = J(p1[k], k); b[k] = J(p2[k], 6 + k); }
      }
      pose_constraint_error<J>(a, b, cons + 6 * e, res);
      const bool f1 = slot[idx1[e]] >= 0, f2 = slot[idx2[e]] >= 0;
      for (int k = 0; k < 6; ++k) {
        r[6 * (size_t)e + k] = res[k].a;
        total += 0.5 * res[k].a * res[k].a;
        for (int j = 0; j < 6; ++j) {
          J1[36 * (size_t)e + 6 * k + j] = f1 ? res[k].v[j] : 0.0;
          J2[36 * (size_t)e + 6 * k + j] = (f2 && idx1[e] != idx2[e]) ? res[k].v[6 + j] : 0.0;
        }
      }
    }
    return total;
  }
  void gradient(double* g) const {
    std::fill(g, g + n(), 0.0);
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const int s1 = slot[idx1[e]], s2 = slot[idx2[e]];
      for (int k = 0; k < 6; ++k) for (int j = 0; j < 6; ++j) {
        if (s1 >= 0) g[6 * s1 + j] += J1[36 * (size_t)e + 6 * k + j] * r[6 * (size_t)e + k];
        if (s2 >= 0) g[6 * s2 + j] += J2[36 * (size_t)e + 6 * k + j] * r[6 * (size_t)e + k];
      }
    }
  }
  void col_sqnorm(const double* scale, double* out) const {
    std::fill(out, out + n(), 0.0);
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const int s1 = slot[idx1[e]], s2 = slot[idx2[e]];
      for (int k = 0; k < 6; ++k) for (int j = 0; j < 6; ++j) {
        if (s1 >= 0) { const double v = J1[36 * (size_t)e + 6 * k + j] * scale[6 * s1 + j]; out[6 * s1 + j] += v * v; }
        if (s2 >= 0) { const double v = J2[36 * (size_t)e + 6 * k + j] * scale[6 * s2 + j]; out[6 * s2 + j] += v * v; }
      }
    }
  }
  double model_change(const double* scale, const double* step) const {
    double acc = 0.0;
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const int s1 = slot[idx1[e]], s2 = slot[idx2[e]];
      for (int k = 0; k < 6; ++k) {
        double m = 0.0;
        for (int j = 0; j < 6; ++j) {
          if (s1 >= 0) m += J1[36 * (size_t)e + 6 * k + j] * scale[6 * s1 + j] * step[6 * s1 + j];
          if (s2 >= 0) m += J2[36 * (size_t)e + 6 * k + j] * scale[6 * s2 + j] * step[6 * s2 + j];
        }
        acc += m * (r[6 * (size_t)e + k] + 0.5 * m);
      }
    }
    return -acc;
  }
  bool solve(const double* scale, const double* D2, double* y) {
    if (solver == 1) return solve_sparse(scale, D2, y);
    const int nn = n();
    std::vector<double> H((size_t)nn * nn, 0.0);
    std::fill(y, y + nn, 0.0);
    for (int e = 0; e < E; ++e) {
      if (!active[e]) continue;
      const int s1 = slot[idx1[e]], s2 = slot[idx2[e]];
      for (int k = 0; k < 6; ++k) {
        int idx[12]; double row[12]; int m = 0;
        if (s1 >= 0) for (int j = 0; j < 6; ++j) { idx[m] = 6 * s1 + j; row[m++] = J1[36 * (size_t)e + 6 * k + j] * scale[6 * s1 + j]; }
        if (s2 >= 0 && idx1[e] != idx2[e]) for (int j = 0; j < 6; ++j) { idx[m] = 6 * s2 + j; row[m++] = J2[36 * (size_t)e + 6 * k + j] * scale[6 * s2 + j]; }
        const double rk = r[6 * (size_t)e + k];
        for (int p = 0; p < m; ++p) {
          y[idx[p]] += row[p] * rk;
          for (int q = 0; q <= p; ++q) {
            const int a = std::max(idx[p], idx[q]), b = std::min(idx[p], idx[q]);
            H[(size_t)a * nn + b] += row[p] * row[q];
          }
        }
      }
    }
    for (int j = 0; j < nn; ++j) H[(size_t)j * nn + j] += D2[j];
    if (!cholesky_lower(H, nn)) return false;
    cholesky_solve(H, nn, y);
    return true;
  }
};

void write_summary(const LMSummary& s, double* out) {
  out[0] = s.initial_cost; out[1] = s.final_cost; out[2] = s.successful; out[3] = s.unsuccessful;
  out[4] = s.termination; out[5] = s.iterations; out[6] = s.fixed_cost; out[7] = s.grad_max;
}

}  // namespace


// ---- RANSAC hypothesis scoring ----
// SLAM::reprojection_error (reference src/slam.cpp:691-726) restated with its mixed precision kept: `error` and `sql`
// are `float` in the reference, everything else double.  line = (closest point, direction) in the previous keyframe's
// frame; T maps that frame to the current one; ft = normalised stereo endpoints in the current frame.
static float ransac_reprojection_error(const double* ft, const double* R, const double* t_in, const double* line, double baseline) {
  float error = 0;
  double t[3] = {t_in[0], t_in[1], t_in[2]};
  const double* cp = line;
  const double* dv = line + 3;
  for (int i = 0; i < 2; ++i) {
    double p1[3], p2[3];
    if (i == 0) {
      p1[0] = ft[0]; p1[1] = ft[1]; p1[2] = 1; p2[0] = ft[2]; p2[1] = ft[3]; p2[2] = 1;
    } else {
      t[0] -= baseline;
      p1[0] = ft[4]; p1[1] = ft[5]; p1[2] = 1; p2[0] = ft[6]; p2[1] = ft[7]; p2[2] = 1;
    }
    double cpc[3], dvc[3];
    for (int r = 0; r < 3; ++r) {
      cpc[r] = (R[3 * r] * cp[0] + R[3 * r + 1] * cp[1] + R[3 * r + 2] * cp[2]) + t[r];     // gc_point_to_pose, src/gc.cpp:55-57
      dvc[r] = R[3 * r] * dv[0] + R[3 * r + 1] * dv[1] + R[3 * r + 2] * dv[2];
    }
    double nc[3] = {cpc[1] * dvc[2] - cpc[2] * dvc[1], cpc[2] * dvc[0] - cpc[0] * dvc[2], cpc[0] * dvc[1] - cpc[1] * dvc[0]};
    const float sql = (float)std::sqrt(nc[0] * nc[0] + nc[1] * nc[1]);
    for (int r = 0; r < 3; ++r) nc[r] /= sql;
    error += std::fabs((nc[0] * p1[0] + nc[1] * p1[1]) + nc[2] * p1[2]);
    error += std::fabs((nc[0] * p2[0] + nc[1] * p2[1]) + nc[2] * p2[2]);
  }
  return error / 4.0;
}

extern "C" {

int oracle_trace_width() { return TRACE_W; }

// residual + AutoDiff Jacobians of one observation (no loss, no scaling). Jc row-major 4x6, Jl 4x4.
void oracle_lba_residual_jacobian(const double* cam, const double* line, const double* ob, double baseline,
                                  double* res, double* Jc, double* Jl) {
  typedef Jet<10> J;
  J c[6], l[4], r[4];
  for (int k = 0; k < 6; ++k) c[k] = J(cam[k], k);
  for (int k = 0; k < 4; ++k) l[k] = J(line[k], 6 + k);
  line_reprojection_error<J>(c, l, ob, baseline, r);
  for (int k = 0; k < 4; ++k) {
    res[k] = r[k].a;
    if (Jc) for (int j = 0; j < 6; ++j) Jc[6 * k + j] = r[k].v[j];
    if (Jl) for (int j = 0; j < 4; ++j) Jl[4 * k + j] = r[k].v[6 + j];
  }
}

void oracle_po_residual_jacobian(const double* p1, const double* p2, const double* c, double* res, double* J1, double* J2) {
  typedef Jet<12> J;
  J a[6], b[6], r[6];
  for (int k = 0; k < 6; ++k) { a[k] = J(p1[k], k); b[k] = J(p2[k], 6 + k); }
  pose_constraint_error<J>(a, b, c, r);
  for (int k = 0; k < 6; ++k) {
    res[k] = r[k].a;
    if (J1) for (int j = 0; j < 6; ++j) J1[6 * k + j] = r[k].v[j];
    if (J2) for (int j = 0; j < 6; ++j) J2[6 * k + j] = r[k].v[6 + j];
  }
}

// total robustified cost 1/2 sum rho(|r_i|^2) over every observation (what Summary::initial_cost reports).
double oracle_lba_cost(int C, int L, int N, const int* cam_idx, const int* line_idx, const double* obs,
                       int robust, double huber_a, double baseline, const double* params) {
  (void)L;
  double total = 0.0;
  for (int i = 0; i < N; ++i) {
    double res[4];
    line_reprojection_error<double>(params + 6 * cam_idx[i], params + 6 * C + 4 * line_idx[i], obs + 8 * i, baseline, res);
    const double s = res[0] * res[0] + res[1] * res[1] + res[2] * res[2] + res[3] * res[3];
    double rho, w; huber(s, huber_a, robust != 0, &rho, &w);
    total += 0.5 * rho;
  }
  return total;
}

// What ceres::Solve does for an LBAProblem (reference slam.cpp:663, 944).  solver: 0 full normal
// equations (reference behaviour, lba_problem.cpp:96-101), 1 Schur on lines (same step).
// lm_opts may be NULL (Ceres 1.7.0 defaults) or {function_tol, gradient_tol, parameter_tol, initial_radius}.
int oracle_lba_solve(int C, int L, int N, int max_iters, const int* cam_idx, const int* line_idx,
                     const int* fixed_idx, const double* obs, int robust, double huber_a, double baseline,
                     int solver, const double* lm_opts, double* params, double* summary8, double* trace) {
  LBAProgram p;
  p.C = C; p.L = L; p.N = N; p.cam_idx = cam_idx; p.line_idx = line_idx; p.obs = obs;
  p.robust = robust != 0; p.huber_a = huber_a; p.baseline = baseline; p.solver = solver; p.params = params;
  p.setup(fixed_idx);
  LMOptions o; o.max_iterations = max_iters;
  if (lm_opts) { o.function_tolerance = lm_opts[0]; o.gradient_tolerance = lm_opts[1]; o.parameter_tolerance = lm_opts[2]; o.initial_radius = lm_opts[3]; }
  LMSummary s;
  levenberg_marquardt(p, o, &s, trace);
  write_summary(s, summary8);
  return 0;
}

double oracle_po_cost(int K, int E, const int* idx1, const int* idx2, const double* cons, const double* params) {
  (void)K;
  double total = 0.0;
  for (int e = 0; e < E; ++e) {
    double res[6];
    pose_constraint_error<double>(params + 6 * idx1[e], params + 6 * idx2[e], cons + 6 * e, res);
    for (int k = 0; k < 6; ++k) total += 0.5 * res[k] * res[k];
  }
  return total;
}

// What ceres::Solve does for a POProblem (reference slam.cpp:1283-1293).  solver 1: sparse Cholesky of the normal equations
// (the reference's SPARSE_NORMAL_CHOLESKY); 0: dense Cholesky (cross-check of the sparse code, same step).
// stats (optional, 2 doubles): flops of the last numeric factorisation, non-zeros of L.
int oracle_po_solve2(int K, int E, int max_iters, const int* idx1, const int* idx2, const double* cons,
                     const double* lm_opts, int solver, double* params, double* summary8, double* trace, double* stats) {
  POProgram p;
  p.K = K; p.E = E; p.idx1 = idx1; p.idx2 = idx2; p.cons = cons; p.params = params; p.solver = solver;
  p.setup();
  LMOptions o; o.max_iterations = max_iters;
  if (lm_opts) { o.function_tolerance = lm_opts[0]; o.gradient_tolerance = lm_opts[1]; o.parameter_tolerance = lm_opts[2]; o.initial_radius = lm_opts[3]; }
  LMSummary s;
  levenberg_marquardt(p, o, &s, trace);
  write_summary(s, summary8);
  if (stats) { stats[0] = (double)p.last_factor_flops(); stats[1] = (double)p.chol.Lx.size(); }
  return 0;
}

int oracle_po_solve(int K, int E, int max_iters, const int* idx1, const int* idx2, const double* cons,
                    const double* lm_opts, double* params, double* summary8, double* trace) {
  POProgram p;
  p.K = K; p.E = E; p.idx1 = idx1; p.idx2 = idx2; p.cons = cons; p.params = params; p.solver = 0;
  p.setup();
  LMOptions o; o.max_iterations = max_iters;
  if (lm_opts) { o.function_tolerance = lm_opts[0]; o.gradient_tolerance = lm_opts[1]; o.parameter_tolerance = lm_opts[2]; o.initial_radius = lm_opts[3]; }
  LMSummary s;
  levenberg_marquardt(p, o, &s, trace);
  write_summary(s, summary8);
  return 0;
}

// Scores n_hyp motion hypotheses (R row-major 9 | t 3 per hypothesis) against n_lines lines: the inner loops of
// SLAM::ransac_motion (reference src/slam.cpp:398-412).  scores[h] = number of inliers, or -1 when the hypothesis is
// skipped (|t| > 1, :400-401); inlier[h][k] = 1 when error < thr.  errors (optional) [n_hyp][n_lines] floats.
void oracle_ransac_score(int n_hyp, const double* poses, int n_lines, const double* lines, const double* obs,
                         double baseline, double thr, int* scores, unsigned char* inlier, float* errors) {
  for (int h = 0; h < n_hyp; ++h) {
    const double* R = poses + 12 * (size_t)h;
    const double* t = R + 9;
    if (std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) > 1) {
      scores[h] = -1;
      for (int k = 0; k < n_lines; ++k) { inlier[(size_t)h * n_lines + k] = 0; if (errors) errors[(size_t)h * n_lines + k] = 0.f; }
      continue;
    }
    int score = 0;
    for (int k = 0; k < n_lines; ++k) {
      const float e = ransac_reprojection_error(obs + 8 * (size_t)k, R, t, lines + 6 * (size_t)k, baseline);
      if (errors) errors[(size_t)h * n_lines + k] = e;
      const bool in = e < thr;
      inlier[(size_t)h * n_lines + k] = in ? 1 : 0;
      score += in ? 1 : 0;
    }
    scores[h] = score;
  }
}

}  // extern "C"
