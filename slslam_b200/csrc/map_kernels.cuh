// Device-resident SLAM map and the window assembly / write-back around an LBA solve: what
// SLAM::bundle_adjustment does on the host before and after ceres::Solve (reference src/slam.cpp:799-920, 957-972) with
// the geometry helpers it calls (src/gc.cpp: gc_Rt_to_wt / gc_wt_to_Rt through ceres/rotation.h :24-50,
// gc_line_to_pose / gc_line_from_pose :63-81, gc_av_to_orth :361-417, gc_orth_to_av :419-460).
//
// Resident state: keyframe poses T = (R row-major 9 | t 3), world(window) -> camera (keyframe_t::T); landmark lines as
// (closest point, direction) in the frame of their initial keyframe (landmark_t::line, init_kfid); the observations of
// every keyframe, appended once when the keyframe is created (landmark id, 8 normalised stereo endpoint coordinates).
// Per keyframe only its new observations (and the re-anchored poses of the <= 2W window keyframes, < 4 KB) go up.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slslam_b200.h"

namespace slslam {

struct MapDev {
  int max_kf, max_lm, max_obs;
  double* kf_T;            // [max_kf][12]
  int* kf_cam;             // [max_kf] camera index of the keyframe in the current window or -1
  int* kf_free;            // [max_kf] 1 when the keyframe is a free camera of the current window
  double* lm_line;         // [max_lm][6] closest point | direction, in the frame of lm_init_kf
  int* lm_init_kf;         // [max_lm] or -1 when the landmark id is unused
  int* lm_count;           // [max_lm] observations in free keyframes of the current window
  int* lm_line_index;      // [max_lm] line index in the current window or -1
  int* obs_lm;             // [max_obs]
  int* obs_kf;             // [max_obs]
  double* obs_xy;          // [max_obs][8]
  // window assembly
  const int2* ranges;      // [n_ranges] (first observation, count) of every window keyframe, chronological
  int n_ranges, n_cand;    // candidates = sum of counts
  const int* cand_off;     // [n_ranges + 1] exclusive prefix of the counts
  int* keep;               // [n_cand] 1 when the candidate enters the window
  int* keep_pos;           // [n_cand] exclusive prefix of keep
  int* sizes;              // [4] L, N, error flags
  // the window, in the reference's array layout (LBAProblem, src/lba_problem.h:188-196)
  int* w_cam; int* w_line; int* w_fixed; double* w_obs; double* w_params;
  int* w_line_lm;          // [L] landmark of every window line
  int C, Cfree;
};

// ---- rotation helpers (ceres/rotation.h semantics, restated in include/ceres/rotation.h for the host) ----
// R row-major here (pose_t::R as Eigen prints it); Ceres works on the column-major array of the same matrix.
__device__ inline void map_R_to_w(const double* R, double* w) {
  w[0] = R[7] - R[5];                      // R21 - R12
  w[1] = R[2] - R[6];                      // R02 - R20
  w[2] = R[3] - R[1];                      // R10 - R01
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
  double s = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) * 0.5;
  s = s > 1.0 ? 1.0 : s;
  const double th = atan2(s, c);
  if (s > 1e-12) {
    const double k = th / (2.0 * s);
    w[0] *= k; w[1] *= k; w[2] *= k;
    return;
  }
  if (c > 0.0) { w[0] *= 0.5; w[1] *= 0.5; w[2] *= 0.5; return; }
  const double inv = 1.0 / (1.0 - c);
  for (int i = 0; i < 3; ++i) {
    double a = (R[4 * i] - c) * inv;
    a = a < 0.0 ? 0.0 : a;
    double v = th * sqrt(a);
    if (w[i] < 0.0) v = -v;
    w[i] = v;
  }
}

__device__ inline void map_w_to_R(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2);
    const double x = w[0] / th, y = w[1] / th, z = w[2] / th;
    double s, c;
    sincos(th, &s, &c);
    const double v = 1.0 - c;
    R[0] = c + x * x * v;      R[1] = x * y * v - z * s;  R[2] = x * z * v + y * s;
    R[3] = y * x * v + z * s;  R[4] = c + y * y * v;      R[5] = y * z * v - x * s;
    R[6] = z * x * v - y * s;  R[7] = z * y * v + x * s;  R[8] = c + z * z * v;
  } else {
    R[0] = 1.0;   R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2];  R[4] = 1.0;   R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0];  R[8] = 1.0;
  }
}

// gc_av_to_orth (src/gc.cpp:361-417): (closest point a, direction v) -> (alpha, beta, gamma, theta)
__device__ inline void map_av_to_orth(const double* av, double* o) {
  const double* a = av; const double* v = av + 3;
  const double n[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
  const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double x[3] = {n[0] / nn, n[1] / nn, n[2] / nn}, y[3] = {v[0] / vn, v[1] / vn, v[2] / vn};
  const double z2 = x[0] * y[1] - x[1] * y[0];
  o[0] = atan2(y[2], z2);
  o[1] = asin(-x[2]);
  o[2] = atan2(x[1], x[0]);
  const double wn = sqrt(nn * nn + vn * vn);
  o[3] = asin(vn / wn);
}

// gc_orth_to_av (src/gc.cpp:419-460)
__device__ inline void map_orth_to_av(const double* o, double* av) {
  double s1, c1, s2, c2, s3, c3, st, ct;
  sincos(o[0], &s1, &c1); sincos(o[1], &s2, &c2); sincos(o[2], &s3, &c3); sincos(o[3], &st, &ct);
  const double d = ct / st;
  av[0] = -(c1 * s2 * c3 + s1 * s3) * d;
  av[1] = -(c1 * s2 * s3 - s1 * c3) * d;
  av[2] = -(c1 * c2) * d;
  av[3] = s1 * s2 * c3 - c1 * s3;
  av[4] = s1 * s2 * s3 + c1 * c3;
  av[5] = s1 * c2;
}

// gc_line_to_pose (src/gc.cpp:63-77): line in frame 0 -> frame 1, x1 = R x0 + t
__device__ inline void map_line_to_pose(const double* l, const double* T, double* out) {
  for (int r = 0; r < 3; ++r) {
    out[r] = T[3 * r] * l[0] + T[3 * r + 1] * l[1] + T[3 * r + 2] * l[2] + T[9 + r];
    out[3 + r] = T[3 * r] * l[3] + T[3 * r + 1] * l[4] + T[3 * r + 2] * l[5];
  }
}
// gc_line_from_pose (src/gc.cpp:79-81): the inverse transform (R^-1 = R^T for the rotations the map holds)
__device__ inline void map_line_from_pose(const double* l, const double* T, double* out) {
  const double p[3] = {l[0] - T[9], l[1] - T[10], l[2] - T[11]};
  for (int r = 0; r < 3; ++r) {
    out[r] = T[r] * p[0] + T[3 + r] * p[1] + T[6 + r] * p[2];
    out[3 + r] = T[r] * l[3] + T[3 + r] * l[4] + T[6 + r] * l[5];
  }
}

// ---- batch conversions (test / utility entry points) ----
__global__ void map_convert_kernel(int n, int mode, const double* in, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) map_av_to_orth(in + 6 * (size_t)i, out + 4 * (size_t)i);
  else if (mode == 1) map_orth_to_av(in + 4 * (size_t)i, out + 6 * (size_t)i);
  else if (mode == 2) map_R_to_w(in + 9 * (size_t)i, out + 3 * (size_t)i);
  else map_w_to_R(in + 3 * (size_t)i, out + 9 * (size_t)i);
}

// ---- window assembly ----
// candidate index -> (range, position) by binary search in the prefix of the range sizes
__device__ inline int map_cand_obs(const MapDev& m, int i) {
  int lo = 0, hi = m.n_ranges;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (m.cand_off[mid] <= i) lo = mid; else hi = mid; }
  return m.ranges[lo].x + (i - m.cand_off[lo]);
}

// 0: the window's cameras: keyframe -> camera index, free flag
__global__ void map_mark_kernel(MapDev m, const int* cam_kf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.C) return;
  m.kf_cam[cam_kf[i]] = i;
  m.kf_free[cam_kf[i]] = i < m.Cfree ? 1 : 0;
}

// 1: per landmark, the number of its observations in FREE keyframes of the window (slam.cpp:819-829 counts the free
// keyframes' member landmarks)
__global__ void map_count_kernel(MapDev m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n_cand) return;
  const int o = map_cand_obs(m, i);
  const int lm = m.obs_lm[o];
  if (m.kf_free[m.obs_kf[o]] && lm >= 0 && lm < m.max_lm && m.lm_init_kf[lm] >= 0) atomicAdd(m.lm_count + lm, 1);
}

// Exclusive scan of `flag(i)` over [0, n) by ONE CTA (n is a few ten thousand at most), result to out, total returned to
// every thread.
template <class F>
__device__ int map_block_scan(int n, F flag, int* out, int* sh /* [blockDim + 1] */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt, b = tid * per, e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += flag(i);
  sh[tid] = s;
  __syncthreads();
  if (tid == 0) { int run = 0; for (int k = 0; k < nt; ++k) { const int v = sh[k]; sh[k] = run; run += v; } sh[nt] = run; }
  __syncthreads();
  int run = sh[tid];
  for (int i = b; i < e; ++i) { out[i] = run; run += flag(i); }
  const int total = sh[nt];
  __syncthreads();
  return total;
}

// 2: landmarks seen by at least two free keyframes become the window's lines, in landmark-id order (slam.cpp:838-845);
//    candidates of those landmarks become its observations, in chronological order (their order inside a line is the
//    landmark's obs_vec order, slam.cpp:848-881)
__global__ void __launch_bounds__(1024) map_select_kernel(MapDev m) {
  __shared__ int sh[1025];
  const int L = map_block_scan(m.max_lm, [&](int l) { return m.lm_count[l] >= 2 ? 1 : 0; }, m.lm_line_index, sh);
  for (int l = threadIdx.x; l < m.max_lm; l += blockDim.x) {
    if (m.lm_count[l] >= 2) m.w_line_lm[m.lm_line_index[l]] = l; else m.lm_line_index[l] = -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m.n_cand; i += blockDim.x) {
    const int o = map_cand_obs(m, i);
    const int lm = m.obs_lm[o];
    m.keep[i] = (lm >= 0 && lm < m.max_lm && m.lm_line_index[lm] >= 0) ? 1 : 0;
  }
  __syncthreads();
  const int N = map_block_scan(m.n_cand, [&](int i) { return m.keep[i]; }, m.keep_pos, sh);
  if (threadIdx.x == 0) { m.sizes[0] = L; m.sizes[1] = N; }
}

// 3: the window's arrays.  fixed_index[2i] = 1 for a keyframe beyond the first W by graph distance (slam.cpp:855-871),
//    fixed_index[2i+1] = 0 (:907)
__global__ void map_emit_kernel(MapDev m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.n_cand || !m.keep[i]) return;
  const int o = map_cand_obs(m, i), p = m.keep_pos[i];
  const int kf = m.obs_kf[o];
  m.w_cam[p] = m.kf_cam[kf];
  m.w_line[p] = m.lm_line_index[m.obs_lm[o]];
  m.w_fixed[2 * p] = m.kf_free[kf] ? 0 : 1;
  m.w_fixed[2 * p + 1] = 0;
  const double2* src = reinterpret_cast<const double2*>(m.obs_xy + 8 * (size_t)o);
  double2* dst = reinterpret_cast<double2*>(m.w_obs + 8 * (size_t)p);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = src[k];
}

// 4: parameters: cameras gc_Rt_to_wt(kf->T) (slam.cpp:830, 862), lines gc_av_to_orth(gc_line_from_pose(line, T_init))
//    (slam.cpp:883-885)
__global__ void map_params_kernel(MapDev m, const int* cam_kf, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m.C) {
    const double* T = m.kf_T + 12 * (size_t)cam_kf[i];
    double* o = m.w_params + 6 * (size_t)i;
    map_R_to_w(T, o);
    o[3] = T[9]; o[4] = T[10]; o[5] = T[11];
  } else if (i < m.C + L) {
    const int l = i - m.C, lm = m.w_line_lm[l];
    double av[6];
    map_line_from_pose(m.lm_line + 6 * (size_t)lm, m.kf_T + 12 * (size_t)m.lm_init_kf[lm], av);
    map_av_to_orth(av, m.w_params + 6 * (size_t)m.C + 4 * (size_t)l);
  }
}

// 5: write-back (slam.cpp:957-972): every camera of the window gets T = gc_wt_to_Rt(pose), then every line
//    lm->line = gc_line_to_pose(gc_orth_to_av(line), T_init) with the UPDATED pose of its initial keyframe
__global__ void map_writeback_cams_kernel(MapDev m, const int* cam_kf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.C) return;
  const double* p = m.w_params + 6 * (size_t)i;
  double* T = m.kf_T + 12 * (size_t)cam_kf[i];
  map_w_to_R(p, T);
  T[9] = p[3]; T[10] = p[4]; T[11] = p[5];
}
__global__ void map_writeback_lines_kernel(MapDev m, int L) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const int lm = m.w_line_lm[l];
  double av[6];
  map_orth_to_av(m.w_params + 6 * (size_t)m.C + 4 * (size_t)l, av);
  map_line_to_pose(av, m.kf_T + 12 * (size_t)m.lm_init_kf[lm], m.lm_line + 6 * (size_t)lm);
}

// clears the per-window marks of the keyframes and landmarks the window touched
__global__ void map_reset_kernel(MapDev m, const int* cam_kf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m.C) { m.kf_cam[cam_kf[i]] = -1; m.kf_free[cam_kf[i]] = 0; }
  if (i < m.max_lm) { m.lm_count[i] = 0; }
}

}  // namespace slslam
