// RANSAC hypothesis scoring on device: the inner loops of SLAM::ransac_motion (reference
// src/slam.cpp:398-412) -- every motion hypothesis against every common line through SLAM::reprojection_error
// (src/slam.cpp:691-726) -- as one launch: thread = (hypothesis, line).  The result is an integer score and an inlier
// mask per hypothesis, so parity with the oracle is BIT-EXACT: the reference's mixed precision is kept (`error` and the
// normaliser `sql` are float, the geometry double) and every product / sum is an explicit round-to-nearest intrinsic,
// so the compiler cannot contract them into FMAs the CPU restatement does not perform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slslam {

__device__ __forceinline__ double rs_dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

// mean absolute endpoint-to-line distance over the four stereo endpoints, in normalised image units
__device__ __forceinline__ float ransac_reprojection_error(const double* __restrict__ ft, const double* __restrict__ R,
                                                           const double* __restrict__ t_in, const double* __restrict__ line,
                                                           double baseline) {
  float error = 0.f;
  double t0 = t_in[0];
  const double t1 = t_in[1], t2 = t_in[2];
  double dvc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) dvc[r] = rs_dot3(R[3 * r], R[3 * r + 1], R[3 * r + 2], line[3], line[4], line[5]);
  double rc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) rc[r] = rs_dot3(R[3 * r], R[3 * r + 1], R[3 * r + 2], line[0], line[1], line[2]);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (i == 1) t0 = __dsub_rn(t0, baseline);
    const double c0 = __dadd_rn(rc[0], t0), c1 = __dadd_rn(rc[1], t1), c2 = __dadd_rn(rc[2], t2);
    double n0 = __dsub_rn(__dmul_rn(c1, dvc[2]), __dmul_rn(c2, dvc[1]));
    double n1 = __dsub_rn(__dmul_rn(c2, dvc[0]), __dmul_rn(c0, dvc[2]));
    double n2 = __dsub_rn(__dmul_rn(c0, dvc[1]), __dmul_rn(c1, dvc[0]));
    const float sql = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1))));
    const double sd = (double)sql;
    n0 = __ddiv_rn(n0, sd); n1 = __ddiv_rn(n1, sd); n2 = __ddiv_rn(n2, sd);
    const double* p = ft + 4 * i;
    const double e1 = fabs(__dadd_rn(__dadd_rn(__dmul_rn(n0, p[0]), __dmul_rn(n1, p[1])), n2));
    const double e2 = fabs(__dadd_rn(__dadd_rn(__dmul_rn(n0, p[2]), __dmul_rn(n1, p[3])), n2));
    error = __double2float_rn(__dadd_rn((double)error, e1));
    error = __double2float_rn(__dadd_rn((double)error, e2));
  }
  return __double2float_rn(__ddiv_rn((double)error, 4.0));
}

// grid (ceil(n_lines / 256), n_hyp).  scores must be zeroed before the launch; integer atomics only (order-free).
__global__ void __launch_bounds__(256) ransac_score_kernel(int n_hyp, const double* __restrict__ poses, int n_lines,
                                                           const double* __restrict__ lines, const double* __restrict__ obs,
                                                           double baseline, double thr, int* __restrict__ scores,
                                                           unsigned char* __restrict__ inlier, float* __restrict__ errors) {
  const int h = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double sp[12];
  if (threadIdx.x < 12) sp[threadIdx.x] = poses[12 * (size_t)h + threadIdx.x];
  __syncthreads();
  // |t| > 1: the hypothesis is skipped (src/slam.cpp:400-401)
  const bool skip = __dsqrt_rn(rs_dot3(sp[9], sp[10], sp[11], sp[9], sp[10], sp[11])) > 1.0;
  bool in = false;
  float e = 0.f;
  if (k < n_lines && !skip) {
    double ft[8], ln[6];
#pragma unroll
    for (int q = 0; q < 8; ++q) ft[q] = obs[8 * (size_t)k + q];
#pragma unroll
    for (int q = 0; q < 6; ++q) ln[q] = lines[6 * (size_t)k + q];
    e = ransac_reprojection_error(ft, sp, sp + 9, ln, baseline);
    in = (double)e < thr;
  }
  if (k < n_lines) {
    if (inlier) inlier[(size_t)h * n_lines + k] = in ? 1 : 0;
    if (errors) errors[(size_t)h * n_lines + k] = e;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0) {
    if (skip) { if (blockIdx.x == 0 && threadIdx.x == 0) scores[h] = -1; }
    else if (bal) atomicAdd(&scores[h], __popc(bal));
  }
}

}  // namespace slslam
