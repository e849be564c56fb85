// Motion-only bundle adjustment on device: ONE free camera, every line constant -- what SLAM::motion_only_ba packs for
// every frame (reference src/slam.cpp:578-675: camera 0 = the pose being refined, camera 1 = identity and constant,
// every line constant, two observations per line).  The problem is the same LBAProblem handed to the same ceres::Solve,
// so the trust-region loop is the one of lba_kernel.cuh with the Schur machinery gone: the normal
// equations are a single 6x6 block.  One CTA per problem, the whole LM loop in one launch; a batch of frames is one
// grid.  Per LM iteration: residual + analytic camera Jacobian per observation of the free camera (thread = observation),
// Huber corrector, Jacobi scaling, 27 sums (H lower triangle + gradient) by shuffle trees in a fixed order, the 6x6
// Cholesky redundantly in every thread, the trial-cost sweep, the accept / reject logic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slslam_b200.h"
#include "lba_math.cuh"

namespace slslam {

constexpr int MOBA_NT = 256;
constexpr int MOBA_NW = MOBA_NT / 32;
constexpr int MOBA_MAX_FREE_OBS = 1024;   // observations of the free camera staged in shared memory (25 doubles each)
constexpr int MOBA_OBS_STRIDE = 25;       // ob[8] | xh yh zh xb (12) | d ist2 s1 | pad(2) -> odd stride, conflict-free
constexpr int MOBA_FIXED_DOUBLES = 12 + 2 * CAM_STRIDE + 6 + (MOBA_NW + 1) * 28 + 2;   // shared memory before the staged observations

struct MobaHdr {
  int C, L, N, free_cam, max_iters, robust;
  double huber_a, baseline, ftol, gtol, ptol, radius0;
  const int* cam_idx;        // [N]
  const int* line_idx;       // [N]
  const double* obs;         // [N][8]
  const double* params_in;   // [6C + 4L]
  double* params_out;        // [6C + 4L]
  slslam_summary* summary;
  double* trace;             // [max_iters][SLSLAM_TRACE_WIDTH] or nullptr
};

__device__ __forceinline__ double moba_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA sum of NV per-thread values in a fixed order (lanes by butterfly, warps 0..7 in order); every thread gets the sums.
// wsc: [MOBA_NW][NV] warp partials, then [NV] totals.
template <int NV>
__device__ __forceinline__ void moba_cta_sum(double* v, double* wsc, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  double* tot = wsc + MOBA_NW * 28;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = moba_warp_sum(v[k]);
    if (lane == 0) wsc[warp * NV + k] = s;
  }
  __syncthreads();
  if (tid < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MOBA_NW; ++w) s += wsc[w * NV + tid];
    tot[tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = tot[k];
}

// The 28-value sum of the linearisation: a transposing warp reduction (31 shuffle-adds for 32 values, lane l ends up
// with the warp total of value l) instead of 28 butterflies of 5, then the same fixed-order combine over the warps.
__device__ __forceinline__ void moba_cta_sum28(double* a, double* wsc, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  double* tot = wsc + MOBA_NW * 28;
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = k < 28 ? a[k] : 0.0;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const double send = upper ? v[k] : v[k + half];
      const double keep = upper ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  if (lane < 28) wsc[warp * 28 + lane] = v[0];
  __syncthreads();
  if (tid < 28) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MOBA_NW; ++w) s += wsc[w * 28 + tid];
    tot[tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 28; ++k) a[k] = tot[k];
}

__device__ __forceinline__ void moba_load_trig(const double* __restrict__ so, LineTrig& lt) {
#pragma unroll
  for (int k = 0; k < 3; ++k) { lt.xh[k] = so[8 + k]; lt.yh[k] = so[11 + k]; lt.zh[k] = so[14 + k]; lt.xb[k] = so[17 + k]; }
  lt.d = so[20]; lt.ist2 = so[21]; lt.s1 = so[22];
}

__global__ void __launch_bounds__(MOBA_NT, 1) lba_motion_only_kernel(const MobaHdr* __restrict__ hdrs) {
  extern __shared__ __align__(16) double sm[];
  const MobaHdr& h = hdrs[blockIdx.x];
  const int tid = threadIdx.x;
  // shared layout
  double* camx = sm;                       // [6]  the free camera at x
  double* camxt = sm + 6;                  // [6]  at the trial point
  double* camR = sm + 12;                  // [CAM_STRIDE] R, dR/dw, t at x
  double* camRt = camR + CAM_STRIDE;       // [CAM_STRIDE] at the trial point
  double* cscale = camRt + CAM_STRIDE;     // [6] Jacobi scale
  double* wsc = cscale + 6;                // [MOBA_NW][28] + [28] reduction scratch
  int* nfree_s = reinterpret_cast<int*>(wsc + (MOBA_NW + 1) * 28);
  double* fobs = sm + MOBA_FIXED_DOUBLES;  // [nfree][MOBA_OBS_STRIDE]
  const int C = h.C, N = h.N, fc = h.free_cam;
  const double* lines = h.params_in + 6 * C;
  const bool robust = h.robust != 0;

  if (tid < 6) camx[tid] = h.params_in[6 * fc + tid];
  if (tid == 0) *nfree_s = 0;
  __syncthreads();
  if (tid == 0) cam_precompute(camx, camR, true);

  // ---- one pass over the caller's observations: those of the free camera are staged (with the trigonometry of their
  // constant line), the others only contribute their (constant) cost.  Slots are assigned in observation order by a
  // ballot scan, so the summation order -- and with it every bit of the result -- is fixed. ----
  double fixed_cost = 0.0;
  for (int base = 0; base < N; base += MOBA_NT) {
    const int i = base + tid;
    const bool in = i < N;
    const int cam = in ? h.cam_idx[i] : -1;
    const bool is_free = in && cam == fc;
    // position among the free observations: warp ballot + per-warp counts through shared memory
    const unsigned bal = __ballot_sync(0xffffffffu, is_free);
    int* wcnt = reinterpret_cast<int*>(wsc);
    if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = *nfree_s;
    for (int w = 0; w < (tid >> 5); ++w) off += wcnt[w];
    const int slot = off + __popc(bal & ((1u << (tid & 31)) - 1u));
    int total = 0;
    for (int w = 0; w < MOBA_NW; ++w) total += wcnt[w];
    __syncthreads();
    if (tid == 0) *nfree_s += total;
    if (in) {
      LineTrig lt;
      line_trig(lines + 4 * h.line_idx[i], lt);
      const double* ob = h.obs + 8 * (size_t)i;
      if (is_free) {
        double* so = fobs + (size_t)slot * MOBA_OBS_STRIDE;
#pragma unroll
        for (int k = 0; k < 8; ++k) so[k] = ob[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) { so[8 + k] = lt.xh[k]; so[11 + k] = lt.yh[k]; so[14 + k] = lt.zh[k]; so[17 + k] = lt.xb[k]; }
        so[20] = lt.d; so[21] = lt.ist2; so[22] = lt.s1;
      } else {
        double cpre[CAM_STRIDE], o8[8], r[4], w;
        cam_precompute(h.params_in + 6 * cam, cpre, false);
#pragma unroll
        for (int k = 0; k < 8; ++k) o8[k] = ob[k];
        obs_eval<false>(cpre, lt, o8, h.baseline, r, nullptr, nullptr);
        fixed_cost += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
      }
    }
    __syncthreads();
  }
  const int nfree = *nfree_s;

  // ---- Jacobi scaling from the column norms at x0; initial cost ----
  double cost;
  {
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 6 squared column norms | cost | fixed cost
    for (int i = tid; i < nfree; i += MOBA_NT) {
      const double* so = fobs + (size_t)i * MOBA_OBS_STRIDE;
      LineTrig lt; moba_load_trig(so, lt);
      double r[4], Jc[24], Jl[16], w;
      obs_eval<true>(camR, lt, so, h.baseline, r, Jc, Jl);
      v[6] += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] += w * w * (Jc[j] * Jc[j] + Jc[6 + j] * Jc[6 + j] + Jc[12 + j] * Jc[12 + j] + Jc[18 + j] * Jc[18 + j]);
    }
    v[7] = fixed_cost;
    moba_cta_sum<8>(v, wsc, tid);
    if (tid < 6) cscale[tid] = 1.0 / (1.0 + sqrt(tid == 0 ? v[0] : tid == 1 ? v[1] : tid == 2 ? v[2] : tid == 3 ? v[3] : tid == 4 ? v[4] : v[5]));
    cost = v[6]; fixed_cost = v[7];
    __syncthreads();
  }
  const double initial_cost = cost + fixed_cost;

  double radius = h.radius0, decrease_factor = 2.0, gmax = 0.0, gtol_abs = 0.0;
  int successful = 0, unsuccessful = 0, invalid = 0, term = SLSLAM_NO_CONVERGENCE, iters = 0;
  bool first_lin = true;
  // one pass more than max_iters when the last allowed iteration accepted its step: Ceres evaluates the gradient (and runs
  // the gradient test) right after every accepted step, so cost / gradient_max_norm / termination_type at the iteration
  // cap come from the final point
  bool grad_pending = false;
  for (int it = 0; it <= h.max_iters; ++it) {
    const bool last = it == h.max_iters;
    if (last && !grad_pending) break;
    // -- linearise at x: H (21, lower) | g (6) | cost --
    double a[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) a[k] = 0.0;
    for (int i = tid; i < nfree; i += MOBA_NT) {
      const double* so = fobs + (size_t)i * MOBA_OBS_STRIDE;
      LineTrig lt; moba_load_trig(so, lt);
      double r[4], Jc[24], Jl[16], w;
      obs_eval<true>(camR, lt, so, h.baseline, r, Jc, Jl);
      a[27] += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[k] *= w;
#pragma unroll
        for (int j = 0; j < 6; ++j) Jc[6 * k + j] *= w * cscale[j];
      }
#pragma unroll
      for (int p = 0; p < 6; ++p) {
#pragma unroll
        for (int q = 0; q <= p; ++q) a[p * (p + 1) / 2 + q] += Jc[p] * Jc[q] + Jc[6 + p] * Jc[6 + q] + Jc[12 + p] * Jc[12 + q] + Jc[18 + p] * Jc[18 + q];
        a[21 + p] += Jc[p] * r[0] + Jc[6 + p] * r[1] + Jc[12 + p] * r[2] + Jc[18 + p] * r[3];
      }
    }
    moba_cta_sum28(a, wsc, tid);
    cost = a[27];
    // gradient max norm with the unscaled Jacobian; |x|^2 of the free camera
    double x_norm2 = 0.0;
    gmax = 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) { gmax = fmax(gmax, fabs(a[21 + p] / cscale[p])); x_norm2 += camx[p] * camx[p]; }
    if (first_lin) { gtol_abs = h.gtol * fmax(gmax, 2.220446049250313e-16); first_lin = false; }
    grad_pending = false;
    if (gmax <= gtol_abs) { term = SLSLAM_GRADIENT_TOLERANCE; break; }
    if (last) break;
    iters = it + 1;
    double* tr = (h.trace && tid == 0) ? h.trace + (size_t)it * SLSLAM_TRACE_WIDTH : nullptr;
    if (tr) { tr[0] = cost; tr[1] = 0; tr[2] = 0; tr[3] = radius; tr[4] = 0; tr[5] = 0; tr[6] = gmax; tr[7] = 0; }
    // -- (H + D) y = g by Cholesky, redundantly in every thread; step = -y * scale --
    double D[6], Lm[21], y[6];
    bool ok = true;
#pragma unroll
    for (int p = 0; p < 6; ++p) D[p] = fmin(fmax(a[p * (p + 1) / 2 + p], 1e-6), 1e32) / radius;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double d = a[k * (k + 1) / 2 + k] + D[k];
#pragma unroll
      for (int m = 0; m < k; ++m) d -= Lm[k * (k + 1) / 2 + m] * Lm[k * (k + 1) / 2 + m];
      ok = ok && (d > 0.0);
      const double inv = pivot_rsqrt(d);
      Lm[k * (k + 1) / 2 + k] = inv;                      // the diagonal holds 1 / l_kk
#pragma unroll
      for (int p = k + 1; p < 6; ++p) {
        double s = a[p * (p + 1) / 2 + k];
#pragma unroll
        for (int m = 0; m < k; ++m) s -= Lm[p * (p + 1) / 2 + m] * Lm[k * (k + 1) / 2 + m];
        Lm[p * (p + 1) / 2 + k] = s * inv;
      }
    }
#pragma unroll
    for (int p = 0; p < 6; ++p) {                         // L z = g
      double s = a[21 + p];
#pragma unroll
      for (int m = 0; m < p; ++m) s -= Lm[p * (p + 1) / 2 + m] * y[m];
      y[p] = s * Lm[p * (p + 1) / 2 + p];
    }
#pragma unroll
    for (int p = 5; p >= 0; --p) {                        // L^T y = z
      double s = y[p];
#pragma unroll
      for (int m = p + 1; m < 6; ++m) s -= Lm[m * (m + 1) / 2 + p] * y[m];
      y[p] = s * Lm[p * (p + 1) / 2 + p];
    }
    double model = 0.0, dn2 = 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      model += 0.5 * y[p] * (a[21 + p] + D[p] * y[p]);
      const double d = y[p] * cscale[p];
      dn2 += d * d;
      if (!isfinite(y[p])) ok = false;
    }
    if (tr) tr[2] = model;
    double new_cost = 0.0;
    if (ok && !(model < 0.0)) {
      __syncthreads();
      if (tid < 6) camxt[tid] = camx[tid] - (tid == 0 ? y[0] : tid == 1 ? y[1] : tid == 2 ? y[2] : tid == 3 ? y[3] : tid == 4 ? y[4] : y[5]) * cscale[tid];
      __syncthreads();
      if (tid == 0) cam_precompute(camxt, camRt, true);     // derivatives too: an accepted step adopts the block by a copy
      __syncthreads();
      double v[1] = {0.0};
      for (int i = tid; i < nfree; i += MOBA_NT) {
        const double* so = fobs + (size_t)i * MOBA_OBS_STRIDE;
        LineTrig lt; moba_load_trig(so, lt);
        double r[4], w;
        obs_eval<false>(camRt, lt, so, h.baseline, r, nullptr, nullptr);
        v[0] += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
      }
      moba_cta_sum<1>(v, wsc, tid);
      new_cost = v[0];
    }
    if (!ok || model < 0.0) {
      ++unsuccessful;
      if (tr) tr[5] = -1.0;
      if (++invalid >= 5) { term = SLSLAM_NUMERICAL_FAILURE; break; }
      radius *= 0.5;
      if (radius < 1e-32) { term = SLSLAM_PARAMETER_TOLERANCE; break; }
      continue;
    }
    invalid = 0;
    const double step_norm = sqrt(dn2), x_norm = sqrt(x_norm2);
    if (tr) { tr[1] = new_cost; tr[4] = step_norm; }
    if (step_norm <= h.ptol * (x_norm + h.ptol)) { term = SLSLAM_PARAMETER_TOLERANCE; break; }
    const double cost_change = cost - new_cost;
    if (fabs(cost_change) < h.ftol * cost) { term = SLSLAM_FUNCTION_TOLERANCE; break; }
    const double rel = cost_change / model;
    if (tr) tr[7] = rel;
    if (rel > 1e-3) {
      ++successful;
      if (tr) tr[5] = 1.0;
      __syncthreads();
      if (tid < 6) camx[tid] = camxt[tid];
      if (tid < CAM_STRIDE) camR[tid] = camRt[tid];
      __syncthreads();
      cost = new_cost;
      grad_pending = true;
      const double t = 2.0 * rel - 1.0;
      radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0;
    } else {
      ++unsuccessful;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
    if (radius < 1e-32) { term = SLSLAM_PARAMETER_TOLERANCE; break; }
  }
  __syncthreads();
  // write back: everything but the free camera is constant
  const int np = 6 * C + 4 * h.L;
  for (int i = tid; i < np; i += MOBA_NT) h.params_out[i] = (i >= 6 * fc && i < 6 * fc + 6) ? camx[i - 6 * fc] : h.params_in[i];
  if (tid == 0) {
    slslam_summary s;
    s.initial_cost = initial_cost; s.final_cost = cost + fixed_cost; s.fixed_cost = fixed_cost; s.gradient_max_norm = gmax;
    s.num_successful_steps = successful; s.num_unsuccessful_steps = unsuccessful; s.termination_type = term; s.iterations = iters;
    *h.summary = s;
  }
}

}  // namespace slslam
