// LBA windows beyond the limits of the tiled solve kernel (lba_kernel.cuh: 32 cameras, 24 free cameras, 32 observations
// per line): the reference's --ba_window_size 20 / 40 runs (matlab_script/result_comp_ancdir_orthonorm/*basize{20,40}*,
// up to 2 W = 80 cameras, 40 of them free, and in the house simulation nearly every camera sees every line) build such
// windows (src/slam.cpp:1376-1382, 811-871).  Same algorithm -- Huber-corrected, Jacobi-scaled LM on the Schur
// complement of the line blocks, Ceres 1.7.0 control -- in a shape that has no such limits:
// a GROUP of G co-resident CTAs per window (cooperative launch; round 2: round 1 of this kernel ran one CTA per window and
// was bound by the latency of its own global loads), every intermediate in global memory (it stays in L2: a W = 40 house
// window is < 3 MB), any number of cameras and of observations per line, up to WIDE_MAX_FREE free cameras (reduced
// system 6 * 64 square).  Work items of a phase are strided over all threads (or warps) of the group; phases are
// separated by an arrive / spin barrier on an L2 counter (release / acquire, as in lba_kernel.cuh); scalars are reduced
// per CTA and then over the group in rank order.  All reductions have a fixed order: results are bit-reproducible.
//   linearise   thread per observation (line-sorted order): residual, analytic Jacobian, corrector, Jacobi scale -> r, Jc, Jl
//   lines       warp per line: H_ll, g_l over its observations (lane-strided partial sums, butterfly), 4x4 Cholesky,
//               Z_i = (Jc_i^T Jl_i) L^-T per observation
//   cameras     thread per (free camera, accumulator): H_cc - Z Z^T, g_c, Z u, diag H_cc over the camera's observations
//   pairs       thread per entry of every off-diagonal block (I > K): - sum over the lines seen by both of Z_I Z_K^T
//               (lookup table line x free camera -> observation)
//   solve       CTA 0 of the group: right-looking block elimination of the dense reduced system with explicit 6x6 pivot
//               inverses, the current panel in shared memory; back-substitution without solves
//   trial       thread per camera: trial pose and rotation table; warp per line: y_l, trial line and its sines / cosines;
//               thread per observation: residual at the trial point
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slslam_b200.h"
#include "lba_math.cuh"

namespace slslam {

constexpr int WIDE_NT = 256;
constexpr int WIDE_MAX_FREE = 64;
constexpr int WIDE_MAX_G = 32;       // CTAs per window
constexpr int WIDE_NPART = 4;        // scalars per group reduction

struct WideHdr {
  int C, Cf, L, N, n, max_iters, robust, pad;
  double huber_a, baseline, ftol, gtol, ptol, radius0;
  // inputs (observations in line-sorted order)
  const int* cam_s;        // [N] camera of sorted observation s
  const int* line_s;       // [N] line of sorted observation s
  const double* obs_s;     // [N][8]
  const int* line_start;   // [L + 1]
  const int* cam_free;     // [C] reduced index or -1
  const int* line_free;    // [L] 1: free, 0: constant or unobserved
  const int* cam_start;    // [Cf + 1] CSR over the observations of every free camera (sorted positions, ascending)
  const int* cam_obs;
  const int* pos;          // [L][Cf] sorted position of the observation of line l by free camera f, or -1
  const double* params_in;
  double* params_out;
  slslam_summary* summary;
  double* trace;
  // scratch: two copies (x and the trial point x') of the parameters, the rotation tables and the line sines / cosines
  double *camx[2], *camR[2], *linex[2], *ltrig[2];
  double *cscale, *lscale;
  double *r, *Jc, *Jl, *Z, *lineLU;        // [4N] [24N] [16N] [24N] [22L]
  double *S, *gc, *zu, *hd, *yc, *ub, *ab; // [n*n] [n] ...
  double* part;                            // [2][WIDE_MAX_G][WIDE_NPART] per-CTA partial scalars (double-buffered)
  double* flag;                            // [1] failure flag of the reduced solve (written by CTA 0)
  unsigned int* bar;                       // arrive counter of the group barrier (zeroed by the host before every launch)
};

struct WideCtx {
  int tid, lane, warp, rank, G;
  int gt, gsize, gw, gwarps;      // thread / warp index inside the group and their counts
  unsigned int target, xchg;
  unsigned int* bar;
  double* sh;                     // [WIDE_NPART][WIDE_NT] block reduction scratch (shared)
};

// Barrier over the G CTAs of a window: arrive (release) + spin (acquire) by thread 0 on a counter in L2; the acquire also
// drops this SM's stale L1 lines, so plain loads after it see what the other CTAs wrote before it.
__device__ __forceinline__ void wide_sync(WideCtx& c) {
  if (c.G == 1) { __syncthreads(); return; }
  c.target += (unsigned int)c.G;
  __syncthreads();
  if (c.tid == 0) {
    __threadfence();
    atomicAdd(c.bar, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c.bar) : "memory");
    } while ((int)(seen - c.target) < 0);
    __threadfence();
  }
  __syncthreads();
}

// v[0..K) of every thread -> the same K totals on every thread of every CTA of the group.  Bit k of `maxmask`: combine
// value k with max (values >= 0) instead of +.  Tree inside the CTA, rank order over the group; includes a group barrier.
template <int K>
__device__ __forceinline__ void wide_group_reduce(WideCtx& c, const WideHdr& h, double* v, unsigned int maxmask) {
  static_assert(K <= WIDE_NPART, "WIDE_NPART");
  double* sh = c.sh;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) sh[k * WIDE_NT + c.tid] = v[k];
  __syncthreads();
  for (int s = WIDE_NT >> 1; s > 0; s >>= 1) {
    if (c.tid < s) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double a = sh[k * WIDE_NT + c.tid], b = sh[k * WIDE_NT + c.tid + s];
        sh[k * WIDE_NT + c.tid] = ((maxmask >> k) & 1u) ? fmax(a, b) : a + b;
      }
    }
    __syncthreads();
  }
  if (c.G > 1) {
    double* slot = h.part + (size_t)(c.xchg & 1u) * WIDE_MAX_G * WIDE_NPART;
    ++c.xchg;
    if (c.tid < K) slot[c.rank * WIDE_NPART + c.tid] = sh[c.tid * WIDE_NT];
    wide_sync(c);
    if (c.tid < K) {
      double a = 0.0;
      for (int r = 0; r < c.G; ++r) {
        const double b = __ldcg(slot + r * WIDE_NPART + c.tid);
        a = ((maxmask >> c.tid) & 1u) ? fmax(a, b) : a + b;
      }
      sh[c.tid * WIDE_NT] = a;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = sh[k * WIDE_NT];
  __syncthreads();
}

// CTA-local reduction (every CTA of the group holds the same data and gets the same bits): no exchange.
template <int K>
__device__ __forceinline__ void wide_block_reduce(WideCtx& c, double* v, unsigned int maxmask) {
  double* sh = c.sh;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) sh[k * WIDE_NT + c.tid] = v[k];
  __syncthreads();
  for (int s = WIDE_NT >> 1; s > 0; s >>= 1) {
    if (c.tid < s) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double a = sh[k * WIDE_NT + c.tid], b = sh[k * WIDE_NT + c.tid + s];
        sh[k * WIDE_NT + c.tid] = ((maxmask >> k) & 1u) ? fmax(a, b) : a + b;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = sh[k * WIDE_NT];
  __syncthreads();
}

// residual (+ Jacobian) sweep at (cam precompute table `cR`, line sines / cosines `ltr`).  MODE 0: column norms for the
// Jacobi scale (J Huber-scaled only), 1: full linearisation (stores r, Jc, Jl scaled), 2: cost only.
template <int MODE>
__device__ void wide_sweep(const WideCtx& c, const WideHdr& h, const double* cR, const double* ltr, double* cost_out, double* fixed_out) {
  double cost = 0.0, fixed = 0.0;
  const bool robust = h.robust != 0;
  for (int s = c.gt; s < h.N; s += c.gsize) {
    const int cam = h.cam_s[s], l = h.line_s[s];
    const int cf = h.cam_free[cam];
    const bool lfree = h.line_free[l] != 0;
    if (MODE == 2 && cf < 0 && !lfree) continue;
    LineTrig lt;
    line_trig_sc(ltr + 8 * (size_t)l, lt);
    double ob[8], r[4], Jc[24], Jl[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) ob[k] = h.obs_s[8 * (size_t)s + k];
    if (MODE == 2) obs_eval<false>(cR + CAM_STRIDE * (size_t)cam, lt, ob, h.baseline, r, nullptr, nullptr);
    else obs_eval<true>(cR + CAM_STRIDE * (size_t)cam, lt, ob, h.baseline, r, Jc, Jl);
    double w;
    const double rho = huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
    if (cf >= 0 || lfree) cost += 0.5 * rho; else fixed += 0.5 * rho;
    if (MODE == 2) continue;
    const double wc = cf >= 0 ? w : 0.0, wl = lfree ? w : 0.0;
    double* Jco = h.Jc + 24 * (size_t)s;
    double* Jlo = h.Jl + 16 * (size_t)s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h.r[4 * (size_t)s + k] = r[k] * w;
#pragma unroll
      for (int j = 0; j < 6; ++j) Jco[6 * k + j] = Jc[6 * k + j] * wc * (MODE == 1 && cf >= 0 ? h.cscale[6 * cf + j] : 1.0);
#pragma unroll
      for (int j = 0; j < 4; ++j) Jlo[4 * k + j] = Jl[4 * k + j] * wl * (MODE == 1 ? h.lscale[4 * (size_t)l + j] : 1.0);
    }
  }
  *cost_out = cost; *fixed_out = fixed;
}

__device__ __forceinline__ double wide_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduced solve (S + D_c) y = g_c - sum Z u by ONE CTA: right-looking block elimination in global memory (L1 / L2), the
// unscaled panel of the current block column in shared memory.  Returns false when a pivot block is not positive.
__device__ bool wide_reduced_solve(const WideHdr& h, double inv_radius, double* pan, double* Wsm, double* bcast, const int* tri) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cf = h.Cf, n = h.n;
  for (int i = tid; i < n; i += WIDE_NT) {
    h.yc[i] = h.gc[i] - h.zu[i];
    h.ab[i] = 0.0;
    h.S[(size_t)i * n + i] += fmin(fmax(h.hd[i], 1e-6), 1e32) * inv_radius;
  }
  if (tid == 0) bcast[0] = 0.0;
  __syncthreads();
  for (int J = 0; J < Cf; ++J) {
    const int nb = Cf - J - 1;
    // pivot inverse by warp 0 (all lanes the same values), to shared memory
    if (warp == 0) {
      double A[21], Ai[6], Si[6], M[9], S3[6], W[21];
#define L6I(p, q) ((p) * ((p) + 1) / 2 + (q))
#define SY3(mm, r, cc) mm[(r) <= (cc) ? ((r) == 0 ? (cc) : (r) == 1 ? 2 + (cc) : 5) : ((cc) == 0 ? (r) : (cc) == 1 ? 2 + (r) : 5)]
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q <= p; ++q) A[L6I(p, q)] = h.S[(size_t)(6 * J + p) * n + 6 * J + q];
      bool ok = spd3_inverse(A[L6I(0, 0)], A[L6I(1, 0)], A[L6I(2, 0)], A[L6I(1, 1)], A[L6I(2, 1)], A[L6I(2, 2)], Ai);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          M[3 * r + cc] = A[L6I(3 + r, 0)] * SY3(Ai, 0, cc) + A[L6I(3 + r, 1)] * SY3(Ai, 1, cc) + A[L6I(3 + r, 2)] * SY3(Ai, 2, cc);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = r; cc < 3; ++cc)
          SY3(S3, r, cc) = A[L6I(3 + cc, 3 + r)] - (M[3 * r] * A[L6I(3 + cc, 0)] + M[3 * r + 1] * A[L6I(3 + cc, 1)] + M[3 * r + 2] * A[L6I(3 + cc, 2)]);
      ok = spd3_inverse(S3[0], S3[1], S3[2], S3[3], S3[4], S3[5], Si) && ok;
      double W21[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          W21[3 * r + cc] = -(SY3(Si, r, 0) * M[cc] + SY3(Si, r, 1) * M[3 + cc] + SY3(Si, r, 2) * M[6 + cc]);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) {
          W[L6I(r, cc)] = SY3(Ai, r, cc) - (M[r] * W21[cc] + M[3 + r] * W21[3 + cc] + M[6 + r] * W21[6 + cc]);
          W[L6I(3 + r, 3 + cc)] = SY3(Si, r, cc);
        }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) W[L6I(3 + r, cc)] = W21[3 * r + cc];
#undef SY3
      if (lane == 0) {
        if (!ok) bcast[0] = 1.0;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
          for (int q = 0; q < 6; ++q) Wsm[6 * p + q] = W[p >= q ? L6I(p, q) : L6I(q, p)];
      }
#undef L6I
    }
    __syncthreads();
    // panel rows P_I = A_IJ W (original rows kept in shared memory), u_J = W b_J
    for (int t = tid; t < 6 * nb + 6; t += WIDE_NT) {
      if (t < 6 * nb) {
        const int bI = t / 6, p = t - 6 * bI, I = J + 1 + bI;
        double* a = h.S + (size_t)(6 * I + p) * n + 6 * J;
        double av[6], pv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) av[k] = a[k];
#pragma unroll
        for (int q = 0; q < 6; ++q) pv[q] = (av[0] * Wsm[q] + av[1] * Wsm[6 + q] + av[2] * Wsm[12 + q]) + (av[3] * Wsm[18 + q] + av[4] * Wsm[24 + q] + av[5] * Wsm[30 + q]);
#pragma unroll
        for (int k = 0; k < 6; ++k) { pan[36 * bI + 6 * p + k] = av[k]; a[k] = pv[k]; }
      } else {
        const int q = t - 6 * nb;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += h.yc[6 * J + k] * Wsm[6 * k + q];
        h.ub[6 * J + q] = sacc;
      }
    }
    __syncthreads();
    // trailing update A_IK -= P_I A_KJ^T (I >= K > J): a thread owns row p of a block; b_I -= P_I b_J
    const int nrow = nb * (nb + 1) / 2 * 6;
    for (int e = tid; e < nrow + 6 * nb; e += WIDE_NT) {
      if (e < nrow) {
        const int blk = e / 6, p = e - 6 * blk;
        const int bi = tri[blk] >> 8, bk = tri[blk] & 0xff;
        const int I = J + 1 + bi, K = J + 1 + bk;
        const double* pi = h.S + (size_t)(6 * I + p) * n + 6 * J;
        const double* ak = pan + 36 * bk;
        double* dst = h.S + (size_t)(6 * I + p) * n + 6 * K;
        double a6[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) a6[k] = pi[k];
#pragma unroll
        for (int q = 0; q < 6; ++q)
          dst[q] -= (a6[0] * ak[6 * q] + a6[1] * ak[6 * q + 1] + a6[2] * ak[6 * q + 2]) + (a6[3] * ak[6 * q + 3] + a6[4] * ak[6 * q + 4] + a6[5] * ak[6 * q + 5]);
      } else {
        const int rI = e - nrow, bI = rI / 6, p = rI - 6 * bI, I = J + 1 + bI;
        const double* pi = h.S + (size_t)(6 * I + p) * n + 6 * J;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += pi[k] * h.yc[6 * J + k];
        h.yc[6 * I + p] -= sacc;
      }
    }
    __syncthreads();
  }
  const bool ok = bcast[0] == 0.0;
  // back-substitution (warp 0): y_J = u_J - acc_J, acc_K += P_JK^T y_J for K < J
  if (warp == 0 && ok) {
    for (int J = Cf - 1; J >= 0; --J) {
      double y[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) y[k] = h.ub[6 * J + k] - h.ab[6 * J + k];
      __syncwarp();
      if (lane < 6) h.yc[6 * J + lane] = h.ub[6 * J + lane] - h.ab[6 * J + lane];
      for (int e = lane; e < 6 * J; e += 32) {
        const int K = e / 6, q = e - 6 * K;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += h.S[(size_t)(6 * J + k) * n + 6 * K + q] * y[k];
        h.ab[6 * K + q] += sacc;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  return ok;
}

// grid = (windows resident at a time) x G CTAs; window w of a launch is solved by the group blockIdx.x / G, then w +
// groups, ... (every window has its own barrier counter and partial-sum slots)
__global__ void __launch_bounds__(WIDE_NT, 1) lba_wide_kernel(const WideHdr* __restrict__ hdrs, int nwin, int G) {
  extern __shared__ __align__(16) double wsm[];
  double* pan = wsm + WIDE_NPART * WIDE_NT;       // [WIDE_MAX_FREE][36] original panel rows of the current block column
  double* Wsm = pan + WIDE_MAX_FREE * 36;         // [36] pivot inverse
  double* bcast = Wsm + 36;                       // [16] broadcast scalars
  int* tri = reinterpret_cast<int*>(bcast + 16);  // block of a lower block triangle -> row << 8 | column
  for (int k = threadIdx.x; k < WIDE_MAX_FREE * (WIDE_MAX_FREE + 1) / 2; k += WIDE_NT) {
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= k) ++I;
    tri[k] = (I << 8) | (k - I * (I + 1) / 2);
  }
  __syncthreads();
  WideCtx c;
  c.sh = wsm;
  c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
  c.G = G; c.rank = (int)(blockIdx.x % (unsigned)G);
  c.gt = c.rank * WIDE_NT + c.tid; c.gsize = G * WIDE_NT;
  c.gw = c.rank * (WIDE_NT / 32) + c.warp; c.gwarps = G * (WIDE_NT / 32);
  const int tid = c.tid, lane = c.lane;
  const int groups = (int)(gridDim.x / (unsigned)G);
  for (int win = (int)(blockIdx.x / (unsigned)G); win < nwin; win += groups) {
  const WideHdr& h = hdrs[win];
  c.bar = h.bar; c.target = 0; c.xchg = 0;
  const int C = h.C, Cf = h.Cf, L = h.L, N = h.N, n = h.n;
  int cur = 0;                                    // which copy holds x (the other one the trial point)
  // ---- parameters, rotation tables, line sines / cosines at x0 ----
  for (int cc = c.gt; cc < C; cc += c.gsize) {
#pragma unroll
    for (int j = 0; j < 6; ++j) h.camx[0][6 * (size_t)cc + j] = h.params_in[6 * (size_t)cc + j];
    cam_precompute(h.camx[0] + 6 * (size_t)cc, h.camR[0] + CAM_STRIDE * (size_t)cc, true);
  }
  for (int i = c.gt; i < 4 * L; i += c.gsize) {
    const double v = h.params_in[6 * (size_t)C + i];
    h.linex[0][i] = v;
    double sv, cv;
    sincos(v, &sv, &cv);
    h.ltrig[0][2 * (size_t)i] = sv; h.ltrig[0][2 * (size_t)i + 1] = cv;
    h.lscale[i] = 1.0;
  }
  for (int i = c.gt; i < n; i += c.gsize) h.cscale[i] = 1.0;
  wide_sync(c);
  // ---- Jacobi scale from the column norms at x0 (Huber-scaled Jacobian), initial and fixed cost ----
  double pc, pf;
  wide_sweep<0>(c, h, h.camR[0], h.ltrig[0], &pc, &pf);
  double cost, fixed_cost;
  {
    double v[2] = {pc, pf};
    wide_group_reduce<2>(c, h, v, 0u);
    cost = v[0]; fixed_cost = v[1];
  }
  const double initial_cost = cost + fixed_cost;
  for (int l = c.gw; l < L; l += c.gwarps) {
    double cn[4] = {0, 0, 0, 0};
    for (int s = h.line_start[l] + lane; s < h.line_start[l + 1]; s += 32) {
      const double* J = h.Jl + 16 * (size_t)s;
#pragma unroll
      for (int j = 0; j < 4; ++j) cn[j] += J[j] * J[j] + J[4 + j] * J[4 + j] + J[8 + j] * J[8 + j] + J[12 + j] * J[12 + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { cn[j] = wide_warp_sum(cn[j]); if (lane == 0) h.lscale[4 * (size_t)l + j] = 1.0 / (1.0 + sqrt(cn[j])); }
  }
  for (int i = c.gt; i < n; i += c.gsize) {
    const int f = i / 6, j = i - 6 * f;
    double cn = 0.0;
    for (int t = h.cam_start[f]; t < h.cam_start[f + 1]; ++t) {
      const double* J = h.Jc + 24 * (size_t)h.cam_obs[t];
      cn += J[j] * J[j] + J[6 + j] * J[6 + j] + J[12 + j] * J[12 + j] + J[18 + j] * J[18 + j];
    }
    h.cscale[i] = 1.0 / (1.0 + sqrt(cn));
  }
  wide_sync(c);

  double radius = h.radius0, decrease_factor = 2.0, gmax = 0.0, gtol_abs = 0.0;
  int successful = 0, unsuccessful = 0, invalid = 0, term = SLSLAM_NO_CONVERGENCE, iters = 0;
  bool first_lin = true, grad_pending = false;
  for (int it = 0; it <= h.max_iters; ++it) {
    const bool last = it == h.max_iters;
    if (last && !grad_pending) break;
    double* camx = h.camx[cur]; double* camxt = h.camx[cur ^ 1];
    double* camR = h.camR[cur]; double* camRt = h.camR[cur ^ 1];
    double* linex = h.linex[cur]; double* linext = h.linex[cur ^ 1];
    double* ltrig = h.ltrig[cur]; double* ltrigt = h.ltrig[cur ^ 1];
    // ---- linearise at x ----
    wide_sweep<1>(c, h, camR, ltrig, &pc, &pf);
    {
      double v[1] = {pc};
      wide_group_reduce<1>(c, h, v, 0u);          // (the barrier inside also publishes r, Jc, Jl)
      cost = v[0];
    }
    // ---- lines: H_ll, g_l, LM diagonal, Cholesky, u; Z per observation; line part of the gradient norm ----
    double gm = 0.0, failf = 0.0;
    const double inv_radius = 1.0 / radius;
    for (int l = c.gw; l < L; l += c.gwarps) {
      const int s0 = h.line_start[l], s1 = h.line_start[l + 1];
      if (!h.line_free[l] || s0 == s1) continue;
      double hg[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) hg[k] = 0.0;
      for (int s = s0 + lane; s < s1; s += 32) {
        const double* J = h.Jl + 16 * (size_t)s;
        const double* r = h.r + 4 * (size_t)s;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
          for (int q = 0; q <= p; ++q) hg[p * (p + 1) / 2 + q] += J[p] * J[q] + J[4 + p] * J[4 + q] + J[8 + p] * J[8 + q] + J[12 + p] * J[12 + q];
          hg[10 + p] += J[p] * r[0] + J[4 + p] * r[1] + J[8 + p] * r[2] + J[12 + p] * r[3];
        }
      }
#pragma unroll
      for (int k = 0; k < 14; ++k) hg[k] = wide_warp_sum(hg[k]);
      double D[4], Lm[10], u[4], inv[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) D[p] = fmin(fmax(hg[p * (p + 1) / 2 + p], 1e-6), 1e32) * inv_radius;
      bool ok = true;
      {
        const double a00 = hg[0] + D[0];
        ok = ok && (a00 > 0.0); inv[0] = pivot_rsqrt(a00); Lm[0] = a00 * inv[0];
        Lm[1] = hg[1] * inv[0]; Lm[3] = hg[3] * inv[0]; Lm[6] = hg[6] * inv[0];
        const double a11 = hg[2] + D[1] - Lm[1] * Lm[1];
        ok = ok && (a11 > 0.0); inv[1] = pivot_rsqrt(a11); Lm[2] = a11 * inv[1];
        Lm[4] = (hg[4] - Lm[3] * Lm[1]) * inv[1]; Lm[7] = (hg[7] - Lm[6] * Lm[1]) * inv[1];
        const double a22 = hg[5] + D[2] - Lm[3] * Lm[3] - Lm[4] * Lm[4];
        ok = ok && (a22 > 0.0); inv[2] = pivot_rsqrt(a22); Lm[5] = a22 * inv[2];
        Lm[8] = (hg[8] - Lm[6] * Lm[3] - Lm[7] * Lm[4]) * inv[2];
        const double a33 = hg[9] + D[3] - Lm[6] * Lm[6] - Lm[7] * Lm[7] - Lm[8] * Lm[8];
        ok = ok && (a33 > 0.0); inv[3] = pivot_rsqrt(a33); Lm[9] = a33 * inv[3];
      }
      if (!ok) failf = 1.0;
      u[0] = hg[10] * inv[0];
      u[1] = (hg[11] - Lm[1] * u[0]) * inv[1];
      u[2] = (hg[12] - Lm[3] * u[0] - Lm[4] * u[1]) * inv[2];
      u[3] = (hg[13] - Lm[6] * u[0] - Lm[7] * u[1] - Lm[8] * u[2]) * inv[3];
      if (lane == 0) {
        double* o = h.lineLU + 22 * (size_t)l;
#pragma unroll
        for (int k = 0; k < 10; ++k) o[k] = Lm[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[10 + k] = u[k]; o[14 + k] = D[k]; o[18 + k] = inv[k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) gm = fmax(gm, fabs(hg[10 + k] / h.lscale[4 * (size_t)l + k]));
      }
      for (int s = s0 + lane; s < s1; s += 32) {
        const double* Jc = h.Jc + 24 * (size_t)s;
        const double* Jl = h.Jl + 16 * (size_t)s;
        double* Zo = h.Z + 24 * (size_t)s;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          double W[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) W[q] = Jc[p] * Jl[q] + Jc[6 + p] * Jl[4 + q] + Jc[12 + p] * Jl[8 + q] + Jc[18 + p] * Jl[12 + q];
          const double z0 = W[0] * inv[0];
          const double z1 = (W[1] - z0 * Lm[1]) * inv[1];
          const double z2 = (W[2] - z0 * Lm[3] - z1 * Lm[4]) * inv[2];
          const double z3 = (W[3] - z0 * Lm[6] - z1 * Lm[7] - z2 * Lm[8]) * inv[3];
          Zo[4 * p] = z0; Zo[4 * p + 1] = z1; Zo[4 * p + 2] = z2; Zo[4 * p + 3] = z3;
        }
      }
    }
    // constant lines: their observations have Jl = 0, no Schur term: Z = 0
    for (int s = c.gt; s < N; s += c.gsize) {
      if (!h.line_free[h.line_s[s]]) { double* Zo = h.Z + 24 * (size_t)s; for (int k = 0; k < 24; ++k) Zo[k] = 0.0; }
    }
    bool line_fail;
    {
      double v[2] = {failf, gm};
      wide_group_reduce<2>(c, h, v, 3u);          // (publishes Z and lineLU)
      line_fail = v[0] != 0.0; gm = v[1];
    }
    // ---- cameras: diagonal blocks, g_c, Z u, diag H_cc ----
    for (int i = c.gt; i < Cf * 39; i += c.gsize) {
      const int f = i / 39, e = i - 39 * f;
      double acc = 0.0;
      int p = 0, q = 0;
      if (e < 21) { while ((p + 1) * (p + 2) / 2 <= e) ++p; q = e - p * (p + 1) / 2; }
      else p = (e - 21) % 6;
      for (int t = h.cam_start[f]; t < h.cam_start[f + 1]; ++t) {
        const int s = h.cam_obs[t];
        const double* Jc = h.Jc + 24 * (size_t)s;
        const double* Z = h.Z + 24 * (size_t)s;
        if (e < 21) {
          acc += Jc[p] * Jc[q] + Jc[6 + p] * Jc[6 + q] + Jc[12 + p] * Jc[12 + q] + Jc[18 + p] * Jc[18 + q]
                 - (Z[4 * p] * Z[4 * q] + Z[4 * p + 1] * Z[4 * q + 1] + Z[4 * p + 2] * Z[4 * q + 2] + Z[4 * p + 3] * Z[4 * q + 3]);
        } else if (e < 27) {
          const double* r = h.r + 4 * (size_t)s;
          acc += Jc[p] * r[0] + Jc[6 + p] * r[1] + Jc[12 + p] * r[2] + Jc[18 + p] * r[3];
        } else if (e < 33) {
          const int l = h.line_s[s];
          if (h.line_free[l]) {
            const double* u = h.lineLU + 22 * (size_t)l + 10;
            acc += Z[4 * p] * u[0] + Z[4 * p + 1] * u[1] + Z[4 * p + 2] * u[2] + Z[4 * p + 3] * u[3];
          }
        } else {
          acc += Jc[p] * Jc[p] + Jc[6 + p] * Jc[6 + p] + Jc[12 + p] * Jc[12 + p] + Jc[18 + p] * Jc[18 + p];
        }
      }
      if (e < 21) { h.S[(size_t)(6 * f + p) * n + 6 * f + q] = acc; h.S[(size_t)(6 * f + q) * n + 6 * f + p] = acc; }
      else if (e < 27) h.gc[6 * f + p] = acc;
      else if (e < 33) h.zu[6 * f + p] = acc;
      else h.hd[6 * f + p] = acc;
    }
    // ---- pairs: off-diagonal blocks (I > K) ----
    if (!last) {
      const int nblk = Cf * (Cf - 1) / 2;
      for (int i = c.gt; i < nblk * 36; i += c.gsize) {
        const int b = i / 36, pq = i - 36 * b, p = pq / 6, q = pq - 6 * p;
        int I = 1; while (I * (I + 1) / 2 <= b) ++I;          // b = I (I - 1) / 2 + K, K < I
        const int K = b - I * (I - 1) / 2;
        double acc = 0.0;
        for (int l = 0; l < L; ++l) {
          const int sa = h.pos[(size_t)l * Cf + I];
          if (sa < 0) continue;
          const int sb = h.pos[(size_t)l * Cf + K];
          if (sb < 0) continue;
          const double* Za = h.Z + 24 * (size_t)sa + 4 * p;
          const double* Zb = h.Z + 24 * (size_t)sb + 4 * q;
          acc += Za[0] * Zb[0] + Za[1] * Zb[1] + Za[2] * Zb[2] + Za[3] * Zb[3];
        }
        h.S[(size_t)(6 * I + p) * n + 6 * K + q] = -acc;
      }
    }
    wide_sync(c);
    // ---- gradient max norm (unscaled Jacobian), |x|^2 of the free blocks: the same on every CTA, no exchange ----
    double x_norm2;
    {
      double pg = 0.0, px = 0.0;
      for (int i = tid; i < n; i += WIDE_NT) pg = fmax(pg, fabs(h.gc[i] / h.cscale[i]));
      for (int cc = tid; cc < C; cc += WIDE_NT) if (h.cam_free[cc] >= 0) for (int j = 0; j < 6; ++j) px += camx[6 * cc + j] * camx[6 * cc + j];
      for (int l = tid; l < L; l += WIDE_NT) if (h.line_free[l] && h.line_start[l + 1] > h.line_start[l]) for (int j = 0; j < 4; ++j) px += linex[4 * l + j] * linex[4 * l + j];
      double v[2] = {pg, px};
      wide_block_reduce<2>(c, v, 1u);
      gmax = fmax(gm, v[0]); x_norm2 = v[1];
    }
    if (first_lin) { gtol_abs = h.gtol * fmax(gmax, 2.220446049250313e-16); first_lin = false; }
    grad_pending = false;
    if (gmax <= gtol_abs) { term = SLSLAM_GRADIENT_TOLERANCE; break; }
    if (last) break;
    iters = it + 1;
    double* tr = (h.trace && c.gt == 0) ? h.trace + (size_t)it * SLSLAM_TRACE_WIDTH : nullptr;
    if (tr) { tr[0] = cost; tr[1] = 0; tr[2] = 0; tr[3] = radius; tr[4] = 0; tr[5] = 0; tr[6] = gmax; tr[7] = 0; }
    // ---- reduced solve by CTA 0 of the group ----
    if (c.rank == 0) {
      const bool sok = wide_reduced_solve(h, inv_radius, pan, Wsm, bcast, tri);
      if (tid == 0) *h.flag = sok ? 0.0 : 1.0;
    }
    wide_sync(c);
    bool ok = !line_fail && __ldcg(h.flag) == 0.0;
    // camera part of the model decrease, |delta|^2, finiteness (the same on every CTA)
    double model = 0.0, dn2 = 0.0;
    {
      double v[3] = {0.0, 0.0, 0.0};
      if (ok) {
        for (int i = tid; i < n; i += WIDE_NT) {
          const double y = h.yc[i];
          v[0] += 0.5 * y * (h.gc[i] + fmin(fmax(h.hd[i], 1e-6), 1e32) * inv_radius * y);
          const double dl = y * h.cscale[i];
          v[1] += dl * dl;
          if (!isfinite(y)) v[2] = 1.0;
        }
      }
      wide_block_reduce<3>(c, v, 4u);
      model = v[0]; dn2 = v[1];
      if (v[2] != 0.0) ok = false;
    }
    double new_cost = 0.0;
    if (ok) {
      // ---- trial point: cameras, lines (y_l = L^-T (u - sum Z^T y_c)), cost ----
      for (int cc = c.gt; cc < C; cc += c.gsize) {
        const int cf = h.cam_free[cc];
#pragma unroll
        for (int j = 0; j < 6; ++j) camxt[6 * (size_t)cc + j] = camx[6 * (size_t)cc + j] - (cf >= 0 ? h.yc[6 * cf + j] * h.cscale[6 * cf + j] : 0.0);
        cam_precompute(camxt + 6 * (size_t)cc, camRt + CAM_STRIDE * (size_t)cc, true);
      }
      double plm = 0.0, pld = 0.0;
      for (int l = c.gw; l < L; l += c.gwarps) {
        const int s0 = h.line_start[l], s1 = h.line_start[l + 1];
        double nl[4];
        if (!h.line_free[l] || s0 == s1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) nl[k] = linex[4 * (size_t)l + k];
        } else {
          double v[4] = {0, 0, 0, 0};
          for (int s = s0 + lane; s < s1; s += 32) {
            const int cf = h.cam_free[h.cam_s[s]];
            if (cf < 0) continue;
            const double* Z = h.Z + 24 * (size_t)s;
            const double* y = h.yc + 6 * cf;
#pragma unroll
            for (int p = 0; p < 6; ++p) { v[0] += Z[4 * p] * y[p]; v[1] += Z[4 * p + 1] * y[p]; v[2] += Z[4 * p + 2] * y[p]; v[3] += Z[4 * p + 3] * y[p]; }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = wide_warp_sum(v[k]);
          const double* lu = h.lineLU + 22 * (size_t)l;
          double yl[4];
          const double w3 = lu[13] - v[3], w2 = lu[12] - v[2], w1 = lu[11] - v[1], w0 = lu[10] - v[0];
          yl[3] = w3 * lu[21];
          yl[2] = (w2 - lu[8] * yl[3]) * lu[20];
          yl[1] = (w1 - lu[4] * yl[2] - lu[7] * yl[3]) * lu[19];
          yl[0] = (w0 - lu[1] * yl[1] - lu[3] * yl[2] - lu[6] * yl[3]) * lu[18];
          const double u0 = lu[10], u1 = lu[11], u2 = lu[12], u3 = lu[13];
          const double g0 = lu[0] * u0, g1 = lu[1] * u0 + lu[2] * u1, g2 = lu[3] * u0 + lu[4] * u1 + lu[5] * u2,
                       g3 = lu[6] * u0 + lu[7] * u1 + lu[8] * u2 + lu[9] * u3;
          if (lane == 0) plm += 0.5 * (yl[0] * (g0 + lu[14] * yl[0]) + yl[1] * (g1 + lu[15] * yl[1]) + yl[2] * (g2 + lu[16] * yl[2]) + yl[3] * (g3 + lu[17] * yl[3]));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double dl = yl[k] * h.lscale[4 * (size_t)l + k];
            if (lane == 0) pld += dl * dl;
            nl[k] = linex[4 * (size_t)l + k] - dl;
          }
        }
        if (lane < 4) {
          const double v = lane == 0 ? nl[0] : lane == 1 ? nl[1] : lane == 2 ? nl[2] : nl[3];
          double sv, cv;
          sincos(v, &sv, &cv);
          linext[4 * (size_t)l + lane] = v;
          ltrigt[8 * (size_t)l + 2 * lane] = sv; ltrigt[8 * (size_t)l + 2 * lane + 1] = cv;
        }
      }
      {
        double v[2] = {plm, pld};
        wide_group_reduce<2>(c, h, v, 0u);        // (publishes the trial cameras and lines)
        model += v[0]; dn2 += v[1];
      }
      wide_sweep<2>(c, h, camRt, ltrigt, &pc, &pf);
      {
        double v[1] = {pc};
        wide_group_reduce<1>(c, h, v, 0u);
        new_cost = v[0];
      }
    }
    if (tr) tr[2] = model;
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
      cur ^= 1;                                   // the trial point becomes x: no copies
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
  // ---- write back: blocks no observation touches keep their input bits (the host pre-copies the input) ----
  wide_sync(c);
  for (int i = c.gt; i < 6 * C; i += c.gsize) h.params_out[i] = h.camx[cur][i];
  for (int i = c.gt; i < 4 * L; i += c.gsize) h.params_out[6 * (size_t)C + i] = h.linex[cur][i];
  if (c.gt == 0) {
    slslam_summary s;
    s.initial_cost = initial_cost; s.final_cost = cost + fixed_cost; s.fixed_cost = fixed_cost; s.gradient_max_norm = gmax;
    s.num_successful_steps = successful; s.num_unsuccessful_steps = unsuccessful; s.termination_type = term; s.iterations = iters;
    *h.summary = s;
  }
  }   // windows of this group
}

}  // namespace slslam
