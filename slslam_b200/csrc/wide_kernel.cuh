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
#include <stdio.h>

#include "../../include/slslam_b200.h"
#include "lba_kernel.cuh"     // warp_reduce_scatter32, warp_sum
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
  double *S, *P, *gc, *zu, *hd, *yc, *ub, *ab; // [n*n] [n*n] [n] ...  (P: the scaled panels A_IJ W_J of the block elimination)
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
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c.bar) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c.bar) : "memory");
    } while ((int)(seen - c.target) < 0);
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

// 16-byte global loads / stores of rows of doubles (rows are 16-byte aligned): lanes read rows of their own, so every
// load instruction touches 32 sectors; pairs of doubles halve the number of sector requests, which is what bounds the
// row-gathering phases
template <int NV>
__device__ __forceinline__ void ld_row(const double* __restrict__ src, double* dst) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int k = 0; k < NV / 2; ++k) { const double2 t = s2[k]; dst[2 * k] = t.x; dst[2 * k + 1] = t.y; }
}
template <int NV>
__device__ __forceinline__ void st_row(double* dst, const double* src) {
  double2* d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
  for (int k = 0; k < NV / 2; ++k) d2[k] = make_double2(src[2 * k], src[2 * k + 1]);
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
    ld_row<8>(h.obs_s + 8 * (size_t)s, ob);
    if (MODE == 2) obs_eval<false>(cR + CAM_STRIDE * (size_t)cam, lt, ob, h.baseline, r, nullptr, nullptr);
    else obs_eval<true>(cR + CAM_STRIDE * (size_t)cam, lt, ob, h.baseline, r, Jc, Jl);
    double w;
    const double rho = huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3], h.huber_a, robust, w);
    if (cf >= 0 || lfree) cost += 0.5 * rho; else fixed += 0.5 * rho;
    if (MODE == 2) continue;
    const double wc = cf >= 0 ? w : 0.0, wl = lfree ? w : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[k] *= w;
#pragma unroll
      for (int j = 0; j < 6; ++j) Jc[6 * k + j] = Jc[6 * k + j] * wc * (MODE == 1 && cf >= 0 ? h.cscale[6 * cf + j] : 1.0);
#pragma unroll
      for (int j = 0; j < 4; ++j) Jl[4 * k + j] = Jl[4 * k + j] * wl * (MODE == 1 ? h.lscale[4 * (size_t)l + j] : 1.0);
    }
    st_row<4>(h.r + 4 * (size_t)s, r);
    st_row<24>(h.Jc + 24 * (size_t)s, Jc);
    st_row<16>(h.Jl + 16 * (size_t)s, Jl);
  }
  *cost_out = cost; *fixed_out = fixed;
}

__device__ __forceinline__ double wide_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// W = A^-1 of a 6x6 pivot block (lower triangle read from A0, row stride ld) by one warp, all lanes the same values;
// lane 0 stores W (full, symmetric) to shared memory.  false when the block is not positive definite.
__device__ __forceinline__ bool wide_pivot_inverse(const double* A0, size_t ld, double* Wsm, int lane) {
  double A[21], Ai[6], Si[6], M[9], S3[6], W[21];
#define L6I(p, q) ((p) * ((p) + 1) / 2 + (q))
#define SY3(mm, r, cc) mm[(r) <= (cc) ? ((r) == 0 ? (cc) : (r) == 1 ? 2 + (cc) : 5) : ((cc) == 0 ? (r) : (cc) == 1 ? 2 + (r) : 5)]
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int q = 0; q <= p; ++q) A[L6I(p, q)] = A0[(size_t)p * ld + q];
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
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) Wsm[6 * p + q] = W[p >= q ? L6I(p, q) : L6I(q, p)];
  }
#undef L6I
  return ok;
}

constexpr int WIDE_TAIL_MAX = 24;    // the last block columns, whose trailing matrix (lower block triangle, 24 * 25 / 2 blocks = 86 KB) is kept in
                                     // shared memory: big columns are cheaper spread over the group, small ones without barriers and L2 trips

// The last m <= WIDE_TAIL_MAX block columns of the elimination by ONE CTA with the trailing matrix in shared memory (packed
// lower block triangle): no group barrier and no L2 round trip per column any more.  Panel rows are scaled in place, their
// unscaled copies kept in `pan` for the trailing update, and also written to P in global memory for the back-substitution;
// right-hand side and u in shared memory, copied back at the end.  Returns false when a pivot block is not positive.
__device__ bool wide_solve_tail(const WideHdr& h, int c0, double* T, double* pan, double* ycs, double* ubs, double* Wsm, const int* tri) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cf = h.Cf, n = h.n, m = Cf - c0;
  bool ok = true;
  for (int i = tid; i < m * (m + 1) / 2 * 36; i += WIDE_NT) {
    const int blk = i / 36, pq = i - 36 * blk, p = pq / 6, q = pq - 6 * p;
    const int bi = tri[blk] >> 8, bk = tri[blk] & 0xff;
    T[i] = h.S[(size_t)(6 * (c0 + bi) + p) * n + 6 * (c0 + bk) + q];
  }
  for (int i = tid; i < 6 * m; i += WIDE_NT) ycs[i] = h.yc[6 * c0 + i];
  __syncthreads();
  for (int Jl = 0; Jl < m; ++Jl) {
    const int nb = m - Jl - 1, J = c0 + Jl;
    double* AJJ = T + (Jl * (Jl + 1) / 2 + Jl) * 36;
    if (warp == 0) {
      const bool pok = wide_pivot_inverse(AJJ, 6, Wsm, lane);
      ok = ok && pok;
    }
    __syncthreads();
    for (int t = tid; t < 6 * nb + 6; t += WIDE_NT) {
      if (t < 6 * nb) {
        const int bI = t / 6, p = t - 6 * bI, Il = Jl + 1 + bI;
        double* a = T + (Il * (Il + 1) / 2 + Jl) * 36 + 6 * p;
        double av[6], pv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) av[k] = a[k];
#pragma unroll
        for (int q = 0; q < 6; ++q) pv[q] = (av[0] * Wsm[q] + av[1] * Wsm[6 + q] + av[2] * Wsm[12 + q]) + (av[3] * Wsm[18 + q] + av[4] * Wsm[24 + q] + av[5] * Wsm[30 + q]);
        double* pg = h.P + (size_t)(6 * (c0 + Il) + p) * n + 6 * J;
#pragma unroll
        for (int k = 0; k < 6; ++k) { pan[36 * bI + 6 * p + k] = av[k]; a[k] = pv[k]; pg[k] = pv[k]; }
      } else {
        const int q = t - 6 * nb;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += ycs[6 * Jl + k] * Wsm[6 * k + q];
        ubs[6 * Jl + q] = sacc;
      }
    }
    __syncthreads();
    const int nrow = nb * (nb + 1) / 2 * 6;
    for (int e = tid; e < nrow + 6 * nb; e += WIDE_NT) {
      if (e < nrow) {
        const int blk = e / 6, p = e - 6 * blk;
        const int bi = tri[blk] >> 8, bk = tri[blk] & 0xff;
        const int Il = Jl + 1 + bi, Kl = Jl + 1 + bk;
        const double* pi = T + (Il * (Il + 1) / 2 + Jl) * 36 + 6 * p;
        const double* ak = pan + 36 * bk;
        double* dst = T + (Il * (Il + 1) / 2 + Kl) * 36 + 6 * p;
        double a6[6], o[6];
        const double2* pi2 = reinterpret_cast<const double2*>(pi);
        const double2* ak2 = reinterpret_cast<const double2*>(ak);
        double2* dst2 = reinterpret_cast<double2*>(dst);
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double2 t0 = pi2[k], t1 = dst2[k]; a6[2 * k] = t0.x; a6[2 * k + 1] = t0.y; o[2 * k] = t1.x; o[2 * k + 1] = t1.y; }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double2 k0 = ak2[3 * q], k1 = ak2[3 * q + 1], k2 = ak2[3 * q + 2];
          o[q] -= (a6[0] * k0.x + a6[1] * k0.y + a6[2] * k1.x) + (a6[3] * k1.y + a6[4] * k2.x + a6[5] * k2.y);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) dst2[k] = make_double2(o[2 * k], o[2 * k + 1]);
      } else {
        const int rI = e - nrow, bI = rI / 6, p = rI - 6 * bI, Il = Jl + 1 + bI;
        const double* pi = T + (Il * (Il + 1) / 2 + Jl) * 36 + 6 * p;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += pi[k] * ycs[6 * Jl + k];
        ycs[6 * Il + p] -= sacc;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < 6 * m; i += WIDE_NT) { h.ub[6 * c0 + i] = ubs[i]; h.yc[6 * c0 + i] = ycs[i]; }
  // every lane of warp 0 saw the same pivots; hand the verdict to the whole CTA
  if (tid == 0) Wsm[0] = ok ? 0.0 : 1.0;
  __syncthreads();
  const bool all_ok = Wsm[0] == 0.0;
  __syncthreads();
  return all_ok;
}

// Reduced solve (S + D_c) y = g_c - sum Z u by the whole group: right-looking block elimination in global memory (L2) with
// explicit 6x6 pivot inverses.  Per block column J ONE group barrier: every CTA inverts the pivot block itself (same
// bits), then the rows of the trailing blocks (I, K), I >= K > J, are strided over all threads of the group; a thread
// recomputes its scaled panel row P_I[p,:] = A_IJ[p,:] W_J (36 FMAs) rather than wait for it, the column-J blocks of S
// stay unscaled (so nobody reads what another thread overwrites), and the owner of block (I, J+1) stores the panel row
// to P and updates the right-hand side row.  Back-substitution without solves by one warp: y_J = u_J - sum P_IJ^T y_I.
__device__ bool wide_reduced_solve(WideCtx& c, const WideHdr& h, double inv_radius, double* Wsm, double* bcast, const int* tri, double* tail_sm) {
  const int tid = c.tid, lane = c.lane, warp = c.warp;
  const int Cf = h.Cf, n = h.n;
  for (int i = c.gt; i < n; i += c.gsize) {
    h.yc[i] = h.gc[i] - h.zu[i];
    h.ab[i] = 0.0;
    h.S[(size_t)i * n + i] += fmin(fmax(h.hd[i], 1e-6), 1e32) * inv_radius;
  }
  if (tid == 0) bcast[0] = 0.0;
  wide_sync(c);
#ifdef SLSLAM_WIDE_PHASES
  long long sp[4] = {0, 0, 0, 0}, st0 = clock64();
#define SPHASE(i) { const long long now_ = clock64(); sp[i] += now_ - st0; st0 = now_; }
#else
#define SPHASE(i)
#endif
  // the first c0 block columns by the whole group in global memory, the rest (<= WIDE_TAIL_MAX) by CTA 0 in shared memory
  const int c0 = Cf > WIDE_TAIL_MAX ? Cf - WIDE_TAIL_MAX : 0;
  for (int J = 0; J < c0; ++J) {
    const int nb = Cf - J - 1;
    if (warp == 0) {
      const bool pok = wide_pivot_inverse(h.S + (size_t)(6 * J) * n + 6 * J, (size_t)n, Wsm, lane);
      if (lane == 0 && !pok) bcast[0] = 1.0;
    }
    __syncthreads();
    SPHASE(0)
    const double* bJ = h.yc + 6 * J;
    const int nrow = nb * (nb + 1) / 2 * 6;
    for (int e = c.gt; e < nrow + 6; e += c.gsize) {
      if (e < nrow) {
        const int blk = e / 6, p = e - 6 * blk;
        const int bi = tri[blk] >> 8, bk = tri[blk] & 0xff;
        const int I = J + 1 + bi, K = J + 1 + bk;
        const double* ai = h.S + (size_t)(6 * I + p) * n + 6 * J;
        double* dst = h.S + (size_t)(6 * I + p) * n + 6 * K;
        // every load first (one L2 round trip for all of them), then the arithmetic, then the stores: the arrays may alias
        // as far as the compiler knows, and a store in between would serialise the round trips
        double av[6], pv[6], o[6], akv[36], bv[6], ycv = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { av[k] = ai[k]; o[k] = dst[k]; }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double* ak = h.S + (size_t)(6 * K + q) * n + 6 * J;      // row q of the unscaled block (K, J)
#pragma unroll
          for (int k = 0; k < 6; ++k) akv[6 * q + k] = ak[k];
        }
        if (bk == 0) {
#pragma unroll
          for (int k = 0; k < 6; ++k) bv[k] = bJ[k];
          ycv = h.yc[6 * I + p];
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) pv[q] = (av[0] * Wsm[q] + av[1] * Wsm[6 + q] + av[2] * Wsm[12 + q]) + (av[3] * Wsm[18 + q] + av[4] * Wsm[24 + q] + av[5] * Wsm[30 + q]);
#pragma unroll
        for (int q = 0; q < 6; ++q)
          o[q] -= (pv[0] * akv[6 * q] + pv[1] * akv[6 * q + 1] + pv[2] * akv[6 * q + 2]) + (pv[3] * akv[6 * q + 3] + pv[4] * akv[6 * q + 4] + pv[5] * akv[6 * q + 5]);
#pragma unroll
        for (int q = 0; q < 6; ++q) dst[q] = o[q];
        if (bk == 0) {
          double* pdst = h.P + (size_t)(6 * I + p) * n + 6 * J;
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) { pdst[k] = pv[k]; sacc += pv[k] * bv[k]; }
          h.yc[6 * I + p] = ycv - sacc;
        }
      } else {
        const int q = e - nrow;                      // u_J = W_J b_J
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += bJ[k] * Wsm[6 * k + q];
        h.ub[6 * J + q] = sacc;
      }
    }
    SPHASE(1)
    wide_sync(c);
    SPHASE(2)
  }
#ifdef SLSLAM_WIDE_PHASES
  if (c.gt == 0) printf("  solve: inverse %lld  items %lld  barrier %lld (cycles, %d columns)\n", sp[0], sp[1], sp[2], Cf);
#endif
#undef SPHASE
  if (c.rank == 0) {
    double* T = tail_sm;                                             // [TAIL (TAIL + 1) / 2][36]
    double* pan = T + WIDE_TAIL_MAX * (WIDE_TAIL_MAX + 1) / 2 * 36;    // [TAIL - 1][36]
    double* ycs = pan + (WIDE_TAIL_MAX - 1) * 36;                    // [6 TAIL]
    double* ubs = ycs + 6 * WIDE_TAIL_MAX;                           // [6 TAIL]
    const bool tok = wide_solve_tail(h, c0, T, pan, ycs, ubs, Wsm, tri);
    if (tid == 0) *h.flag = (tok && bcast[0] == 0.0) ? 0.0 : 1.0;
  }
  wide_sync(c);
  const bool ok = c.G == 1 ? (*h.flag == 0.0) : (__ldcg(h.flag) == 0.0);
  // back-substitution by CTA 0: y_J = u_J - acc_J, acc_K += P_JK^T y_J for K < J.  u and acc live in shared memory
  // meanwhile; thread e owns entry (K, q) = (e / 6, e % 6) (and e + WIDE_NT), and the P values of the next column are
  // loaded while the current one is being applied, so the dependent chain runs through shared memory only
  if (c.rank == 0 && ok) {
    double* ubs = c.sh;              // [n]  (the reduction scratch is free here: 2 n <= 12 * WIDE_MAX_FREE <= WIDE_NT * WIDE_NPART)
    double* abs_ = c.sh + n;         // [n]
    for (int i = tid; i < n; i += WIDE_NT) { ubs[i] = h.ub[i]; abs_[i] = 0.0; }
    const int e0 = tid, e1 = tid + WIDE_NT;
    const int K0 = e0 / 6, q0 = e0 - 6 * K0, K1 = e1 / 6, q1 = e1 - 6 * K1;
    double pa[6], pb[6];
    auto fetch = [&](int J) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        pa[k] = (J >= 0 && e0 < 6 * J) ? h.P[(size_t)(6 * J + k) * n + 6 * K0 + q0] : 0.0;
        pb[k] = (J >= 0 && e1 < 6 * J) ? h.P[(size_t)(6 * J + k) * n + 6 * K1 + q1] : 0.0;
      }
    };
    fetch(Cf - 1);
    __syncthreads();
    for (int J = Cf - 1; J >= 0; --J) {
      double y[6], ca[6], cb[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { y[k] = ubs[6 * J + k] - abs_[6 * J + k]; ca[k] = pa[k]; cb[k] = pb[k]; }
      fetch(J - 1);
      __syncthreads();               // everybody has read u_J and acc_J
      if (tid < 6) ubs[6 * J + tid] = (tid == 0 ? y[0] : tid == 1 ? y[1] : tid == 2 ? y[2] : tid == 3 ? y[3] : tid == 4 ? y[4] : y[5]);
      if (e0 < 6 * J) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += ca[k] * y[k];
        abs_[e0] += sacc;
      }
      if (e1 < 6 * J) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += cb[k] * y[k];
        abs_[e1] += sacc;
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += WIDE_NT) h.yc[i] = ubs[i];     // y over u
  }
  wide_sync(c);
  return ok;
}

// grid = (windows resident at a time) x G CTAs; window w of a launch is solved by the group blockIdx.x / G, then w +
// groups, ... (every window has its own barrier counter and partial-sum slots)
__global__ void __launch_bounds__(WIDE_NT, 1) lba_wide_kernel(const WideHdr* __restrict__ hdrs, int nwin, int G) {
  extern __shared__ __align__(16) double wsm[];
  double* Wsm = wsm + WIDE_NPART * WIDE_NT;       // [36] pivot inverse
  double* bcast = Wsm + 36;                       // [16] broadcast scalars
  int* tri = reinterpret_cast<int*>(bcast + 16);  // block of a lower block triangle -> row << 8 | column
  double* tail_sm = bcast + 16 + (WIDE_MAX_FREE * (WIDE_MAX_FREE + 1) / 2 + 1) / 2;   // trailing matrix of the reduced solve
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
  c.gw = c.warp * G + c.rank; c.gwarps = G * (WIDE_NT / 32);      // interleaved: consecutive warp-items land on different SMs
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
#ifdef SLSLAM_WIDE_PHASES
  {
    const long long b0 = clock64();
    for (int k = 0; k < 20; ++k) wide_sync(c);
    if (c.gt == 0) printf("wide_sync: %lld cycles each (G = %d)\n", (clock64() - b0) / 20, c.G);
  }
  long long wph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wt0 = clock64();
#define WPHASE(i) { const long long now_ = clock64(); wph[i] += now_ - wt0; wt0 = now_; }
#else
#define WPHASE(i)
#endif
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
    WPHASE(0)
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
        double J[16], r[4];
        ld_row<16>(h.Jl + 16 * (size_t)s, J);
        ld_row<4>(h.r + 4 * (size_t)s, r);
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
        const double* Jcp = h.Jc + 24 * (size_t)s;
        const double* Jlp = h.Jl + 16 * (size_t)s;
        double* Zo = h.Z + 24 * (size_t)s;
        double Jc[24], Jl[16], Zv[24];           // loads first, stores last (see the reduced solve)
        ld_row<24>(Jcp, Jc);
        ld_row<16>(Jlp, Jl);
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          double W[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) W[q] = Jc[p] * Jl[q] + Jc[6 + p] * Jl[4 + q] + Jc[12 + p] * Jl[8 + q] + Jc[18 + p] * Jl[12 + q];
          const double z0 = W[0] * inv[0];
          const double z1 = (W[1] - z0 * Lm[1]) * inv[1];
          const double z2 = (W[2] - z0 * Lm[3] - z1 * Lm[4]) * inv[2];
          const double z3 = (W[3] - z0 * Lm[6] - z1 * Lm[7] - z2 * Lm[8]) * inv[3];
          Zv[4 * p] = z0; Zv[4 * p + 1] = z1; Zv[4 * p + 2] = z2; Zv[4 * p + 3] = z3;
        }
        st_row<24>(Zo, Zv);
      }
    }
    // constant lines: their observations have Jl = 0, no Schur term: Z = 0
    for (int s = c.gt; s < N; s += c.gsize) {
      if (!h.line_free[h.line_s[s]]) { double2* Zo = reinterpret_cast<double2*>(h.Z + 24 * (size_t)s); for (int k = 0; k < 12; ++k) Zo[k] = make_double2(0.0, 0.0); }
    }
    bool line_fail;
    {
      double v[2] = {failf, gm};
      wide_group_reduce<2>(c, h, v, 3u);          // (publishes Z and lineLU)
      line_fail = v[0] != 0.0; gm = v[1];
    }
    WPHASE(1)
    // ---- cameras: diagonal blocks, g_c, Z u, diag H_cc.  Warp per free camera, lanes over its observations, transposing
    // warp reduction (lane k ends up with accumulator k) ----
    for (int f = c.gw; f < Cf; f += c.gwarps) {
      double acc[39];
#pragma unroll
      for (int k = 0; k < 39; ++k) acc[k] = 0.0;
      for (int t = h.cam_start[f] + lane; t < h.cam_start[f + 1]; t += 32) {
        const int s = h.cam_obs[t];
        const double* Jcp = h.Jc + 24 * (size_t)s;
        const double* Zp = h.Z + 24 * (size_t)s;
        const double* rp = h.r + 4 * (size_t)s;
        double Jc[24], Z[24], r[4], u[4];
        ld_row<24>(Jcp, Jc);
        ld_row<24>(Zp, Z);
        ld_row<4>(rp, r);
        const int l = h.line_s[s];
        const bool lf = h.line_free[l] != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = lf ? h.lineLU[22 * (size_t)l + 10 + k] : 0.0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
#pragma unroll
          for (int q = 0; q <= p; ++q)
            acc[p * (p + 1) / 2 + q] += Jc[p] * Jc[q] + Jc[6 + p] * Jc[6 + q] + Jc[12 + p] * Jc[12 + q] + Jc[18 + p] * Jc[18 + q]
                                        - (Z[4 * p] * Z[4 * q] + Z[4 * p + 1] * Z[4 * q + 1] + Z[4 * p + 2] * Z[4 * q + 2] + Z[4 * p + 3] * Z[4 * q + 3]);
          acc[21 + p] += Jc[p] * r[0] + Jc[6 + p] * r[1] + Jc[12 + p] * r[2] + Jc[18 + p] * r[3];
          acc[27 + p] += Z[4 * p] * u[0] + Z[4 * p + 1] * u[1] + Z[4 * p + 2] * u[2] + Z[4 * p + 3] * u[3];
          acc[33 + p] += Jc[p] * Jc[p] + Jc[6 + p] * Jc[6 + p] + Jc[12 + p] * Jc[12 + p] + Jc[18 + p] * Jc[18 + p];
        }
      }
      double tail[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) tail[k] = warp_sum(acc[32 + k]);
      warp_reduce_scatter32(acc, lane);
      const double tot = acc[0];                 // accumulator `lane`
      if (lane < 21) {
        int p = 0; while ((p + 1) * (p + 2) / 2 <= lane) ++p;
        const int q = lane - p * (p + 1) / 2;
        h.S[(size_t)(6 * f + p) * n + 6 * f + q] = tot; h.S[(size_t)(6 * f + q) * n + 6 * f + p] = tot;
      } else if (lane < 27) h.gc[6 * f + lane - 21] = tot;
      else h.zu[6 * f + lane - 27] = tot;        // lanes 27..31: entries 0..4
      if (lane == 0) {
        h.zu[6 * f + 5] = tail[0];
#pragma unroll
        for (int k = 0; k < 6; ++k) h.hd[6 * f + k] = tail[1 + k];
      }
    }
    // ---- pairs: off-diagonal blocks (I > K).  Warp per block, lanes over the lines (lookup table line x free camera) ----
    if (!last) {
      const int nblk = Cf * (Cf - 1) / 2;
      for (int b = c.gw; b < nblk; b += c.gwarps) {
        const int I = (tri[b] >> 8) + 1, K = tri[b] & 0xff;      // b = I (I - 1) / 2 + K, K < I
        double acc[36];
#pragma unroll
        for (int k = 0; k < 36; ++k) acc[k] = 0.0;
        for (int l = lane; l < L; l += 32) {
          const int sa = h.pos[(size_t)l * Cf + I], sb = h.pos[(size_t)l * Cf + K];
          if (sa < 0 || sb < 0) continue;
          const double* Zap = h.Z + 24 * (size_t)sa;
          const double* Zbp = h.Z + 24 * (size_t)sb;
          double Za[24], Zb[24];
          ld_row<24>(Zap, Za);
          ld_row<24>(Zbp, Zb);
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q)
              acc[6 * p + q] += Za[4 * p] * Zb[4 * q] + Za[4 * p + 1] * Zb[4 * q + 1] + Za[4 * p + 2] * Zb[4 * q + 2] + Za[4 * p + 3] * Zb[4 * q + 3];
        }
        double tail[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tail[k] = warp_sum(acc[32 + k]);
        warp_reduce_scatter32(acc, lane);
        {
          const int p = lane / 6, q = lane - 6 * p;
          h.S[(size_t)(6 * I + p) * n + 6 * K + q] = -acc[0];
        }
        if (lane < 4) h.S[(size_t)(6 * I + 5) * n + 6 * K + 2 + lane] = -(lane == 0 ? tail[0] : lane == 1 ? tail[1] : lane == 2 ? tail[2] : tail[3]);
      }
    }
    wide_sync(c);
    WPHASE(2)
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
    WPHASE(3)
    // ---- reduced solve by the whole group (the pivot test is evaluated identically on every CTA) ----
    const bool sok = wide_reduced_solve(c, h, inv_radius, Wsm, bcast, tri, tail_sm);
    WPHASE(4)
    bool ok = !line_fail && sok;
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
            double Z[24];
            ld_row<24>(h.Z + 24 * (size_t)s, Z);
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
      WPHASE(5)
      wide_sweep<2>(c, h, camRt, ltrigt, &pc, &pf);
      {
        double v[1] = {pc};
        wide_group_reduce<1>(c, h, v, 0u);
        new_cost = v[0];
      }
      WPHASE(6)
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
#ifdef SLSLAM_WIDE_PHASES
  if (c.gt == 0) printf("wide phases (cycles, %d iterations, G=%d): linearise %lld  lines %lld  cameras+pairs %lld  gradient %lld  solve %lld  trial point %lld  trial sweep %lld\n",
                        iters, c.G, wph[0], wph[1], wph[2], wph[3], wph[4], wph[5], wph[6]);
#endif
#undef WPHASE
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
