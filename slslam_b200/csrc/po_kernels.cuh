// Pose-graph optimisation on device: what ceres::Solve does for a POProblem (reference src/slam.cpp:1283-1293,
// src/po_problem.cpp:40-77; LM semantics restated in SURVEY.md App. A3).
//
// The whole Levenberg-Marquardt loop is enqueued on one stream with NO host round trip: the trust-region state lives
// in a PoState record in HBM, the accept / reject / terminate decision is a one-CTA kernel, and every other kernel
// starts by reading `done` and returns at once after termination.
//   K5  po_linearize      thread per edge: residual + dual-number Jacobians (two 6x6 per edge)
//       po_colnorm_grad   thread per unknown: column norms and gradient over the pose's incident edges (CSR, fixed order)
//   K6  po_assemble       thread per entry of each non-zero 6x6 block of J^T J (+ LM diagonal), dense lower storage,
//                         right-hand side appended as row n so that the factorisation also forward-substitutes it
//       po_chol_panel     32-wide panel: diagonal block factorised in shared memory by every CTA, rows below solved
//       po_chol_syrk      trailing update, 64x64 register-tiled
//       po_backsolve      L^T y = z, right-looking, one CTA
//       po_step / po_decide / po_accept / po_refresh   trial point, model decrease, step acceptance, termination
//   K6' block-sparse path (the default): the pose graph's J^T J is block sparse (odometry band + loop edges) and the
//       reference factors it sparsely (SPARSE_NORMAL_CHOLESKY, po_problem.cpp:68).  The host orders the free poses by
//       minimum degree and lays out the 6x6 blocks of L (fill included) with, per column, the list of (source, source,
//       destination) block updates; po_sp_assemble fills the blocks and po_sp_factor_solve -- ONE CTA, one launch --
//       runs the whole right-looking block factorisation, carries the right-hand side along and back-substitutes.
//       The dense kernels above remain as the path for graphs whose factor is nearly full.
// Every reduction has a fixed order (no atomics): results are bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slslam_b200.h"
#include "po_math.cuh"
#include "lba_math.cuh"

namespace slslam {

constexpr int PO_NB = 32;      // panel width of the dense factorisation
constexpr int PO_TR = 128;     // rows per CTA in the panel solve
constexpr int PO_TS = 64;      // trailing-update tile

struct PoState {
  double cost, new_cost, fixed_cost, initial_cost, radius, decrease_factor, gmax, gtol_abs, x_norm;
  double ftol, gtol, ptol;
  int done, term, successful, unsuccessful, invalid, iters, it, max_iters, cur, chol_fail, accepted, first;
};

struct PoDev {
  int K, E, n, M, ld, nblk;       // poses, edges, unknowns (6*free poses), M = n+1 rows, leading dimension, H blocks
  const int *idx1, *idx2;         // [E]
  const double* cons;             // [6E]
  const int* slot;                // [K] reduced block index or -1
  const int* slot_pose;           // [n/6] pose of each reduced block
  const unsigned char* active;    // [E] edge has a free pose
  const int* inc_off;             // [n/6 + 1] CSR of incident edges per reduced block
  const int* inc;                 // edge << 1 | (1 if the block is pose2 of the edge)
  const int *blk_i, *blk_j;       // [nblk] block coordinates, blk_i >= blk_j
  const int* blk_off;             // [nblk + 1]
  const int* contrib;             // edge << 1 | (1 if blk_i is pose2 of the edge)
  double *x, *xt;                 // [6K]
  double *r, *J1, *J2, *cost_e;   // two sets each: [2][6E], [2][36E], [2][36E], [2][E]
  double *scale, *cn, *g, *y, *mval;   // [n], [n], [n], [n], [E]
  double *H, *Ld;                 // [M][ld] dense lower + rhs row; [ceil(n/32)][32][32] factored diagonal blocks
  PoState* st;
  double* trace;                  // [max_iters][SLSLAM_TRACE_WIDTH]
  slslam_summary* summary;
  // ---- block-sparse factorisation (sparse != 0); `pos` = position of a reduced block in the elimination order ----
  int sparse, Kf, nsb;            // nsb = blocks of L: Kf diagonal blocks (block id = position) then the off-diagonal ones
  double* Hb;                     // [nsb][36] row-major 6x6; off-diagonal block (row pos > col pos): rows of the row block
  double *bz, *us, *yp;           // [n] in position order: right-hand side, u = W b (see the kernel), solution
  const int* slot_pos;            // [Kf] position of reduced block (slot) s
  const int* col_off;             // [Kf + 1] first off-diagonal block of column c (block id = Kf + col_off[c] + a)
  const int* row_pos;             // [nsb - Kf] row position of every off-diagonal block, ascending inside a column
  const int* tri_off;             // [Kf + 1] updates of column c
  const int2* tri;                // x: destination block id, y: a | b << 16 (sources: off-diagonal blocks a >= b of the column)
  const int* blk_dst;             // [nblk] where po_sp_assemble writes the structurally non-zero blocks of J^T J
  const int* bs_chunk;            // [bs_nchunk + 1] descending column boundaries of the back-substitution's staging chunks
  int bs_nchunk;
  long long* sp_cycles;           // [4] diagnostics of po_sp_factor_solve: phase 1, phase 2, back-substitution, total (SM cycles)
  const int* stage_off;           // level order: columns [stage_off[s], stage_off[s + 1]) are eliminated side by side
  int nstage;
};

// ------------------------------------------------------------------------------------------------------------
// K5: residual and Jacobians of every edge at x (which_x = 0) or at the trial point (1), into buffer set `use_cur ?
// cur : 1 - cur`.  all_free = 1 ignores the constant pose (evaluate-only entry point).
// ------------------------------------------------------------------------------------------------------------
// Four threads per edge (round 2; one thread carried all 12 partials in two passes before: 38 us per call, twice per LM
// iteration): thread `part` evaluates the residual with the three partials (part & 1) * 3 .. + 2 of pose1 (part < 2) or of
// pose2 (part >= 2).  Every partial is computed by the same operations as before, so the Jacobians keep their bits.
constexpr int PO_LIN_PARTS = 4;
__global__ void po_linearize(PoDev d, int which_x, int use_cur, int all_free) {
  if (d.st->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t / PO_LIN_PARTS, part = t - e * PO_LIN_PARTS;
  if (e >= d.E) return;
  const int set = use_cur ? d.st->cur : 1 - d.st->cur;
  const double* x = which_x ? d.xt : d.x;
  const int a = d.idx1[e], b = d.idx2[e];
  const bool f1 = all_free || d.slot[a] >= 0, f2 = all_free || d.slot[b] >= 0;
  double* r = d.r + ((size_t)set * d.E + e) * 6;
  double* J1 = d.J1 + ((size_t)set * d.E + e) * 36;
  double* J2 = d.J2 + ((size_t)set * d.E + e) * 36;
  double p1[6], p2[6], c[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { p1[k] = x[6 * a + k]; p2[k] = x[6 * b + k]; c[k] = d.cons[6 * (size_t)e + k]; }
  const int d0 = 3 * (part & 1);            // first direction of this thread
  const bool second = part >= 2;            // derivatives with respect to pose2
  if (second && (a == b || !f2)) {
    // (a self edge aliases one block for both arguments: its directions coincide and go to J1)
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int j = 0; j < 3; ++j) J2[6 * k + d0 + j] = 0.0;
    return;
  }
  typedef Dual<3> D3;
  D3 A[6], B[6], res[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const bool mine = k >= d0 && k < d0 + 3;
    const bool va = mine && !second, vb = mine && (second || a == b);
    A[k] = va ? dvar<3>(p1[k], k - d0) : dconst<3>(p1[k]);
    B[k] = vb ? dvar<3>(p2[k], k - d0) : dconst<3>(p2[k]);
  }
  pose_constraint_residual<3>(A, B, c, res);
  if (part == 0) {
    double cost = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { r[k] = res[k].a; cost += 0.5 * res[k].a * res[k].a; }
    d.cost_e[(size_t)set * d.E + e] = cost;
  }
  double* J = second ? J2 : J1;
  const bool fr = second ? true : f1;
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[6 * k + d0 + j] = fr ? res[k].v[j] : 0.0;
}

// squared column norms and gradient of the unscaled Jacobian (current set), thread per unknown
__global__ void po_colnorm_grad(PoDev d, int only_if_accepted) {
  if (d.st->done || (only_if_accepted && !d.st->accepted)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  const int s = i / 6, j = i % 6, set = d.st->cur;
  double cn = 0.0, g = 0.0;
  for (int t = d.inc_off[s]; t < d.inc_off[s + 1]; ++t) {
    const int code = d.inc[t], e = code >> 1;
    const double* J = ((code & 1) ? d.J2 : d.J1) + ((size_t)set * d.E + e) * 36;
    const double* r = d.r + ((size_t)set * d.E + e) * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double v = J[6 * k + j]; cn += v * v; g += v * r[k]; }
  }
  d.cn[i] = cn; d.g[i] = g;
}

// block-wide sum / max in a fixed order; result valid on every thread
template <bool MAX>
__device__ double po_block_reduce(double v, double* sh) {
  const int tid = threadIdx.x, nt = blockDim.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = nt >> 1; s > 0; s >>= 1) {
    if (tid < s) sh[tid] = MAX ? fmax(sh[tid], sh[tid + s]) : sh[tid] + sh[tid + s];
    __syncthreads();
  }
  const double out = sh[0];
  __syncthreads();
  return out;
}

// After a (re-)linearisation: cost sums, Jacobi scaling (first call), gradient max norm, |x|, gradient test.
__global__ void po_refresh(PoDev d, int first) {
  __shared__ double sh[256];
  PoState* st = d.st;
  if (st->done || (!first && !st->accepted)) return;
  const int tid = threadIdx.x, set = st->cur;
  double c_act = 0.0, c_fix = 0.0, gm = 0.0, xn = 0.0;
  if (first) {
    for (int e = tid; e < d.E; e += 256) { if (d.active[e]) c_act += d.cost_e[(size_t)set * d.E + e]; else c_fix += d.cost_e[(size_t)set * d.E + e]; }
    for (int i = tid; i < d.n; i += 256) d.scale[i] = 1.0 / (1.0 + sqrt(d.cn[i]));
  }
  for (int i = tid; i < d.n; i += 256) {
    gm = fmax(gm, fabs(d.g[i]));
    const double xv = d.x[6 * d.slot_pose[i / 6] + i % 6];
    xn += xv * xv;
  }
  gm = po_block_reduce<true>(gm, sh);
  xn = po_block_reduce<false>(xn, sh);
  if (first) { c_act = po_block_reduce<false>(c_act, sh); c_fix = po_block_reduce<false>(c_fix, sh); }
  if (tid == 0) {
    st->gmax = gm; st->x_norm = sqrt(xn);
    if (first) {
      st->cost = c_act; st->fixed_cost = c_fix; st->initial_cost = c_act + c_fix;
      st->gtol_abs = st->gtol * fmax(gm, 2.220446049250313e-16);
      if (d.n == 0 || gm <= st->gtol_abs) { st->term = SLSLAM_GRADIENT_TOLERANCE; st->done = 1; }
    } else if (gm <= st->gtol_abs) {
      st->term = SLSLAM_GRADIENT_TOLERANCE; st->done = 1;
    }
    st->accepted = 0;
  }
}

// K6a: scaled normal equations.  One thread per entry of every structurally non-zero block; the right-hand side
// J^T r goes to row n.  H must have been zeroed.
__global__ void po_assemble(PoDev d) {
  const PoState* st = d.st;
  if (st->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int set = st->cur;
  if (t < d.nblk * 36) {
    const int b = t / 36, p = (t % 36) / 6, q = t % 6;
    const int bi = d.blk_i[b], bj = d.blk_j[b];
    double s = 0.0;
    for (int u = d.blk_off[b]; u < d.blk_off[b + 1]; ++u) {
      const int code = d.contrib[u], e = code >> 1;
      const double* Ja = ((code & 1) ? d.J2 : d.J1) + ((size_t)set * d.E + e) * 36;
      const double* Jb = (bi == bj) ? Ja : (((code & 1) ? d.J1 : d.J2) + ((size_t)set * d.E + e) * 36);
#pragma unroll
      for (int k = 0; k < 6; ++k) s += Ja[6 * k + p] * Jb[6 * k + q];
    }
    const int row = 6 * bi + p, col = 6 * bj + q;
    s *= d.scale[row] * d.scale[col];
    if (row == col) {
      const double sc = d.scale[row];
      s += fmin(fmax(d.cn[row] * sc * sc, 1e-6), 1e32) / st->radius;
    }
    if (col <= row) d.H[(size_t)row * d.ld + col] = s;
  } else {
    const int i = t - d.nblk * 36;
    if (i < d.n) d.H[(size_t)d.n * d.ld + i] = d.g[i] * d.scale[i];
  }
}

// K6b: panel [k0, k0+nb).  Every CTA factors the diagonal block in shared memory (same data, same order, same bits);
// CTA 0 stores it to Ld, CTA b >= 1 solves PO_TR rows below it: X L_kk^T = A.
__global__ void __launch_bounds__(PO_TR) po_chol_panel(PoDev d, int k0) {
  if (d.st->done) return;
  __shared__ double Ls[PO_NB][PO_NB + 1];
  __shared__ double inv[PO_NB];
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int nb = min(PO_NB, d.n - k0);
  if (tid < 32) {
    // warp 0 factors the block in registers: lane i holds row i; column k needs one broadcast of a_kk and one shuffle
    // per trailing column, no block-wide barrier on the pivot chain
    const int i = tid;
    double a[PO_NB];
#pragma unroll
    for (int j = 0; j < PO_NB; ++j) {
      double v = (i == j) ? 1.0 : 0.0;                     // identity padding beyond nb
      if (i < nb && j < nb && j <= i) v = d.H[(size_t)(k0 + i) * d.ld + k0 + j];
      a[j] = v;
    }
    bool notpd = false;
    double myinv = 1.0;
#pragma unroll
    for (int k = 0; k < PO_NB; ++k) {
      const double dkk = __shfl_sync(0xffffffffu, a[k], k);
      notpd = notpd || !(dkk > 0.0) || !isfinite(dkk);
      const double ik = rsqrt(dkk);
      const double lik = a[k] * ik;                        // lane k: l_kk = d_kk / sqrt(d_kk)
      a[k] = lik;
      if (i == k) myinv = ik;
#pragma unroll
      for (int j = k + 1; j < PO_NB; ++j) {
        const double ljk = __shfl_sync(0xffffffffu, lik, j);
        a[j] -= lik * ljk;
      }
    }
#pragma unroll
    for (int j = 0; j < PO_NB; ++j) Ls[i][j] = (j <= i) ? a[j] : 0.0;
    inv[i] = myinv;
    if (i == 0) bad = notpd ? 1 : 0;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    double* out = d.Ld + (size_t)(k0 / PO_NB) * PO_NB * PO_NB;
    for (int t = tid; t < PO_NB * PO_NB; t += PO_TR) out[t] = Ls[t / PO_NB][t % PO_NB];
    if (tid == 0 && bad) d.st->chol_fail = 1;
    return;
  }
  const int row = k0 + nb + (blockIdx.x - 1) * PO_TR + tid;
  if (row >= d.M) return;
  double* a = d.H + (size_t)row * d.ld + k0;
  double x[PO_NB];
#pragma unroll
  for (int q = 0; q < PO_NB; ++q) x[q] = (q < nb) ? a[q] : 0.0;
#pragma unroll
  for (int q = 0; q < PO_NB; ++q) {
    x[q] *= inv[q];
#pragma unroll
    for (int m = q + 1; m < PO_NB; ++m) x[m] -= x[q] * Ls[m][q];
  }
#pragma unroll
  for (int q = 0; q < PO_NB; ++q) if (q < nb) a[q] = x[q];
}

// K6c: trailing update A_ij -= sum_k L_ik L_jk over the panel columns, lower tiles only, rhs row included.
__global__ void __launch_bounds__(256) po_chol_syrk(PoDev d, int k0, int nb, int t0, int ntile) {
  if (d.st->done) return;
  __shared__ double Pi[PO_NB][PO_TS + 2];
  __shared__ double Pj[PO_NB][PO_TS + 2];
  // tile (ti, tj), tj <= ti, from the linear block index
  int ti = 0, rem = blockIdx.x;
  while (rem > ti) { rem -= ti + 1; ++ti; }
  const int tj = rem;
  (void)ntile;
  const int tid = threadIdx.x;
  const int r0 = t0 + ti * PO_TS, c0 = t0 + tj * PO_TS;
  for (int t = tid; t < PO_TS * PO_NB; t += 256) {
    const int rr = t / PO_NB, k = t % PO_NB;
    const int gi = r0 + rr, gj = c0 + rr;
    Pi[k][rr] = (gi < d.M && k < nb) ? d.H[(size_t)gi * d.ld + k0 + k] : 0.0;
    Pj[k][rr] = (gj < d.M && k < nb) ? d.H[(size_t)gj * d.ld + k0 + k] : 0.0;
  }
  __syncthreads();
  const int ty = tid / 16, tx = tid % 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
  for (int k = 0; k < PO_NB; ++k) {
    double ai[4], bj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ai[i] = Pi[k][ty * 4 + i]; bj[i] = Pj[k][tx * 4 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += ai[i] * bj[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = r0 + ty * 4 + i;
    if (gi >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = c0 + tx * 4 + j;
      if (gj <= gi && gj < d.n) d.H[(size_t)gi * d.ld + gj] -= acc[i][j];
    }
  }
}

// L^T y = z with z = row n of H (the forward-substituted right-hand side), right-looking over 32-blocks from the last,
// spread over several co-resident CTAs (cooperative launch): block kb belongs to CTA kb mod gridDim.  The owner solves
// its diagonal block from Ld (warp 0, shared memory, reciprocal diagonal), publishes y_kb and raises flag[kb]; every
// CTA then removes y_kb's contribution from the blocks it owns (lane = column, the 32 rows of the block are independent
// coalesced loads, issued BEFORE the wait so only the flag and 32 doubles are on the critical path).  One CTA did all
// of this before: 0.73 ms per call, L2-latency bound on a single SM.
constexpr int PO_BS_MAXOWN = 8;    // 32-column blocks a CTA may own (n <= 32 * 8 * gridDim)

__global__ void __launch_bounds__(256) po_backsolve(PoDev d, unsigned int* flags, unsigned int gen) {
  if (d.st->done) return;
  __shared__ double yk[PO_NB];
  __shared__ double Lsm[PO_NB][PO_NB + 1];
  __shared__ double invd[PO_NB];
  __shared__ double wown[PO_BS_MAXOWN][PO_NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = d.n;
  const int nblocks = (n + PO_NB - 1) / PO_NB, ncta = gridDim.x, me = blockIdx.x;
  // my blocks: me, me + ncta, ...  (slot s <-> block me + s * ncta); start from z
  for (int t = tid; t < PO_BS_MAXOWN * PO_NB; t += 256) {
    const int sidx = t / PO_NB, j = (me + sidx * ncta) * PO_NB + t % PO_NB;
    wown[sidx][t % PO_NB] = (me + sidx * ncta < nblocks && j < n) ? d.H[(size_t)n * d.ld + j] : 0.0;
  }
  __syncthreads();
  for (int kb = nblocks - 1; kb >= 0; --kb) {
    const int k0 = kb * PO_NB, nb = min(PO_NB, n - k0);
    const bool mine = (kb % ncta) == me;
    // prefetch the rows of block kb over the columns of the blocks I own below it: warp w handles owned slot w
    double hrow[PO_NB];
    const int myblk = me + warp * ncta;
    const bool upd = warp < PO_BS_MAXOWN && myblk < kb;
    if (upd) {
      const int j = myblk * PO_NB + lane;
#pragma unroll
      for (int q = 0; q < PO_NB; ++q) hrow[q] = (q < nb) ? __ldcg(d.H + (size_t)(k0 + q) * d.ld + j) : 0.0;
    }
    if (mine) {
      const double* L = d.Ld + (size_t)kb * PO_NB * PO_NB;
      for (int t = tid; t < PO_NB * PO_NB; t += 256) Lsm[t / PO_NB][t % PO_NB] = L[t];
      __syncthreads();
      if (tid < 32) invd[tid] = 1.0 / Lsm[tid][tid];
      __syncthreads();
      if (tid < 32) {
        double wv = (tid < nb) ? wown[kb / ncta][tid] : 0.0;
        for (int q = nb - 1; q >= 0; --q) {
          const double yq = __shfl_sync(0xffffffffu, wv, q) * invd[q];
          if (tid == q) wv = yq;
          else if (tid < q) wv -= Lsm[q][tid] * yq;
        }
        yk[tid] = (tid < nb) ? wv : 0.0;
        if (tid < nb) { d.y[k0 + tid] = wv; __threadfence(); }
        __syncwarp();
        if (tid == 0) {
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + kb), "r"(gen) : "memory");
        }
      }
      __syncthreads();
    } else {
      if (tid == 0) {
        unsigned int seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + kb) : "memory"); } while (seen != gen);
      }
      __syncthreads();
      if (tid < 32) yk[tid] = (tid < nb) ? __ldcg(d.y + k0 + tid) : 0.0;
      __syncthreads();
    }
    if (upd) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q = 0; q < PO_NB; q += 2) { s0 += hrow[q] * yk[q]; s1 += hrow[q + 1] * yk[q + 1]; }
      wown[warp][lane] -= s0 + s1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// K6' block-sparse: assembly.  One thread per entry of every structurally non-zero block of J^T J (blk_i = row
// block, the later one in the elimination order; contribution lists as in po_assemble); Hb must have been zeroed (fill).
// ------------------------------------------------------------------------------------------------------------
__global__ void po_sp_assemble(PoDev d) {
  const PoState* st = d.st;
  if (st->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int set = st->cur;
  if (t < d.nblk * 36) {
    const int b = t / 36, p = (t % 36) / 6, q = t % 6;
    const int bi = d.blk_i[b], bj = d.blk_j[b];
    double s = 0.0;
    for (int u = d.blk_off[b]; u < d.blk_off[b + 1]; ++u) {
      const int code = d.contrib[u], e = code >> 1;
      const double* Ja = ((code & 1) ? d.J2 : d.J1) + ((size_t)set * d.E + e) * 36;
      const double* Jb = (bi == bj) ? Ja : (((code & 1) ? d.J1 : d.J2) + ((size_t)set * d.E + e) * 36);
#pragma unroll
      for (int k = 0; k < 6; ++k) s += Ja[6 * k + p] * Jb[6 * k + q];
    }
    const int row = 6 * bi + p, col = 6 * bj + q;
    s *= d.scale[row] * d.scale[col];
    if (row == col) {
      const double sc = d.scale[row];
      s += fmin(fmax(d.cn[row] * sc * sc, 1e-6), 1e32) / st->radius;
    }
    d.Hb[(size_t)d.blk_dst[b] * 36 + 6 * p + q] = s;
  } else {
    const int i = t - d.nblk * 36;
    if (i < d.n) d.bz[6 * d.slot_pos[i / 6] + i % 6] = d.g[i] * d.scale[i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// K6' block-sparse: factorisation + both substitutions in ONE launch of ONE CTA (the elimination order is a chain of
// dependent block columns; what parallelism a column has -- its panel rows and its block updates -- fits one CTA, and
// a single CTA keeps the blocks it touches in its L1 between consecutive columns).  Block elimination with explicit
// inverses of the 6x6 pivot blocks, as in the LBA reduced solve (lba_kernel.cuh): per column c
//   phase 1  every panel thread forms W = A_cc^-1 in registers (two closed-form 3x3 cofactor inverses + a 3x3 Schur
//            complement); thread (a, p): row p of P_a = A_ac W, original row kept in shared memory, P written over A_ac;
//            six threads: u_c = W b_c
//   phase 2  the planned updates  A_dst -= P_a A_bc^T  (thread per entry, 36 per update) and  b_row(a) -= P_a b_c
// then, descending, y_c = u_c - sum_a P_ac^T y_row(a) by warp 0.
// ------------------------------------------------------------------------------------------------------------
constexpr int PO_SP_NT = 512;
constexpr int PO_SP_MAXROWS = 160;   // off-diagonal blocks per column (the shared-memory column cache holds 1 + MAXROWS blocks)
constexpr int PO_SP_YMAX = 6144;     // unknowns whose solution vector is kept in shared memory during the back-substitution
constexpr int PO_SP_TRICAP = 1024;   // updates of one column staged in shared memory (longer lists are read from global memory)
constexpr int PO_SP_CACHE = (1 + PO_SP_MAXROWS) * 36;                       // doubles per column cache
constexpr int PO_SP_SMEM_DOUBLES = 2 * PO_SP_CACHE + PO_SP_MAXROWS * 36 + PO_SP_YMAX + 72 + 24 + 2 * PO_SP_TRICAP + PO_SP_MAXROWS + 400;
constexpr size_t PO_SP_SMEM = (size_t)PO_SP_SMEM_DOUBLES * 8;

__device__ __forceinline__ double po_ld_strong(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// W = A^-1 of a symmetric positive definite 6x6 block given by its lower triangle (row-major 6x6 source): two closed-form
// 3x3 cofactor inverses and a 3x3 Schur complement (one reciprocal each), as in the LBA reduced solve.
__device__ __forceinline__ bool po_spd6_inverse(const double* A36, double* W /* [21], lower, L6(p,q) */) {
#define L6I(p, q) ((p) * ((p) + 1) / 2 + (q))
#define SY3(mm, r, cc) mm[(r) <= (cc) ? ((r) == 0 ? (cc) : (r) == 1 ? 2 + (cc) : 5) : ((cc) == 0 ? (r) : (cc) == 1 ? 2 + (r) : 5)]
  double A[21];
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int q = 0; q <= p; ++q) A[L6I(p, q)] = A36[6 * p + q];
  double Ai[6], Si[6], M[9], S[6];
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
      SY3(S, r, cc) = A[L6I(3 + cc, 3 + r)] - (M[3 * r] * A[L6I(3 + cc, 0)] + M[3 * r + 1] * A[L6I(3 + cc, 1)] + M[3 * r + 2] * A[L6I(3 + cc, 2)]);
  ok = spd3_inverse(S[0], S[1], S[2], S[3], S[4], S[5], Si) && ok;
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
#undef L6I
  return ok;
}

// The dependent chain of the factorisation is, per block column c: P_c = A_c W_c (panel) -> the update of the NEXT
// pivot block -> its inverse W_{c+1}.  Everything on that chain lives in shared memory:
//   panel phase   threads (a, p): row p of P_a = A_ac W_c (W_c from shared memory), P written to the panel buffer and to
//                 global memory (the back-substitution needs it); six threads u_c = W_c b_c.  Meanwhile the idle warps
//                 fetch column c + 1 (pivot block, panel, right-hand side, its update list and row positions) into the
//                 second column cache -- every earlier update of it has been issued before the previous barrier.
//   update phase  warp 0 applies column c's update of the pivot block c + 1 first (the first entry of the list when it
//                 exists) and inverts it into W_{c+1} while warps 1..31 apply the other updates: into the cache when
//                 they target column c + 1, otherwise as fire-and-forget reductions (red.global.add.f64; nobody waits
//                 for L2, and the barriers between columns order the reductions of one address).
// Round 2, first version: every phase read and wrote global memory and every panel thread inverted the pivot itself,
// 752 us per launch for 260 poses; the phase counters (slslam_po_stats::factor_cycles) guided this form.
__global__ void __launch_bounds__(PO_SP_NT, 1) po_sp_factor_solve(PoDev d) {
  if (d.st->done) return;
  extern __shared__ __align__(16) double spsm[];
  double* cache0 = spsm;                                // column caches: [0..36) pivot block, then the panel blocks
  double* cache1 = cache0 + PO_SP_CACHE;
  double* Pm = cache1 + PO_SP_CACHE;                    // [MAXROWS][36] P = A W of the current column
  double* ysm = Pm + PO_SP_MAXROWS * 36;                // [YMAX] solution in position order (back-substitution)
  double* Wsm = ysm + PO_SP_YMAX;                       // [2][36] pivot inverse of the current / next column
  double* bzc = Wsm + 72;                               // [2][8] right-hand-side block of the current / next column
  double* bc = bzc + 16;                                // [8] b_c
  int2* tri_s = reinterpret_cast<int2*>(bc + 8);        // [2][TRICAP] staged update lists
  int* rp_s = reinterpret_cast<int*>(tri_s + 2 * PO_SP_TRICAP);   // [2][MAXROWS] staged row positions
  int* bs_i = rp_s + 2 * PO_SP_MAXROWS;                 // back-substitution: [2][400] ints (column offsets, row positions)
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, Kf = d.Kf;
  if (tid == 0) bad = 0;
  long long t_p1 = 0, t_p2 = 0, t_mark = 0, t_start = 0;
  if (tid == 0) { t_start = clock64(); t_mark = t_start; }
  // What the idle warps (or, for column 0, everybody) fetch for column cn into cache `dst`.  ONE concatenated index space
  // (blocks | right-hand side | update list | row positions): a thread issues its load(s) before it stores anything, so
  // the whole fetch is one L2 round trip instead of one per array (it sits on the barrier of the panel phase).
  auto fetch_column = [&](int cn, double* dst, double* bdst, int first_thread, int nthreads) {
    const int oa = d.col_off[cn], mm = d.col_off[cn + 1] - oa;
    const int ta = d.tri_off[cn], ntr = d.tri_off[cn + 1] - ta;
    const int nblk = 36 * (1 + mm), n1 = nblk + 6, n2 = n1 + (ntr <= PO_SP_TRICAP ? ntr : 0), n3 = n2 + mm;
    int2* ts = tri_s + (cn & 1) * PO_SP_TRICAP;
    int* rs = rp_s + (cn & 1) * PO_SP_MAXROWS;
    for (int i0 = tid - first_thread; i0 < n3; i0 += 2 * nthreads) {
      const int i1 = i0 + nthreads;
      double v0 = 0.0, v1 = 0.0; int2 w0 = make_int2(0, 0), w1 = make_int2(0, 0); int r0 = 0, r1 = 0;
      if (i0 < 36) v0 = po_ld_strong(d.Hb + (size_t)cn * 36 + i0);
      else if (i0 < nblk) v0 = po_ld_strong(d.Hb + (size_t)(Kf + oa) * 36 + (i0 - 36));
      else if (i0 < n1) v0 = po_ld_strong(d.bz + 6 * cn + (i0 - nblk));
      else if (i0 < n2) w0 = d.tri[ta + (i0 - n1)];
      else r0 = d.row_pos[oa + (i0 - n2)];
      if (i1 < n3) {
        if (i1 < 36) v1 = po_ld_strong(d.Hb + (size_t)cn * 36 + i1);
        else if (i1 < nblk) v1 = po_ld_strong(d.Hb + (size_t)(Kf + oa) * 36 + (i1 - 36));
        else if (i1 < n1) v1 = po_ld_strong(d.bz + 6 * cn + (i1 - nblk));
        else if (i1 < n2) w1 = d.tri[ta + (i1 - n1)];
        else r1 = d.row_pos[oa + (i1 - n2)];
      }
      if (i0 < nblk) dst[i0] = v0; else if (i0 < n1) bdst[i0 - nblk] = v0; else if (i0 < n2) ts[i0 - n1] = w0; else rs[i0 - n2] = r0;
      if (i1 < n3) { if (i1 < nblk) dst[i1] = v1; else if (i1 < n1) bdst[i1 - nblk] = v1; else if (i1 < n2) ts[i1 - n1] = w1; else rs[i1 - n2] = r1; }
    }
  };
  fetch_column(0, cache0, bzc, 0, PO_SP_NT);
  __syncthreads();
  if (tid < 32) {
    double W[21];
    const bool ok = po_spd6_inverse(cache0, W);
    if (lane == 0) {
      if (!ok) bad = 1;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) Wsm[6 * p + q] = W[p >= q ? p * (p + 1) / 2 + q : q * (q + 1) / 2 + p];
    }
  }
  __syncthreads();
  for (int c = 0; c < Kf; ++c) {
    // (offsets from the one shared-memory base, not selected pointers: the accesses then stay LDS / STS instead of
    // generic loads, and nothing is spilled to an indexed local array)
    double* cur = spsm + (c & 1) * PO_SP_CACHE;
    double* nxt = spsm + ((c + 1) & 1) * PO_SP_CACHE;
    const double* Wc = Wsm + 36 * (c & 1);
    double* Wn = Wsm + 36 * ((c + 1) & 1);
    double* bcur = bzc + 8 * (c & 1);
    double* bnxt = bzc + 8 * ((c + 1) & 1);
    const int o0 = d.col_off[c], o1 = d.col_off[c + 1], m = o1 - o0;
    const int o2 = (c + 1 < Kf) ? d.col_off[c + 2] : o1;
    const int nwork = 6 * m + 6, work_end = (nwork + 31) & ~31;
    // ---- panel phase ----
    if (tid < nwork) {
      if (tid < 6 * m) {
        const int a = tid / 6, p = tid - 6 * a;
        const double* arow = cur + 36 * (1 + a) + 6 * p;            // the original row stays in the cache for the updates
        double* grow = d.Hb + (size_t)(Kf + o0 + a) * 36 + 6 * p;   // P is kept in global memory for the back-substitution
        double av[6], pv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) av[k] = arow[k];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double s0 = av[0] * Wc[q] + av[1] * Wc[6 + q] + av[2] * Wc[12 + q];
          const double s1 = av[3] * Wc[18 + q] + av[4] * Wc[24 + q] + av[5] * Wc[30 + q];
          pv[q] = s0 + s1;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) { Pm[36 * a + 6 * p + k] = pv[k]; grow[k] = pv[k]; }
      } else {
        const int q = tid - 6 * m;
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double bk = bcur[k]; sum += bk * Wc[6 * k + q]; if (q == 0) bc[k] = bk; }
        d.us[6 * c + q] = sum;
      }
    } else if (tid >= work_end && c + 1 < Kf) {
      fetch_column(c + 1, nxt, bnxt, work_end, PO_SP_NT - work_end);
    }
    __syncthreads();
    if (work_end >= PO_SP_NT && c + 1 < Kf) {
      // a column so full that no warp was idle: fetch the next one now (rare: the last, dense columns of a full factor)
      fetch_column(c + 1, nxt, bnxt, 0, PO_SP_NT);
      __syncthreads();
    }
    if (tid == 0) { const long long now = clock64(); t_p1 += now - t_mark; t_mark = now; }
    // ---- update phase ----
    const int t0 = d.tri_off[c], nt = d.tri_off[c + 1] - t0;
    const int2* tl = (nt <= PO_SP_TRICAP) ? (tri_s + (c & 1) * PO_SP_TRICAP) : (d.tri + t0);
    const int* rps = rp_s + (c & 1) * PO_SP_MAXROWS;
    // the update of the next pivot block, when there is one, is the first entry (row c + 1 sorts first in the column)
    const bool pivot_first = nt > 0 && tl[0].x == c + 1;
    if (tid < 32) {
      if (c + 1 < Kf) {
        if (pivot_first) {
          const int2 tr = tl[0];
          for (int pq = lane; pq < 36; pq += 32) {
            const int p = pq / 6, q = pq - 6 * p;
            const double* pa = Pm + 36 * (tr.y & 0xffff) + 6 * p;
            const double* ob = cur + 36 * (1 + (tr.y >> 16)) + 6 * q;
            const double s0 = pa[0] * ob[0] + pa[1] * ob[1] + pa[2] * ob[2];
            const double s1 = pa[3] * ob[3] + pa[4] * ob[4] + pa[5] * ob[5];
            nxt[pq] -= s0 + s1;
          }
          __syncwarp();
        }
        double W[21];
        const bool ok = po_spd6_inverse(nxt, W);
        if (lane == 0) {
          if (!ok) bad = 1;
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q) Wn[6 * p + q] = W[p >= q ? p * (p + 1) / 2 + q : q * (q + 1) / 2 + p];
        }
      }
    } else {
      const int first = pivot_first ? 36 : 0;
      const int pan_lo = Kf + o1, pan_hi = Kf + o2;     // block ids of column c + 1's panel
      for (int e = first + tid - 32; e < 36 * nt + 6 * m; e += PO_SP_NT - 32) {
        if (e < 36 * nt) {
          const int t = e / 36, pq = e - 36 * t, p = pq / 6, q = pq - 6 * p;
          const int2 tr = tl[t];
          const double* pa = Pm + 36 * (tr.y & 0xffff) + 6 * p;
          const double* ob = cur + 36 * (1 + (tr.y >> 16)) + 6 * q;
          const double s0 = pa[0] * ob[0] + pa[1] * ob[1] + pa[2] * ob[2];
          const double s1 = pa[3] * ob[3] + pa[4] * ob[4] + pa[5] * ob[5];
          const double sv = s0 + s1;
          if (tr.x >= pan_lo && tr.x < pan_hi) nxt[36 * (1 + tr.x - pan_lo) + pq] -= sv;
          else if (tr.x == c + 1) nxt[pq] -= sv;                       // (only when the list was not sorted as expected)
          else atomicAdd(d.Hb + (size_t)tr.x * 36 + pq, -sv);          // result unused: a reduction, nobody waits for it
        } else {
          const int r = e - 36 * nt, a = r / 6, p = r - 6 * a;
          const double* pa = Pm + 36 * a + 6 * p;
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sum += pa[k] * bc[k];
          const int rp = rps[a];
          if (rp == c + 1) bnxt[p] -= sum; else atomicAdd(d.bz + 6 * rp + p, -sum);
        }
      }
    }
    __syncthreads();
    if (tid == 0) { const long long now = clock64(); t_p2 += now - t_mark; t_mark = now; }
  }
  if (tid == 0 && bad) d.st->chol_fail = 1;
  // ---- back-substitution, descending: y_c = u_c - sum_a P_ac^T y_row(a).  The P blocks of a chunk of columns (a
  // contiguous range of Hb), their u, row positions and offsets are staged in shared memory by all threads, then warp 0
  // walks the chunk's columns while the other warps stage the next chunk.
  const bool y_in_smem = d.n <= PO_SP_YMAX;
  double* yv = y_in_smem ? ysm : d.yp;
  double* us_s = reinterpret_cast<double*>(tri_s);                // [2][6 * MAXROWS] (the update lists are dead by now)
  const int nchunk = d.bs_nchunk;
  auto stage_chunk = [&](int k, int first_thread, int nthreads) {
    if (k >= nchunk) return;
    const int c_hi = d.bs_chunk[k], c_lo = d.bs_chunk[k + 1];           // columns [c_lo, c_hi)
    const int b0 = d.col_off[c_lo], nb = d.col_off[c_hi] - b0, ncol = c_hi - c_lo;
    double* dst = spsm + (k & 1) * PO_SP_CACHE;
    for (int i = tid - first_thread; i < 36 * nb; i += nthreads) dst[i] = __ldcg(d.Hb + (size_t)(Kf + b0) * 36 + i);
    double* ud = us_s + (k & 1) * 6 * PO_SP_MAXROWS;
    for (int i = tid - first_thread; i < 6 * ncol; i += nthreads) ud[i] = __ldcg(d.us + 6 * c_lo + i);
    int* id = bs_i + (k & 1) * 400;                                     // [0, ncol]: column offsets, [200, 200 + nb): row positions
    for (int i = tid - first_thread; i <= ncol; i += nthreads) id[i] = d.col_off[c_lo + i] - b0;
    for (int i = tid - first_thread; i < nb; i += nthreads) id[200 + i] = d.row_pos[b0 + i];
  };
  stage_chunk(0, 0, PO_SP_NT);
  __syncthreads();
  for (int k = 0; k < nchunk; ++k) {
    if (tid >= 32) {
      stage_chunk(k + 1, 32, PO_SP_NT - 32);
    } else {
      const int c_hi = d.bs_chunk[k], c_lo = d.bs_chunk[k + 1];
      const double* Ps = spsm + (k & 1) * PO_SP_CACHE;
      const double* ud = us_s + (k & 1) * 6 * PO_SP_MAXROWS;
      const int* id = bs_i + (k & 1) * 400;
      // lane (g, q), g < 5, q < 6, sums the (row block, row) pairs j = g, g + 5, ... of entry q of P^T y; four shuffles
      // fold the five partial sums in a fixed order (the instruction count of this single-warp chain is what it costs:
      // six 5-level butterflies per column were 3x slower)
      const int g = lane / 6, q = lane - 6 * g;
      for (int c = c_hi - 1; c >= c_lo; --c) {
        const int o0 = id[c - c_lo], m = id[c - c_lo + 1] - o0;
        double acc = 0.0;
        if (lane < 30) {
          for (int j = g; j < 6 * m; j += 5) {
            const int a = j / 6, p = j - 6 * a;
            const int yi = 6 * id[200 + o0 + a] + p;
            const double yr = y_in_smem ? ysm[yi] : d.yp[yi];
            acc += Ps[36 * (o0 + a) + 6 * p + q] * yr;
          }
        }
        const double a1 = __shfl_down_sync(0xffffffffu, acc, 6), a2 = __shfl_down_sync(0xffffffffu, acc, 12);
        const double a3 = __shfl_down_sync(0xffffffffu, acc, 18), a4 = __shfl_down_sync(0xffffffffu, acc, 24);
        if (lane < 6) {
          const double yc_ = ud[6 * (c - c_lo) + lane] - ((((acc + a1) + a2) + a3) + a4);
          if (y_in_smem) ysm[6 * c + lane] = yc_; else d.yp[6 * c + lane] = yc_;
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (tid == 0 && d.sp_cycles) {
    const long long now = clock64();
    d.sp_cycles[0] = t_p1; d.sp_cycles[1] = t_p2; d.sp_cycles[2] = now - t_mark; d.sp_cycles[3] = now - t_start;
  }
  // solution back in slot order
  for (int i = tid; i < d.n; i += PO_SP_NT) d.y[i] = yv[6 * d.slot_pos[i / 6] + i % 6];
}

// ------------------------------------------------------------------------------------------------------------
// Level-ordered factorisation (round 2, second form).  The minimum-degree order peels a trajectory graph from its ends:
// the elimination tree is a path and po_sp_factor_solve walks ~250 dependent block columns at ~3.3 k cycles each.  The
// host now orders by LEVELS of independent, near-minimum-degree poses (every other pose of the chain, then every other
// of what is left, ...; po_host.cu) and splits a level into stages whose columns touch disjoint blocks.  Here ONE WARP
// eliminates a column -- pivot inverse, scaled panel (written over the column: the back-substitution reads it), update of
// the blocks between its rows, right-hand side -- and the 16 warps of the CTA take the columns of a stage side by side:
// plain read-modify-writes, fixed order, no atomics; a __syncthreads separates the stages.  Back-substitution by stages in
// reverse, warp per column.  Everything the stages exchange goes through L2 (loads with .cg).
// ------------------------------------------------------------------------------------------------------------
constexpr int PO_LV_NT = 256;         // 8 warps: the column code wants more than the 128 registers 512 threads would leave
constexpr int PO_LV_CLUSTER = 8;      // CTAs of the cluster the kernel is launched as (portable maximum)
constexpr int PO_LV_COOP = 4;         // stages with fewer columns than this: the whole CTA works on one column at a time
constexpr int PO_LV_MAXROWS = 12;       // off-diagonal blocks per column the per-warp staging holds
constexpr int PO_LV_MAXTRI = PO_LV_MAXROWS * (PO_LV_MAXROWS + 1) / 2;
constexpr int PO_LV_WARP_DOUBLES = 2 * PO_LV_MAXROWS * 36 + PO_LV_MAXTRI;            // per-warp staging: blocks, panel, update list
constexpr int PO_LV_STAGE_DOUBLES = (PO_LV_NT / 32) * PO_LV_WARP_DOUBLES;
constexpr int PO_LV_TAILCOLS = 96, PO_LV_TAILBLK = 400;                               // dense end kept in shared memory by the back-substitution
constexpr size_t PO_LV_SMEM = (size_t)(PO_LV_STAGE_DOUBLES + PO_LV_TAILBLK * 36 + 6 * PO_LV_TAILCOLS) * 8 + (size_t)(PO_LV_TAILBLK + PO_LV_TAILCOLS + 2) * 4;

__global__ void __launch_bounds__(PO_LV_NT, 1) po_sp_factor_levels(PoDev d) {
  if (d.st->done) return;
  extern __shared__ __align__(16) double lvsm[];
  __shared__ int bad;
  __shared__ double cW[36], cbc[8];                   // cooperative mode: pivot inverse and right-hand-side block
  // launched as ONE thread-block cluster of gridDim.x CTAs (<= 8, on the SMs of one GPC): the columns of a stage are
  // spread over the warps of all of them, a hardware cluster barrier (release / acquire) separates the stages, and
  // everything the CTAs exchange goes through L2 (every load of factor data is .cg)
  const int ncta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) + (PO_LV_NT / 32) * cta, nw = (PO_LV_NT / 32) * ncta, Kf = d.Kf;
  auto stage_sync = [&]() {
    if (ncta == 1) { __syncthreads(); return; }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  };
  double* stA = lvsm + (size_t)(tid >> 5) * PO_LV_WARP_DOUBLES;     // unscaled blocks of the warp's column
  double* stP = stA + PO_LV_MAXROWS * 36;                         // scaled panel P = A W
  int2* tri_s = reinterpret_cast<int2*>(stP + PO_LV_MAXROWS * 36);   // the column's update list
  if (tid == 0) bad = 0;
  long long t_start = 0, t_mid = 0, t_coop = 0;
  if (tid == 0) t_start = clock64();
  __syncthreads();
  for (int s = 0; s < d.nstage; ++s) {
    const int c0 = d.stage_off[s], c1 = d.stage_off[s + 1];
    long long t_st = 0;
    if (tid == 0) t_st = clock64();
    if (c1 - c0 < PO_LV_COOP) {
      // few columns (the dense end of the elimination: long columns, one per stage): CTA 0 takes them one at a time --
      // warp 0 inverts the pivot, 6 m threads scale the panel, the update rows are spread over all threads -- and it
      // takes the whole run of consecutive short stages in one go (one cluster barrier for the run)
      double* cA = lvsm;                                   // warp 0's staging, shared by everybody here
      double* cP = cA + PO_LV_MAXROWS * 36;
      int2* ctri = reinterpret_cast<int2*>(cP + PO_LV_MAXROWS * 36);
      int s_end = s + 1;
      while (s_end < d.nstage && d.stage_off[s_end + 1] - d.stage_off[s_end] < PO_LV_COOP) ++s_end;
      const int c_run = d.stage_off[s_end];
      s = s_end - 1;
      for (int c = c0; c < c_run && cta == 0; ++c) {
        const int o0 = d.col_off[c], m = d.col_off[c + 1] - o0;
        const int t0 = d.tri_off[c], nt = d.tri_off[c + 1] - t0;
        for (int i = tid; i < nt; i += PO_LV_NT) ctri[i] = d.tri[t0 + i];
        if (tid < 6) cbc[tid] = __ldcg(d.bz + 6 * c + tid);
        if (warp == 0) {
          double A36[36], W[21];
#pragma unroll
          for (int pp = 0; pp < 6; ++pp)
#pragma unroll
            for (int q = 0; q <= pp; ++q) A36[6 * pp + q] = __ldcg(d.Hb + (size_t)c * 36 + 6 * pp + q);
          const bool ok = po_spd6_inverse(A36, W);
          if (lane == 0) {
            if (!ok) bad = 1;
#pragma unroll
            for (int pp = 0; pp < 6; ++pp)
#pragma unroll
              for (int q = 0; q < 6; ++q) cW[6 * pp + q] = W[pp >= q ? pp * (pp + 1) / 2 + q : q * (q + 1) / 2 + pp];
          }
        }
        __syncthreads();
        if (tid < 6 * m) {
          const int a = tid / 6, pr = tid - 6 * a;
          double* blk = d.Hb + (size_t)(Kf + o0 + a) * 36 + 6 * pr;
          double av[6], pv[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) av[k] = __ldcg(blk + k);
          const int rrow = 6 * d.row_pos[o0 + a] + pr;
          const double bold = __ldcg(d.bz + rrow);
#pragma unroll
          for (int q = 0; q < 6; ++q)
            pv[q] = (av[0] * cW[q] + av[1] * cW[6 + q] + av[2] * cW[12 + q]) + (av[3] * cW[18 + q] + av[4] * cW[24 + q] + av[5] * cW[30 + q]);
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) { cA[36 * a + 6 * pr + k] = av[k]; cP[36 * a + 6 * pr + k] = pv[k]; sacc += pv[k] * cbc[k]; }
#pragma unroll
          for (int k = 0; k < 6; ++k) blk[k] = pv[k];
          d.bz[rrow] = bold - sacc;
        } else if (tid >= 96 && tid < 102) {
          const int q = tid - 96;
          double u = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) u += cW[6 * q + k] * cbc[k];
          d.us[6 * c + q] = u;
        }
        __syncthreads();
        for (int base = 0; base < 6 * nt; base += 2 * PO_LV_NT) {
          double o[2][6];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = base + PO_LV_NT * u + tid;
            if (i < 6 * nt) {
              const int t = i / 6, pr = i - 6 * t;
              const double* dst = d.Hb + (size_t)ctri[t].x * 36 + 6 * pr;
#pragma unroll
              for (int k = 0; k < 6; ++k) o[u][k] = __ldcg(dst + k);
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = base + PO_LV_NT * u + tid;
            if (i < 6 * nt) {
              const int t = i / 6, pr = i - 6 * t;
              const int2 tr = ctri[t];
              const int a = tr.y & 0xffff, b = tr.y >> 16;
              const double* Pa = cP + 36 * a + 6 * pr;
              const double* Ab = cA + 36 * b;
              double* dst = d.Hb + (size_t)tr.x * 36 + 6 * pr;
              double pa[6];
#pragma unroll
              for (int k = 0; k < 6; ++k) pa[k] = Pa[k];
#pragma unroll
              for (int q = 0; q < 6; ++q)
                o[u][q] -= (pa[0] * Ab[6 * q] + pa[1] * Ab[6 * q + 1] + pa[2] * Ab[6 * q + 2]) + (pa[3] * Ab[6 * q + 3] + pa[4] * Ab[6 * q + 4] + pa[5] * Ab[6 * q + 5]);
#pragma unroll
              for (int q = 0; q < 6; ++q) dst[q] = o[u][q];
            }
          }
        }
        __syncthreads();
      }
      stage_sync();
      if (tid == 0) t_coop += clock64() - t_st;
      continue;
    }
    for (int c = c0 + warp; c < c1; c += nw) {
      const int o0 = d.col_off[c], m = d.col_off[c + 1] - o0;
      const int t0 = d.tri_off[c], nt = d.tri_off[c + 1] - t0;
      for (int i = lane; i < nt; i += 32) tri_s[i] = d.tri[t0 + i];          // in flight with the pivot loads
      // pivot inverse (every lane the same values) and right-hand-side block
      double A36[36], W[21], bc[6];
#pragma unroll
      for (int pp = 0; pp < 6; ++pp)
#pragma unroll
        for (int q = 0; q <= pp; ++q) A36[6 * pp + q] = __ldcg(d.Hb + (size_t)c * 36 + 6 * pp + q);
#pragma unroll
      for (int k = 0; k < 6; ++k) bc[k] = __ldcg(d.bz + 6 * c + k);
      const bool ok = po_spd6_inverse(A36, W);
      if (!ok && lane == 0) bad = 1;
#define WF(i, j) W[(i) >= (j) ? (i) * ((i) + 1) / 2 + (j) : (j) * ((j) + 1) / 2 + (i)]
      if (lane < 6) {
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          // (W symmetric) row `lane` of W times b_c, with compile-time indices into W
          const double w = lane == 0 ? WF(0, k) : lane == 1 ? WF(1, k) : lane == 2 ? WF(2, k) : lane == 3 ? WF(3, k) : lane == 4 ? WF(4, k) : WF(5, k);
          u += w * bc[k];
        }
        d.us[6 * c + lane] = u;
      }
      // panel: row p of P_a = A_a W; the unscaled row kept for the updates; right-hand side of the row block.  At most
      // three rows per lane (m <= 12): every load first, then the arithmetic, then the stores (one L2 round trip)
      {
        double av[3][6], bold[3];
        int rrow[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i = lane + 32 * u;
          rrow[u] = 0; bold[u] = 0.0;
          if (i < 6 * m) {
            const int a = i / 6, pr = i - 6 * a;
            const double* blk = d.Hb + (size_t)(Kf + o0 + a) * 36 + 6 * pr;
#pragma unroll
            for (int k = 0; k < 6; ++k) av[u][k] = __ldcg(blk + k);
            rrow[u] = 6 * d.row_pos[o0 + a] + pr;
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) if (lane + 32 * u < 6 * m) bold[u] = __ldcg(d.bz + rrow[u]);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i = lane + 32 * u;
          if (i < 6 * m) {
            const int a = i / 6, pr = i - 6 * a;
            double* blk = d.Hb + (size_t)(Kf + o0 + a) * 36 + 6 * pr;
            double pv[6];
#pragma unroll
            for (int q = 0; q < 6; ++q)
              pv[q] = (av[u][0] * WF(0, q) + av[u][1] * WF(1, q) + av[u][2] * WF(2, q)) + (av[u][3] * WF(3, q) + av[u][4] * WF(4, q) + av[u][5] * WF(5, q));
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { stA[36 * a + 6 * pr + k] = av[u][k]; stP[36 * a + 6 * pr + k] = pv[k]; sacc += pv[k] * bc[k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) blk[k] = pv[k];
            d.bz[rrow[u]] = bold[u] - sacc;
          }
        }
      }
#undef WF
      __syncwarp();
      // updates: destination block (rows of a, columns of b) -= P_a A_b^T, one row per lane-item; two items per lane at a
      // time with all their destination loads issued first (no two items of a column share a destination row)
      for (int base = 0; base < 6 * nt; base += 64) {
        double o[2][6];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = base + 32 * u + lane;
          if (i < 6 * nt) {
            const int t = i / 6, pr = i - 6 * t;
            const double* dst = d.Hb + (size_t)tri_s[t].x * 36 + 6 * pr;
#pragma unroll
            for (int k = 0; k < 6; ++k) o[u][k] = __ldcg(dst + k);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = base + 32 * u + lane;
          if (i < 6 * nt) {
            const int t = i / 6, pr = i - 6 * t;
            const int2 tr = tri_s[t];
            const int a = tr.y & 0xffff, b = tr.y >> 16;
            const double* Pa = stP + 36 * a + 6 * pr;
            const double* Ab = stA + 36 * b;
            double* dst = d.Hb + (size_t)tr.x * 36 + 6 * pr;
            double pa[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) pa[k] = Pa[k];
#pragma unroll
            for (int q = 0; q < 6; ++q)
              o[u][q] -= (pa[0] * Ab[6 * q] + pa[1] * Ab[6 * q + 1] + pa[2] * Ab[6 * q + 2]) + (pa[3] * Ab[6 * q + 3] + pa[4] * Ab[6 * q + 4] + pa[5] * Ab[6 * q + 5]);
#pragma unroll
            for (int q = 0; q < 6; ++q) dst[q] = o[u][q];
          }
        }
      }
      __syncwarp();
    }
    stage_sync();
  }
  if (tid == 0) { t_mid = clock64(); if (bad) d.st->chol_fail = 1; }
  // back-substitution.  First the dense end (the final run of short stages, columns [tail0, Kf)): CTA 0 stages all their
  // panel blocks, row positions and u in shared memory with one round of loads, then one warp walks the columns downwards
  // with y in shared memory -- no L2 round trip on the chain.
  int s_hi = d.nstage - 1;
  {
    int s_lo = d.nstage;
    while (s_lo > 0 && d.stage_off[s_lo] - d.stage_off[s_lo - 1] < PO_LV_COOP) --s_lo;
    const int tail0 = d.stage_off[s_lo], ntail = Kf - tail0;
    const int b0 = ntail > 0 ? d.col_off[tail0] : 0, nbt = ntail > 0 ? d.col_off[Kf] - b0 : 0;
    if (ntail > 0 && ntail <= PO_LV_TAILCOLS && nbt <= PO_LV_TAILBLK) {
      if (cta == 0) {
        double* Pt = lvsm + PO_LV_STAGE_DOUBLES;             // [nbt][36]
        double* yt = Pt + PO_LV_TAILBLK * 36;                // [6 ntail] u, then y
        int* rpt = reinterpret_cast<int*>(yt + 6 * PO_LV_TAILCOLS);   // [nbt] row positions, then [ntail + 1] column offsets
        int* cot = rpt + PO_LV_TAILBLK;
        for (int i = tid; i < 36 * nbt; i += PO_LV_NT) Pt[i] = __ldcg(d.Hb + (size_t)(Kf + b0) * 36 + i);
        for (int i = tid; i < 6 * ntail; i += PO_LV_NT) yt[i] = __ldcg(d.us + 6 * tail0 + i);
        for (int i = tid; i < nbt; i += PO_LV_NT) rpt[i] = d.row_pos[b0 + i];
        for (int i = tid; i <= ntail; i += PO_LV_NT) cot[i] = d.col_off[tail0 + i] - b0;
        __syncthreads();
        if ((tid >> 5) == 0) {
          for (int c = Kf - 1; c >= tail0; --c) {
            const int o0 = cot[c - tail0], m = cot[c - tail0 + 1] - o0;
            // lane (g, q), g < 5, q < 6, sums the (row block, row) pairs j = g, g + 5, ... of entry q of P^T y; four
            // shuffles fold the five partial sums in a fixed order (six 5-level butterflies cost three times as much)
            const int g = lane / 6, q = lane - 6 * g;
            double acc = 0.0;
            if (lane < 30) {
              for (int j = g; j < 6 * m; j += 5) {
                const int a = j / 6, pr = j - 6 * a;
                acc += Pt[36 * (o0 + a) + 6 * pr + q] * yt[6 * (rpt[o0 + a] - tail0) + pr];   // rows of a tail column are later tail columns
              }
            }
            const double a1 = __shfl_down_sync(0xffffffffu, acc, 6), a2 = __shfl_down_sync(0xffffffffu, acc, 12);
            const double a3 = __shfl_down_sync(0xffffffffu, acc, 18), a4 = __shfl_down_sync(0xffffffffu, acc, 24);
            if (lane < 6) yt[6 * (c - tail0) + lane] -= (((acc + a1) + a2) + a3) + a4;
            __syncwarp();
          }
        }
        __syncthreads();
        for (int i = tid; i < 6 * ntail; i += PO_LV_NT) d.yp[6 * tail0 + i] = yt[i];
      }
      stage_sync();
      s_hi = s_lo - 1;
    }
  }
  // the other stages in reverse: y_c = u_c - sum_a P_ac^T y_row(a); warp per column, lanes over (row block, row), butterfly
  for (int s = s_hi; s >= 0; --s) {
    const int c0 = d.stage_off[s], c1 = d.stage_off[s + 1];
    // a short stage: CTA 0 alone, and no cluster barrier until the run of short stages ends
    const bool small = c1 - c0 < PO_LV_COOP;
    const bool next_small = small && s > 0 && d.stage_off[s] - d.stage_off[s - 1] < PO_LV_COOP;
    const int w0 = small ? (tid >> 5) : warp, wn = small ? PO_LV_NT / 32 : nw;
    if (small && cta != 0) { if (!next_small) stage_sync(); continue; }
    for (int c = c0 + w0; c < c1; c += wn) {
      const int o0 = d.col_off[c], m = d.col_off[c + 1] - o0;
      const int g = lane / 6, q = lane - 6 * g;          // (see the tail above)
      double acc = 0.0;
      if (lane < 30) {
        for (int j = g; j < 6 * m; j += 5) {
          const int a = j / 6, pr = j - 6 * a;
          acc += __ldcg(d.Hb + (size_t)(Kf + o0 + a) * 36 + 6 * pr + q) * __ldcg(d.yp + 6 * d.row_pos[o0 + a] + pr);
        }
      }
      const double a1 = __shfl_down_sync(0xffffffffu, acc, 6), a2 = __shfl_down_sync(0xffffffffu, acc, 12);
      const double a3 = __shfl_down_sync(0xffffffffu, acc, 18), a4 = __shfl_down_sync(0xffffffffu, acc, 24);
      if (lane < 6) d.yp[6 * c + lane] = __ldcg(d.us + 6 * c + lane) - ((((acc + a1) + a2) + a3) + a4);
    }
    if (small && next_small) __syncthreads(); else stage_sync();
  }
  if (tid == 0 && cta == 0 && d.sp_cycles) {
    const long long now = clock64();
    d.sp_cycles[0] = t_mid - t_start - t_coop; d.sp_cycles[1] = t_coop; d.sp_cycles[2] = now - t_mid; d.sp_cycles[3] = now - t_start;
  }
  // solution back in slot order
  for (int i = tid + PO_LV_NT * cta; i < d.n; i += PO_LV_NT * ncta) d.y[i] = __ldcg(d.yp + 6 * d.slot_pos[i / 6] + i % 6);
}

// trial point x' = x - scale*y on the free poses, and the per-edge part of the model decrease -(m.(r + m/2)), m = J delta
__global__ void po_step(PoDev d) {
  if (d.st->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 6 * d.K) {
    const int s = d.slot[t / 6];
    double v = d.x[t];
    if (s >= 0) v -= d.y[6 * s + t % 6] * d.scale[6 * s + t % 6];
    d.xt[t] = v;
  } else {
    const int e = t - 6 * d.K;
    if (e >= d.E) return;
    double acc = 0.0;
    if (d.active[e]) {
      const int set = d.st->cur;
      const int s1 = d.slot[d.idx1[e]], s2 = d.slot[d.idx2[e]];
      const double* J1 = d.J1 + ((size_t)set * d.E + e) * 36;
      const double* J2 = d.J2 + ((size_t)set * d.E + e) * 36;
      const double* r = d.r + ((size_t)set * d.E + e) * 6;
      double d1[6], d2[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        d1[j] = s1 >= 0 ? -d.y[6 * s1 + j] * d.scale[6 * s1 + j] : 0.0;
        d2[j] = s2 >= 0 ? -d.y[6 * s2 + j] * d.scale[6 * s2 + j] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) m += J1[6 * k + j] * d1[j] + J2[6 * k + j] * d2[j];
        acc += m * (r[k] + 0.5 * m);
      }
    }
    d.mval[e] = acc;
  }
}

// Step acceptance and trust-region update (TrustRegionMinimizer / LevenbergMarquardtStrategy semantics).
__global__ void po_decide(PoDev d) {
  __shared__ double sh[256];
  PoState* st = d.st;
  if (st->done) return;
  const int tid = threadIdx.x, tset = 1 - st->cur;
  double nc = 0.0, mv = 0.0, dn = 0.0, bad = 0.0;
  for (int e = tid; e < d.E; e += 256) if (d.active[e]) { nc += d.cost_e[(size_t)tset * d.E + e]; mv += d.mval[e]; }
  for (int i = tid; i < d.n; i += 256) {
    const double y = d.y[i], dl = y * d.scale[i];
    dn += dl * dl;
    if (!isfinite(y)) bad = 1.0;
  }
  nc = po_block_reduce<false>(nc, sh);
  mv = po_block_reduce<false>(mv, sh);
  dn = po_block_reduce<false>(dn, sh);
  bad = po_block_reduce<true>(bad, sh);
  if (tid != 0) return;
  const int it = st->it;
  st->it = it + 1;
  st->iters = it + 1;
  double* tr = d.trace ? d.trace + (size_t)it * SLSLAM_TRACE_WIDTH : nullptr;
  if (tr) { tr[0] = st->cost; tr[1] = 0; tr[2] = 0; tr[3] = st->radius; tr[4] = 0; tr[5] = 0; tr[6] = st->gmax; tr[7] = 0; }
  const bool ok = !st->chol_fail && bad == 0.0;
  st->chol_fail = 0;
  st->accepted = 0;
  const double model = ok ? -mv : 0.0;
  if (tr) tr[2] = model;
  if (!ok || model < 0.0) {
    ++st->unsuccessful;
    if (tr) tr[5] = -1.0;
    if (++st->invalid >= 5) { st->term = SLSLAM_NUMERICAL_FAILURE; st->done = 1; return; }
    st->radius *= 0.5;
    if (st->radius < 1e-32) { st->term = SLSLAM_PARAMETER_TOLERANCE; st->done = 1; }
    if (it + 1 >= st->max_iters) st->done = 1;
    return;
  }
  st->invalid = 0;
  const double step_norm = sqrt(dn);
  if (tr) { tr[1] = nc; tr[4] = step_norm; }
  if (step_norm <= st->ptol * (st->x_norm + st->ptol)) { st->term = SLSLAM_PARAMETER_TOLERANCE; st->done = 1; return; }
  const double change = st->cost - nc;
  if (fabs(change) < st->ftol * st->cost) { st->term = SLSLAM_FUNCTION_TOLERANCE; st->done = 1; return; }
  const double rel = change / model;
  if (tr) tr[7] = rel;
  if (rel > 1e-3) {
    ++st->successful;
    if (tr) tr[5] = 1.0;
    st->cost = nc;
    st->cur = tset;            // the trial linearisation becomes the current one
    st->accepted = 1;
    const double q = 2.0 * rel - 1.0;
    st->radius = fmin(1e16, st->radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
    st->decrease_factor = 2.0;
  } else {
    ++st->unsuccessful;
    st->radius /= st->decrease_factor;
    st->decrease_factor *= 2.0;
  }
  if (st->radius < 1e-32) { st->term = SLSLAM_PARAMETER_TOLERANCE; st->done = 1; return; }
  if (it + 1 >= st->max_iters && !st->accepted) st->done = 1;
}

__global__ void po_accept(PoDev d) {
  const PoState* st = d.st;
  if (st->done || !st->accepted) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 6 * d.K) d.x[t] = d.xt[t];
}

__global__ void po_finish(PoDev d) {
  const PoState* st = d.st;
  slslam_summary s;
  s.initial_cost = st->initial_cost; s.final_cost = st->cost + st->fixed_cost; s.fixed_cost = st->fixed_cost;
  s.gradient_max_norm = st->gmax; s.num_successful_steps = st->successful; s.num_unsuccessful_steps = st->unsuccessful;
  s.termination_type = st->term; s.iterations = st->iters;
  *d.summary = s;
}

}  // namespace slslam
