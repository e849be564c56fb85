// Device math for the LBA hot path: the Pluecker-line stereo reprojection residual of
// LineReprojectionError::operator() (reference src/lba_problem.h:46-118) and its ANALYTIC Jacobian with respect to
// the 6-dof camera (angle-axis, translation) and the 4-dof orthonormal line (a, b, g, t).  The reference gets the
// Jacobian from Ceres dual numbers (src/lba_problem.cpp:65-74); here it is derived by hand
// so that all trigonometry is per camera / per line and the per-observation work is ~350 FMAs.
#pragma once
#include <cuda_runtime.h>

namespace slslam {

// 1 / sqrt(d) for the pivot chains of the block Choleskys: the hardware approximation (MUFU.RSQ64H, ~20 bits) and two
// Newton steps y <- y (1.5 - 0.5 d y^2), accurate to a few ulp.  The library rsqrt() spends most of its ~200 cycles of
// latency on special cases that cannot occur here (the callers test d > 0 separately; a non-positive or non-finite
// pivot still yields NaN / inf, which the step-validity checks catch).
__device__ __forceinline__ double pivot_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// 1 / d from the hardware approximation (MUFU.RCP64H) and two Newton steps; same remarks as pivot_rsqrt.
__device__ __forceinline__ double pivot_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// Per-camera block staged in shared memory: R (9, row-major), dR/dw_k (3 x 9), t (3), pad.  The lanes of a tile read
// the same element of different cameras with 8-byte loads; an odd stride (41 doubles = 82 words, 18 c mod 32 banks)
// keeps up to 16 cameras on distinct bank pairs, where 40 put every other camera on the same pair.
constexpr int CAM_STRIDE = 41;

// R = exp([w]x) with the first-order branch R = I + [w]x when |w|^2 == 0, which is what
// ceres::AngleAxisRotatePoint evaluates (reference src/lba_problem.h:75-76; the newest keyframe is exactly
// identity, src/slam.cpp:1322).  dR[k] = dR/dw_k with respect to the GLOBAL angle-axis vector.
__device__ __forceinline__ void cam_precompute(const double* __restrict__ cam, double* __restrict__ out, bool with_jac) {
  const double w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
  double* R = out;
  double* dR = out + 9;
  out[36] = cam[3]; out[37] = cam[4]; out[38] = cam[5]; out[39] = 0.0;
  const double w[3] = {w0, w1, w2};
  if (th2 > 0.0) {
    const double th = sqrt(th2);
    double s, c, sh, ch;
    sincos(th, &s, &c);
    sincos(0.5 * th, &sh, &ch);
    const double A = s / th;
    const double hs = sh / (0.5 * th);
    const double B = 0.5 * hs * hs;                 // (1 - cos th)/th^2 without cancellation
    R[0] = c + B * w0 * w0;   R[1] = B * w0 * w1 - A * w2; R[2] = B * w0 * w2 + A * w1;
    R[3] = B * w1 * w0 + A * w2; R[4] = c + B * w1 * w1;   R[5] = B * w1 * w2 - A * w0;
    R[6] = B * w2 * w0 - A * w1; R[7] = B * w2 * w1 + A * w0; R[8] = c + B * w2 * w2;
    if (with_jac) {
      double E, F;                                   // E = (cos - A)/th^2, F = (A - 2B)/th^2
      if (th2 < 1e-3) {
        E = -1.0 / 3.0 + th2 * (1.0 / 30.0 - th2 * (1.0 / 840.0 - th2 * (1.0 / 45360.0)));
        F = -1.0 / 12.0 + th2 * (1.0 / 180.0 - th2 * (1.0 / 6720.0 - th2 * (1.0 / 453600.0)));
      } else {
        E = (c - A) / th2;
        F = (A - 2.0 * B) / th2;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double wk = w[k];
        double* D = dR + 9 * k;
        // -A wk I + E wk [w]x + A [e_k]x + F wk w w^T + B (e_k w^T + w e_k^T)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double v = F * wk * w[i] * w[j];
            if (i == j) v -= A * wk;
            if (i == k) v += B * w[j];
            if (j == k) v += B * w[i];
            D[3 * i + j] = v;
          }
        // E wk [w]x
        D[1] -= E * wk * w2; D[2] += E * wk * w1;
        D[3] += E * wk * w2; D[5] -= E * wk * w0;
        D[6] -= E * wk * w1; D[7] += E * wk * w0;
      }
      // A [e_k]x
      dR[5] -= A; dR[7] += A;                        // k = 0: [e0]x = [[0,0,0],[0,0,-1],[0,1,0]]
      dR[9 + 2] += A; dR[9 + 6] -= A;                // k = 1: [[0,0,1],[0,0,0],[-1,0,0]]
      dR[18 + 1] -= A; dR[18 + 3] += A;              // k = 2: [[0,-1,0],[1,0,0],[0,0,0]]
    }
  } else {
    R[0] = 1.0; R[1] = -w2; R[2] = w1;
    R[3] = w2;  R[4] = 1.0; R[5] = -w0;
    R[6] = -w1; R[7] = w0;  R[8] = 1.0;
    if (with_jac) {
#pragma unroll
      for (int i = 0; i < 27; ++i) dR[i] = 0.0;
      dR[5] = -1.0; dR[7] = 1.0;
      dR[9 + 2] = 1.0; dR[9 + 6] = -1.0;
      dR[18 + 1] = -1.0; dR[18 + 3] = 1.0;
    }
  }
}

// Per-line trigonometry (reference src/lba_problem.h:56-72): [xh yh zh] = Rz(g) Ry(b) Rx(a), d = cot t.
// closest point cp = -zh d, direction dv = yh, Pluecker normal cp x dv = d xh.
struct LineTrig {
  double xh[3], yh[3], zh[3];
  double xb[3];      // d xh / d b
  double d, ist2;    // cot t, 1/sin^2 t
  double s1;
};

// from the eight sines / cosines sc = {s1, c1, s2, c2, s3, c3, st, ct} of (a, b, g, t)
__device__ __forceinline__ void line_trig_sc(const double* sc, LineTrig& lt) {
  const double s1 = sc[0], c1 = sc[1], s2 = sc[2], c2 = sc[3], s3 = sc[4], c3 = sc[5], st = sc[6], ct = sc[7];
  lt.xh[0] = c2 * c3; lt.xh[1] = c2 * s3; lt.xh[2] = -s2;
  lt.yh[0] = s1 * s2 * c3 - c1 * s3; lt.yh[1] = s1 * s2 * s3 + c1 * c3; lt.yh[2] = s1 * c2;
  lt.zh[0] = c1 * s2 * c3 + s1 * s3; lt.zh[1] = c1 * s2 * s3 - s1 * c3; lt.zh[2] = c1 * c2;
  lt.xb[0] = -s2 * c3; lt.xb[1] = -s2 * s3; lt.xb[2] = -c2;
  const double ist = pivot_rcp(st);
  lt.d = ct * ist;
  lt.ist2 = ist * ist;
  lt.s1 = s1;
}

__device__ __forceinline__ void line_trig(const double* __restrict__ ln, LineTrig& lt) {
  double sc[8];
  sincos(ln[0], &sc[0], &sc[1]);
  sincos(ln[1], &sc[2], &sc[3]);
  sincos(ln[2], &sc[4], &sc[5]);
  sincos(ln[3], &sc[6], &sc[7]);
  line_trig_sc(sc, lt);
}

__device__ __forceinline__ void mv3(const double* __restrict__ M, const double* v, double* o) {
  o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}

// One stereo observation.  cpre: CAM_STRIDE doubles from cam_precompute.  ob: x0 y0 x1 y1 | x2 y2 x3 y3.
// r[4]; Jc[24] row-major 4x6 (w0 w1 w2 t0 t1 t2); Jl[16] row-major 4x4 (a b g t).
template <bool JAC>
__device__ __forceinline__ void obs_eval(const double* __restrict__ cpre, const LineTrig& lt, const double* ob,
                                         double bl, double* r, double* Jc, double* Jl) {
  const double* R = cpre;
  const double t0 = cpre[36], t1 = cpre[37], t2 = cpre[38];
  double m[3], q[3];
  mv3(R, lt.xh, m);
  mv3(R, lt.yh, q);
  // n = (R cp + t - o) x (R dv) = d R xh + (t - o) x R yh ; o_A = 0, o_B = (bl,0,0)
  const double nA0 = lt.d * m[0] + (t1 * q[2] - t2 * q[1]);
  const double nA1 = lt.d * m[1] + (t2 * q[0] - t0 * q[2]);
  const double nA2 = lt.d * m[2] + (t0 * q[1] - t1 * q[0]);
  const double nB0 = nA0, nB1 = nA1 + bl * q[2], nB2 = nA2 - bl * q[1];
  // 1 / sqrt by the hardware seed + two Newton steps (a few ulp; ~8 instructions instead of the ~50 of sqrt + divide)
  const double isA = pivot_rsqrt(nA0 * nA0 + nA1 * nA1);
  const double isB = pivot_rsqrt(nB0 * nB0 + nB1 * nB1);
  const double hA0 = nA0 * isA, hA1 = nA1 * isA, hA2 = nA2 * isA;
  const double hB0 = nB0 * isB, hB1 = nB1 * isB, hB2 = nB2 * isB;
  r[0] = -(ob[0] * hA0 + ob[1] * hA1 + hA2);
  r[1] = -(ob[2] * hA0 + ob[3] * hA1 + hA2);
  r[2] = -(ob[4] * hB0 + ob[5] * hB1 + hB2);
  r[3] = -(ob[6] * hB0 + ob[7] * hB1 + hB2);
  if (!JAC) return;
  // dr/dn = -([x y 1] + r [h0 h1 0]) / s
  double G[4][3];
  G[0][0] = -(ob[0] + r[0] * hA0) * isA; G[0][1] = -(ob[1] + r[0] * hA1) * isA; G[0][2] = -isA;
  G[1][0] = -(ob[2] + r[1] * hA0) * isA; G[1][1] = -(ob[3] + r[1] * hA1) * isA; G[1][2] = -isA;
  G[2][0] = -(ob[4] + r[2] * hB0) * isB; G[2][1] = -(ob[5] + r[2] * hB1) * isB; G[2][2] = -isB;
  G[3][0] = -(ob[6] + r[3] * hB0) * isB; G[3][1] = -(ob[7] + r[3] * hB1) * isB; G[3][2] = -isB;
#define SLSLAM_COL(J, stride, col, a0, a1, a2, b1, b2)                                   \
  {                                                                                      \
    J[0 * stride + col] = G[0][0] * (a0) + G[0][1] * (a1) + G[0][2] * (a2);              \
    J[1 * stride + col] = G[1][0] * (a0) + G[1][1] * (a1) + G[1][2] * (a2);              \
    J[2 * stride + col] = G[2][0] * (a0) + G[2][1] * (b1) + G[2][2] * (b2);              \
    J[3 * stride + col] = G[3][0] * (a0) + G[3][1] * (b1) + G[3][2] * (b2);              \
  }
  // rotation columns: m' = dR_k xh, q' = dR_k yh ; dnA = d m' + t x q' ; dnB = dnA + (0, bl q'2, -bl q'1)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double mp[3], qp[3];
    mv3(cpre + 9 + 9 * k, lt.xh, mp);
    mv3(cpre + 9 + 9 * k, lt.yh, qp);
    const double a0 = lt.d * mp[0] + (t1 * qp[2] - t2 * qp[1]);
    const double a1 = lt.d * mp[1] + (t2 * qp[0] - t0 * qp[2]);
    const double a2 = lt.d * mp[2] + (t0 * qp[1] - t1 * qp[0]);
    SLSLAM_COL(Jc, 6, k, a0, a1, a2, a1 + bl * qp[2], a2 - bl * qp[1]);
  }
  // translation columns: dn = e_j x q, same for both cameras
  SLSLAM_COL(Jc, 6, 3, 0.0, -q[2], q[1], -q[2], q[1]);
  SLSLAM_COL(Jc, 6, 4, q[2], 0.0, -q[0], 0.0, -q[0]);
  SLSLAM_COL(Jc, 6, 5, -q[1], q[0], 0.0, q[0], 0.0);
  {  // a: xh fixed, q' = R zh
    double zz[3];
    mv3(R, lt.zh, zz);
    const double a0 = t1 * zz[2] - t2 * zz[1], a1 = t2 * zz[0] - t0 * zz[2], a2 = t0 * zz[1] - t1 * zz[0];
    SLSLAM_COL(Jl, 4, 0, a0, a1, a2, a1 + bl * zz[2], a2 - bl * zz[1]);
  }
  {  // b: m' = R xb, q' = s1 m
    double mp[3];
    mv3(R, lt.xb, mp);
    const double s1 = lt.s1;
    const double a0 = lt.d * mp[0] + s1 * (t1 * m[2] - t2 * m[1]);
    const double a1 = lt.d * mp[1] + s1 * (t2 * m[0] - t0 * m[2]);
    const double a2 = lt.d * mp[2] + s1 * (t0 * m[1] - t1 * m[0]);
    SLSLAM_COL(Jl, 4, 1, a0, a1, a2, a1 + bl * s1 * m[2], a2 - bl * s1 * m[1]);
  }
  {  // g: m' = R (-xh1, xh0, 0), q' = R (-yh1, yh0, 0)
    const double xg[3] = {-lt.xh[1], lt.xh[0], 0.0}, yg[3] = {-lt.yh[1], lt.yh[0], 0.0};
    double mp[3], qp[3];
    mv3(R, xg, mp);
    mv3(R, yg, qp);
    const double a0 = lt.d * mp[0] + (t1 * qp[2] - t2 * qp[1]);
    const double a1 = lt.d * mp[1] + (t2 * qp[0] - t0 * qp[2]);
    const double a2 = lt.d * mp[2] + (t0 * qp[1] - t1 * qp[0]);
    SLSLAM_COL(Jl, 4, 2, a0, a1, a2, a1 + bl * qp[2], a2 - bl * qp[1]);
  }
  {  // t: d' = -1/sin^2 t
    const double a0 = -lt.ist2 * m[0], a1 = -lt.ist2 * m[1], a2 = -lt.ist2 * m[2];
    SLSLAM_COL(Jl, 4, 3, a0, a1, a2, a1, a2);
  }
#undef SLSLAM_COL
}

// Inverse of a symmetric positive definite 3x3 [a b c; b d e; c e f] by cofactors: one reciprocal instead of three
// dependent square roots.  o = {i00, i01, i02, i11, i12, i22}.  Returns false if a leading minor is not positive.
__device__ __forceinline__ bool spd3_inverse(double a, double b, double c, double d, double e, double f, double* o) {
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  const double det = a * c00 + b * c01 + c * c02;
  const double r = pivot_rcp(det);
  o[0] = c00 * r; o[1] = c01 * r; o[2] = c02 * r; o[3] = c11 * r; o[4] = c12 * r; o[5] = c22 * r;
  return a > 0.0 && c22 > 0.0 && det > 0.0;
}

// HuberLoss(a) on s = |r|^2 with the rho'' <= 0 corrector: returns rho, sets sqrt(rho').
__device__ __forceinline__ double huber_rho(double s, double a, bool robust, double& sqrt_rho1) {
  if (!robust || s <= a * a) { sqrt_rho1 = 1.0; return s; }
  const double ri = pivot_rsqrt(s), rs = s * ri;        // sqrt(s) = s / sqrt(s)
  const double q = a * ri;                              // rho' = a / sqrt(s)
  sqrt_rho1 = q * pivot_rsqrt(q);
  return 2.0 * a * rs - a * a;
}

}  // namespace slslam
