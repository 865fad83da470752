// Small dense FP64 linear algebra kept entirely in registers: Jacobi symmetric
// eigensolves, the rotation-variant SVD used by the analytic eigensystems, and a
// generic cyclic Jacobi for standalone psd_project.
//
// Everything is SKB_HD (host + device) so tests/host_harness.cu can run the exact
// same code on the CPU of the build container (test infrastructure only; the
// product never executes these on the host).
#pragma once
#include <math.h>
#include <float.h>

// Build configuration of the assembly path (every csrc header includes this file first).  The three switches below
// started as A/B experiments of round 1; timed on a B200 in round 2 (C5, profiles/r02c_ab_element_order.txt) they are
// worth 2 % of the step together and are now the default.  -DSKB_BASELINE_KERNELS restores the round-1 kernels.
#if !defined(SKB_BASELINE_KERNELS)
#ifndef SKB_EXP_SUM0
#define SKB_EXP_SUM0 1      // corner-0 pairs of the local stiffness by read-back (sum_a K_ab = 0) instead of U M U^T
#endif
#ifndef SKB_EXP_SRCBASE
#define SKB_EXP_SRCBASE 1   // phase-2 sources carry the pair's staging offset + diagonal flag (no pair_base arithmetic)
#endif
#ifndef SKB_FIN_ITEMS
#define SKB_FIN_ITEMS 2     // level 2: two items per thread, loads of every stage batched
#endif
#endif

#if defined(__CUDACC__)
#define SKB_HD __host__ __device__ __forceinline__
#else
#define SKB_HD inline
#endif

namespace skb {

template <int N>
struct Vec {
  double v[N];
  SKB_HD double& operator[](int i) { return v[i]; }
  SKB_HD const double& operator[](int i) const { return v[i]; }
};

// Row-major N x N matrix in registers (all loops are fully unrolled).
template <int N>
struct Mat {
  double m[N][N];
  SKB_HD double& operator()(int i, int j) { return m[i][j]; }
  SKB_HD const double& operator()(int i, int j) const { return m[i][j]; }
};

template <int N>
SKB_HD Mat<N> identity() {
  Mat<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
  return r;
}

template <int N>
SKB_HD Mat<N> matmul(const Mat<N>& a, const Mat<N>& b) {
  Mat<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(a.m[i][k], b.m[k][j], s);
      r.m[i][j] = s;
    }
  return r;
}

// a * b^T
template <int N>
SKB_HD Mat<N> matmul_nt(const Mat<N>& a, const Mat<N>& b) {
  Mat<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(a.m[i][k], b.m[j][k], s);
      r.m[i][j] = s;
    }
  return r;
}

// a^T * b
template <int N>
SKB_HD Mat<N> matmul_tn(const Mat<N>& a, const Mat<N>& b) {
  Mat<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(a.m[k][i], b.m[k][j], s);
      r.m[i][j] = s;
    }
  return r;
}

SKB_HD double det(const Mat<1>& a) { return a.m[0][0]; }
SKB_HD Mat<1> cofactor(const Mat<1>&) {
  Mat<1> c;
  c.m[0][0] = 1.0;
  return c;
}
SKB_HD double det(const Mat<2>& a) { return a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]; }
SKB_HD double det(const Mat<3>& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// cofactor matrix c = d det / dF
SKB_HD Mat<2> cofactor(const Mat<2>& f) {
  Mat<2> c;
  c.m[0][0] = f.m[1][1];
  c.m[0][1] = -f.m[1][0];
  c.m[1][0] = -f.m[0][1];
  c.m[1][1] = f.m[0][0];
  return c;
}
SKB_HD Mat<3> cofactor(const Mat<3>& f) {
  Mat<3> c;
  c.m[0][0] = f.m[1][1] * f.m[2][2] - f.m[1][2] * f.m[2][1];
  c.m[0][1] = f.m[1][2] * f.m[2][0] - f.m[1][0] * f.m[2][2];
  c.m[0][2] = f.m[1][0] * f.m[2][1] - f.m[1][1] * f.m[2][0];
  c.m[1][0] = f.m[0][2] * f.m[2][1] - f.m[0][1] * f.m[2][2];
  c.m[1][1] = f.m[0][0] * f.m[2][2] - f.m[0][2] * f.m[2][0];
  c.m[1][2] = f.m[0][1] * f.m[2][0] - f.m[0][0] * f.m[2][1];
  c.m[2][0] = f.m[0][1] * f.m[1][2] - f.m[0][2] * f.m[1][1];
  c.m[2][1] = f.m[0][2] * f.m[1][0] - f.m[0][0] * f.m[1][2];
  c.m[2][2] = f.m[0][0] * f.m[1][1] - f.m[0][1] * f.m[1][0];
  return c;
}

// --------------------------------------------------------------------------
// Jacobi rotation (c, s) that zeroes the (p,q) entry of a symmetric matrix
// with diagonal entries app, aqq and off-diagonal apq.  Standard stable form
// (Golub & Van Loan 8.5): t is the smaller root so |theta| <= pi/4.
// --------------------------------------------------------------------------
SKB_HD void sym_schur2(double app, double aqq, double apq, double& c, double& s, double& t) {
  if (apq == 0.0) {
    c = 1.0;
    s = 0.0;
    t = 0.0;
    return;
  }
  double tau = (aqq - app) / (2.0 * apq);
  t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
  c = 1.0 / sqrt(fma(t, t, 1.0));
  s = t * c;
}

// reciprocal square root: MUFU.RSQ64H + Newton steps on the device (no slow-path division)
SKB_HD double rsqrt_f64(double x) {
#if defined(__CUDA_ARCH__)
  // hardware seed (MUFU.RSQ64H, ~22 bits) + one third-order step  y <- y + y e (1/2 + 3/8 e),
  // e = 1 - x y^2: full double accuracy for normal positive x; no range checks or slow path
  // (callers guarantee x > 0 and finite)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#else
  return 1.0 / sqrt(x);
#endif
}

// Same rotation as sym_schur2, division-free: with d = aqq - app, o = 2 apq, h = hypot(d, o),
//   cos 2t = |d| / h,  c = sqrt((1 + cos 2t) / 2),  s = sign(d) o / (2 h c),  t = s / c.
// Two rsqrt and ~12 flops instead of one division and two square roots; no cancellation
// because (1 + cos 2t) / 2 lies in [1/2, 1].
SKB_HD void sym_schur2_fast(double app, double aqq, double apq, double& c, double& s, double& t) {
  const double d = aqq - app, o = apq + apq;
  const double h2 = fma(d, d, o * o);
  if (!(h2 > 0.0) || apq == 0.0) {  // nothing to rotate (also NaN / underflow)
    c = 1.0;
    s = 0.0;
    t = 0.0;
    return;
  }
  const double rh = rsqrt_f64(h2);
  const double x = fma(0.5 * fabs(d), rh, 0.5);
  const double rc = rsqrt_f64(x);
  c = x * rc;
  s = (d >= 0.0 ? 0.5 : -0.5) * o * rh * rc;
  t = s * rc;
}

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix held in
// registers.  On exit  a = V diag(w) V^T,  V orthogonal with det +1.
// All index arithmetic is compile-time (loops unrolled), so nothing spills to
// local memory because of dynamic indexing.
template <int N>
SKB_HD void jacobi_eig(Mat<N> a, Vec<N>& w, Mat<N>& V, int max_sweeps = 12, double tol2 = 1e-32) {
  V = identity<N>();
  // sweeps stay a loop (1-3 run; unrolled, the 12 copies of the sweep body were 40 KB of never-executed instructions)
#pragma unroll 1
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      diag = fma(a.m[i][i], a.m[i][i], diag);
#pragma unroll
      for (int j = i + 1; j < N; ++j) off = fma(a.m[i][j], a.m[i][j], off);
    }
    // converged when the off-diagonal mass is below double rounding of the diagonal;
    // `!(off > ...)` also stops on NaN input
    if (!(off > tol2 * diag) ) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        double c, s, t;
        sym_schur2_fast(a.m[p][p], a.m[q][q], a.m[p][q], c, s, t);
        // A <- J^T A J with J = [[c, s], [-s, c]] on (p,q)
        double apq = a.m[p][q];
        a.m[p][p] = fma(-t, apq, a.m[p][p]);
        a.m[q][q] = fma(t, apq, a.m[q][q]);
        a.m[p][q] = 0.0;
        a.m[q][p] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          if (k != p && k != q) {
            double akp = a.m[k][p], akq = a.m[k][q];
            double np_ = fma(c, akp, -s * akq);
            double nq_ = fma(s, akp, c * akq);
            a.m[k][p] = np_;
            a.m[p][k] = np_;
            a.m[k][q] = nq_;
            a.m[q][k] = nq_;
          }
          double vkp = V.m[k][p], vkq = V.m[k][q];
          V.m[k][p] = fma(c, vkp, -s * vkq);
          V.m[k][q] = fma(s, vkp, c * vkq);
        }
      }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = a.m[i][i];
}

// Runtime-sized variant operating on memory (standalone psd_project for block
// sizes that are not 4 or 9).  a is n x n row-major, overwritten; V n x n.
SKB_HD void jacobi_eig_dyn(double* a, double* w, double* V, int n, int max_sweeps = 30) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += a[i * n + i] * a[i * n + i];
      for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
    }
    if (!(off > 1e-32 * diag)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double c, s, t;
        sym_schur2(a[p * n + p], a[q * n + q], a[p * n + q], c, s, t);
        double apq = a[p * n + q];
        a[p * n + p] -= t * apq;
        a[q * n + q] += t * apq;
        a[p * n + q] = 0.0;
        a[q * n + p] = 0.0;
        for (int k = 0; k < n; ++k) {
          if (k != p && k != q) {
            double akp = a[k * n + p], akq = a[k * n + q];
            double np_ = c * akp - s * akq;
            double nq_ = s * akp + c * akq;
            a[k * n + p] = np_;
            a[p * n + k] = np_;
            a[k * n + q] = nq_;
            a[q * n + k] = nq_;
          }
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
}

// swap columns p,q of A and V, negating the new column q of both (keeps
// A V^T and det V unchanged)
template <int N>
SKB_HD void swap_cols_signed(Mat<N>& A, Mat<N>& V, Vec<N>& nrm, int p, int q, bool doit) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double ap = A.m[k][p], aq = A.m[k][q];
    A.m[k][p] = doit ? aq : ap;
    A.m[k][q] = doit ? -ap : aq;
    double vp = V.m[k][p], vq = V.m[k][q];
    V.m[k][p] = doit ? vq : vp;
    V.m[k][q] = doit ? -vp : vq;
  }
  double np_ = nrm[p], nq_ = nrm[q];
  nrm[p] = doit ? nq_ : np_;
  nrm[q] = doit ? np_ : nq_;
}

// --------------------------------------------------------------------------
// Rotation-variant SVD  F = U diag(sig) V^T  with U, V proper rotations, the
// singular values ordered by decreasing magnitude and the sign of det F carried
// by the LAST one -- the convention of the reference's svd_rv
// (/root/reference/simkit/svd_rv.py:33-53, polar_svd.py:34-56).
//
// Method: Jacobi eigensolve of C = F^T F for V, then A = F V and ONE one-sided
// (Hestenes) clean-up sweep on the columns of A, which restores high relative
// accuracy of the small singular triplets.  U's last column is the cross product
// of the others, so the result is well defined for rank-deficient and inverted F.
// --------------------------------------------------------------------------
template <int N>
SKB_HD void hestenes_sweep(Mat<N>& A, Mat<N>& V) {
#pragma unroll
  for (int p = 0; p < N - 1; ++p)
#pragma unroll
    for (int q = p + 1; q < N; ++q) {
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        al = fma(A.m[k][p], A.m[k][p], al);
        be = fma(A.m[k][q], A.m[k][q], be);
        ga = fma(A.m[k][p], A.m[k][q], ga);
      }
      if (ga * ga > 1e-34 * al * be) {
        double c, s, t;
        sym_schur2_fast(al, be, ga, c, s, t);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double ap = A.m[k][p], aq = A.m[k][q];
          A.m[k][p] = fma(c, ap, -s * aq);
          A.m[k][q] = fma(s, ap, c * aq);
          double vp = V.m[k][p], vq = V.m[k][q];
          V.m[k][p] = fma(c, vp, -s * vq);
          V.m[k][q] = fma(s, vp, c * vq);
        }
      }
    }
}

// F^T F, computing only the upper triangle
template <int N>
SKB_HD Mat<N> gram(const Mat<N>& F) {
  Mat<N> C;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(F.m[k][i], F.m[k][j], s);
      C.m[i][j] = s;
      C.m[j][i] = s;
    }
  return C;
}

// reciprocal: MUFU.RCP64H seed + one third-order step  y <- y (1 + e + e^2),  e = 1 - x y  (full double
// accuracy for normal x, 1/0 = inf and 1/inf = 0 kept; no denormal slow path)
SKB_HD double rcp_f64(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return (e == e) ? fma(y, fma(e, e, e), y) : y;  // x = 0 / inf / NaN: the seed is already the answer
#else
  return 1.0 / x;
#endif
}

SKB_HD float rsqrt_f32(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}

// FP32 warm start of the right singular vectors: cyclic Jacobi on the single-precision Gram matrix
// of F (the FP32 pipe runs beside the FP64 pipe, so these sweeps are nearly free), stopped at a
// relative off-diagonal mass of ~3e-7.  Returns V as a single-precision matrix that is orthogonal to
// ~1e-6; the caller re-orthonormalises in double.  NaN / overflow / zero input leaves V = I (the
// double-precision sweeps that follow then do all the work).
SKB_HD void jacobi_warm_f32(const Mat<3>& F, float V[3][3]) {
  float f[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) f[i][j] = (float)F.m[i][j];
  float a[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      float s = f[0][i] * f[0][j];
      s = fmaf(f[1][i], f[1][j], s);
      s = fmaf(f[2][i], f[2][j], s);
      a[i][j] = s;
      a[j][i] = s;
    }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0f : 0.0f;
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    const float off = fmaf(a[0][1], a[0][1], fmaf(a[0][2], a[0][2], a[1][2] * a[1][2]));
    const float dg = fmaf(a[0][0], a[0][0], fmaf(a[1][1], a[1][1], a[2][2] * a[2][2]));
    if (!(off > 1e-13f * dg) || !(dg < 3e38f)) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const float apq = a[p][q];
        const float d = a[q][q] - a[p][p], o = apq + apq;
        const float h2 = fmaf(d, d, o * o);
        float c = 1.0f, s = 0.0f, t = 0.0f;
        if (h2 > 1e-37f && h2 < 3e38f) {
          const float rh = rsqrt_f32(h2);
          const float x = fmaf(0.5f * fabsf(d), rh, 0.5f);
          const float rc = rsqrt_f32(x);
          c = x * rc;
          s = (d >= 0.0f ? 0.5f : -0.5f) * o * rh * rc;
          t = s * rc;
        }
        a[p][p] = fmaf(-t, apq, a[p][p]);
        a[q][q] = fmaf(t, apq, a[q][q]);
        a[p][q] = 0.0f;
        a[q][p] = 0.0f;
        const int k = 3 - p - q;
        const float akp = a[k][p], akq = a[k][q];
        const float np_ = fmaf(c, akp, -s * akq), nq_ = fmaf(s, akp, c * akq);
        a[k][p] = np_;
        a[p][k] = np_;
        a[k][q] = nq_;
        a[q][k] = nq_;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float vp = V[r][p], vq = V[r][q];
          V[r][p] = fmaf(c, vp, -s * vq);
          V[r][q] = fmaf(s, vp, c * vq);
        }
      }
  }
}

// One-sided (Hestenes) sweep over the column pairs of A = F V.  A pair is rotated when its relative
// inner product exceeds sqrt(thr2); returns whether anything was rotated.
template <int N>
SKB_HD bool hestenes_sweep_thr(Mat<N>& A, Mat<N>& V, double thr2) {
  bool any = false;
#pragma unroll
  for (int p = 0; p < N - 1; ++p)
#pragma unroll
    for (int q = p + 1; q < N; ++q) {
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        al = fma(A.m[k][p], A.m[k][p], al);
        be = fma(A.m[k][q], A.m[k][q], be);
        ga = fma(A.m[k][p], A.m[k][q], ga);
      }
      if (ga * ga > thr2 * al * be) {
        any = true;
        double c, s, t;
        sym_schur2_fast(al, be, ga, c, s, t);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double ap = A.m[k][p], aq = A.m[k][q];
          A.m[k][p] = fma(c, ap, -s * aq);
          A.m[k][q] = fma(s, ap, c * aq);
          double vp = V.m[k][p], vq = V.m[k][q];
          V.m[k][p] = fma(c, vp, -s * vq);
          V.m[k][q] = fma(s, vp, c * vq);
        }
      }
    }
  return any;
}

// 3x3 rotation-variant SVD.  V starts from the FP32 Jacobi warm start (re-orthonormalised in double by
// Gram-Schmidt + cross product, so det V = +1 to rounding), then one-sided Hestenes sweeps on the
// columns of A = F V in double: the first always rotates (quadratic convergence takes the ~1e-6 warm
// start to ~1e-12), later sweeps only touch pairs whose relative inner product still exceeds 1e-13 and
// the loop ends when a sweep rotates nothing.  Working on the columns themselves keeps high relative
// accuracy of the small singular triplets (inverted / flat elements).
SKB_HD void svd_rv(const Mat<3>& F, Mat<3>& U, Vec<3>& sig, Mat<3>& V) {
#if defined(SKB_SVD_F64)
#define SKB_SVD_JACOBI_TOL2 1e-22
  Mat<3> C = gram(F);
  Vec<3> w;
  jacobi_eig<3>(C, w, V, 12, SKB_SVD_JACOBI_TOL2);
  Mat<3> A = matmul(F, V);
  hestenes_sweep<3>(A, V);
#elif defined(SKB_EXP_NOSVD)  // timing experiment only (wrong results): no warm start, no sweeps
  V = identity<3>();
  Mat<3> A = F;
#else
  {
    float Vf[3][3];
    jacobi_warm_f32(F, Vf);
    double v0[3], v1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v0[k] = (double)Vf[k][0];
      v1[k] = (double)Vf[k][1];
    }
    const double r0 = rsqrt_f64(fma(v0[0], v0[0], fma(v0[1], v0[1], v0[2] * v0[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k) v0[k] *= r0;
    const double d01 = fma(v0[0], v1[0], fma(v0[1], v1[1], v0[2] * v1[2]));
#pragma unroll
    for (int k = 0; k < 3; ++k) v1[k] = fma(-d01, v0[k], v1[k]);
    const double r1 = rsqrt_f64(fma(v1[0], v1[0], fma(v1[1], v1[1], v1[2] * v1[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k) v1[k] *= r1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      V.m[k][0] = v0[k];
      V.m[k][1] = v1[k];
    }
    V.m[0][2] = v0[1] * v1[2] - v0[2] * v1[1];
    V.m[1][2] = v0[2] * v1[0] - v0[0] * v1[2];
    V.m[2][2] = v0[0] * v1[1] - v0[1] * v1[0];
  }
  Mat<3> A = matmul(F, V);
  // the first sweep always rotates (threshold 1e-34), later ones only what is left above 1e-13 relative; ONE copy of
  // the sweep body in the instruction stream (the per-element code is ~100 KB: instruction fetch shows in the profile)
#pragma unroll 1
  for (int it = 0; it < 13; ++it)
    if (!hestenes_sweep_thr<3>(A, V, it == 0 ? 1e-34 : 1e-26) && it > 0) break;
#endif
  Vec<3> n2;  // squared column norms; square roots are taken once, through rsqrt
#pragma unroll
  for (int j = 0; j < 3; ++j) n2[j] = fma(A.m[0][j], A.m[0][j], fma(A.m[1][j], A.m[1][j], A.m[2][j] * A.m[2][j]));
  // sort columns by decreasing norm (3-element network)
  swap_cols_signed<3>(A, V, n2, 0, 1, n2[0] < n2[1]);
  swap_cols_signed<3>(A, V, n2, 1, 2, n2[1] < n2[2]);
  swap_cols_signed<3>(A, V, n2, 0, 1, n2[0] < n2[1]);
  // leading two columns of U by normalisation (Gram-Schmidt guards rank <= 1)
  double u0[3], u1[3], u2[3];
  double s0 = 0.0, s1 = 0.0;
  if (n2[0] > 0.0) {
    const double inv = rsqrt_f64(n2[0]);
    s0 = n2[0] * inv;
#pragma unroll
    for (int k = 0; k < 3; ++k) u0[k] = A.m[k][0] * inv;
  } else {
    u0[0] = 1.0; u0[1] = 0.0; u0[2] = 0.0;
  }
  if (n2[1] > 1e-300 * n2[0] && n2[1] > 0.0) {
    const double inv = rsqrt_f64(n2[1]);
    s1 = n2[1] * inv;
#pragma unroll
    for (int k = 0; k < 3; ++k) u1[k] = A.m[k][1] * inv;
    // re-orthogonalise against u0 (no-op to rounding when converged)
    double d = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) u1[k] = fma(-d, u0[k], u1[k]);
    double n1 = rsqrt_f64(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) u1[k] *= n1;
  } else {
    // any unit vector orthogonal to u0
    int kmin = (fabs(u0[0]) <= fabs(u0[1]) && fabs(u0[0]) <= fabs(u0[2])) ? 0 : (fabs(u0[1]) <= fabs(u0[2]) ? 1 : 2);
    double e[3] = {kmin == 0 ? 1.0 : 0.0, kmin == 1 ? 1.0 : 0.0, kmin == 2 ? 1.0 : 0.0};
    double d = u0[kmin];
#pragma unroll
    for (int k = 0; k < 3; ++k) u1[k] = e[k] - d * u0[k];
    double n1 = rsqrt_f64(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) u1[k] *= n1;
  }
  u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
  u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
  u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    U.m[k][0] = u0[k];
    U.m[k][1] = u1[k];
    U.m[k][2] = u2[k];
  }
  sig[0] = s0;
  sig[1] = s1;
  sig[2] = u2[0] * A.m[0][2] + u2[1] * A.m[1][2] + u2[2] * A.m[2][2];  // signed
}

SKB_HD void svd_rv(const Mat<2>& F, Mat<2>& U, Vec<2>& sig, Mat<2>& V) {
  Mat<2> C = gram(F);
  Vec<2> w;
  jacobi_eig<2>(C, w, V, 2);
  Mat<2> A = matmul(F, V);
  hestenes_sweep<2>(A, V);
  Vec<2> nrm;
#pragma unroll
  for (int j = 0; j < 2; ++j) nrm[j] = sqrt(A.m[0][j] * A.m[0][j] + A.m[1][j] * A.m[1][j]);
  swap_cols_signed<2>(A, V, nrm, 0, 1, nrm[0] < nrm[1]);
  double u0[2];
  if (nrm[0] > 0.0) {
    double inv = 1.0 / nrm[0];
    u0[0] = A.m[0][0] * inv;
    u0[1] = A.m[1][0] * inv;
  } else {
    u0[0] = 1.0;
    u0[1] = 0.0;
  }
  double u1[2] = {-u0[1], u0[0]};  // +90 degrees: det U = +1
  U.m[0][0] = u0[0];
  U.m[1][0] = u0[1];
  U.m[0][1] = u1[0];
  U.m[1][1] = u1[1];
  sig[0] = nrm[0];
  sig[1] = u1[0] * A.m[0][1] + u1[1] * A.m[1][1];
}

// symmetric rebuild  V diag(w) V^T
template <int N>
SKB_HD Mat<N> rebuild_sym(const Mat<N>& V, const Vec<N>& w) {
  Mat<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(V.m[i][k] * w[k], V.m[j][k], s);
      r.m[i][j] = s;
      r.m[j][i] = s;
    }
  return r;
}

}  // namespace skb
