// Constitutive models of the hot path, per element, in FP64 registers.
//
// Two views of each model:
//   * direct formulas for psi(F) and P(F) = dpsi/dF (energy / gradient kernels,
//     no decomposition except ARAP's polar factor);
//   * the principal-stretch view used for the Hessian: with the rotation-variant
//     SVD F = U diag(sig) V^T every isotropic model's d2psi/dF2 is block
//     diagonal in the basis {U e_p e_q^T V^T}: a d x d "scaling" block
//     S_pr = d2psi/dsig_p dsig_r and, per pair (p,q), a 2x2 block [[a,b],[b,a]]
//     whose eigenvalues are the flip (a+b) and twist (a-b) eigenvalues
//     (SURVEY.md §7 "Analytic eigensystem").  PSD projection clamps those
//     eigenvalues instead of running a 9x9 eigh per element
//     (/root/reference/simkit/psd_project.py:28-45).
//
// Reference formulas restated here:
//   stable neo-Hookean  energies/stable_neo_hookean.py:65-129,132-218,221-443
//   neo-Hookean         energies/neo_hookean.py:64-97,100-131,134-178
//   ARAP                energies/arap.py:71-145, rotation_gradient.py:12-75
//   StVK                energies/stvk.py:61-93,96-127,130-179
//   linear elasticity   energies/linear_elasticity.py:43-70,73-100,103-136
//   FCR                 energies/fcr.py:41-62,65-125,128-302 (ARAP shear part x2 + volumetric part)
//   Macklin-Mueller NH  energies/macklin_mueller_neo_hookean.py:66-122,125-185,188-370
#pragma once
#include "smallmat.cuh"

namespace skb {

enum Material : int {
  MAT_STABLE_NEO_HOOKEAN = 0,
  MAT_NEO_HOOKEAN = 1,
  MAT_ARAP = 2,
  MAT_STVK = 3,
  MAT_LINEAR_ELASTICITY = 4,
  MAT_FCR = 5,                          // fixed corotational: 2 psi_arap + lam/2 (J-1)^2
  MAT_MACKLIN_MUELLER_NEO_HOOKEAN = 6,  // mu (1-J) + lam/2 (1-J)^2 + mu/2 (I_C - d)
  MAT_COUNT = 7
};

// models whose stress needs the polar rotation R = U V^T
SKB_HD bool material_uses_rotation(int mat) { return mat == MAT_ARAP || mat == MAT_FCR; }

// psd_mode: how eigenvalues are floored relative to the quadrature weight
enum PsdMode : int {
  PSD_NONE = 0,
  PSD_AFTER_VOL = 1,   // per-material *_hessian_x/_u: psd_project(vol * He)   (stable_neo_hookean.py:533-535)
  PSD_BEFORE_VOL = 2,  // elastic dispatcher / _z tier: vol * psd_project(He)   (elastic.py:663-664)
  PSD_ABS_AFTER_VOL = 3  // psd_project(method='abs') semantics on vol*He (psd_project.py:34-35)
};

#define SKB_PSD_FLOOR 1e-6  // psd_project.py:37

template <int D>
SKB_HD double frob2(const Mat<D>& F) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) s = fma(F.m[i][j], F.m[i][j], s);
  return s;
}

template <int D>
SKB_HD Mat<D> polar_R(const Mat<D>& F) {
  Mat<D> U, V;
  Vec<D> sig;
  svd_rv(F, U, sig, V);
  return matmul_nt(U, V);
}

// ---------------------------------------------------------------- energy ---
template <int D>
SKB_HD double energy_density(int mat, const Mat<D>& F, double mu, double lam) {
  switch (mat) {
    case MAT_STABLE_NEO_HOOKEAN: {
      double IC = frob2(F);
      double J = det(F);
      double alpha = 1.0 + D * mu * rcp_f64((D + 1) * lam);
      double d = J - alpha;
      return 0.5 * mu * (IC - D) - 0.5 * mu * log(IC + 1.0) + 0.5 * lam * d * d;
    }
    case MAT_NEO_HOOKEAN: {
      double IC = frob2(F);
      double lJ = log(det(F));
      return 0.5 * mu * (IC - D) - mu * lJ + 0.5 * lam * lJ * lJ;
    }
    case MAT_ARAP: {
      Mat<D> R = polar_R(F);
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double d = F.m[i][j] - R.m[i][j];
          s = fma(d, d, s);
        }
      return 0.5 * mu * s;
    }
    case MAT_FCR: {
      Mat<D> R = polar_R(F);
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double d = F.m[i][j] - R.m[i][j];
          s = fma(d, d, s);
        }
      double dj = det(F) - 1.0;
      return mu * s + 0.5 * lam * dj * dj;
    }
    case MAT_MACKLIN_MUELLER_NEO_HOOKEAN: {
      double IC = frob2(F);
      double dj = 1.0 - det(F);
      return mu * dj + 0.5 * lam * dj * dj + 0.5 * mu * (IC - D);
    }
    case MAT_STVK: {
      Mat<D> C = matmul_tn(F, F);
      double tr = 0.0, e2 = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double e = 0.5 * (C.m[i][j] - (i == j ? 1.0 : 0.0));
          e2 = fma(e, e, e2);
          if (i == j) tr += e;
        }
      return mu * e2 + 0.5 * lam * tr * tr;
    }
    default: {  // MAT_LINEAR_ELASTICITY
      double tr = 0.0, e2 = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double e = 0.5 * (F.m[i][j] + F.m[j][i]) - (i == j ? 1.0 : 0.0);
          e2 = fma(e, e, e2);
          if (i == j) tr += e;
        }
      return mu * e2 + 0.5 * lam * tr * tr;
    }
  }
}

// ------------------------------------------------------------------ PK1 ----
template <int D>
SKB_HD Mat<D> pk1(int mat, const Mat<D>& F, double mu, double lam) {
  Mat<D> P;
  switch (mat) {
    case MAT_STABLE_NEO_HOOKEAN: {
      double IC = frob2(F);
      double J = det(F);
      Mat<D> c = cofactor(F);
      double alpha = 1.0 + D * mu * rcp_f64((D + 1) * lam);
      double A = mu * (1.0 - rcp_f64(IC + 1.0));
      double Dc = lam * (J - alpha);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = fma(A, F.m[i][j], Dc * c.m[i][j]);
      return P;
    }
    case MAT_NEO_HOOKEAN: {
      double J = det(F);
      Mat<D> c = cofactor(F);
      double k = (lam * log(J) - mu) * rcp_f64(J);  // F^-T = cof / J
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = fma(mu, F.m[i][j], k * c.m[i][j]);
      return P;
    }
    case MAT_ARAP: {
      Mat<D> R = polar_R(F);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = mu * (F.m[i][j] - R.m[i][j]);
      return P;
    }
    case MAT_FCR: {
      Mat<D> R = polar_R(F);
      Mat<D> c = cofactor(F);
      double k = lam * (det(F) - 1.0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = fma(2.0 * mu, F.m[i][j] - R.m[i][j], k * c.m[i][j]);
      return P;
    }
    case MAT_MACKLIN_MUELLER_NEO_HOOKEAN: {
      Mat<D> c = cofactor(F);
      double k = lam * (det(F) - 1.0) - mu;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = fma(mu, F.m[i][j], k * c.m[i][j]);
      return P;
    }
    case MAT_STVK: {
      Mat<D> C = matmul_tn(F, F);
      Mat<D> S2;
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += 0.5 * (C.m[i][i] - 1.0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j)
          S2.m[i][j] = mu * (C.m[i][j] - (i == j ? 1.0 : 0.0)) + (i == j ? lam * tr : 0.0);
      return matmul(F, S2);
    }
    default: {  // linear elasticity
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += F.m[i][i] - 1.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j)
          P.m[i][j] = mu * (F.m[i][j] + F.m[j][i] - (i == j ? 2.0 : 0.0)) + (i == j ? lam * tr : 0.0);
      return P;
    }
  }
}

// ------------------------------------------------ principal-stretch view ---
// NP = number of index pairs (p<q): 1 in 2D, 3 in 3D.  Pair k of 3D is
// (0,1),(0,2),(1,2); its "third" index is 2,1,0.
template <int D>
struct Principal {
  static constexpr int NP = D * (D - 1) / 2;
  Mat<D> S;        // scaling block d2psi/dsig dsig (symmetric)
  double a[NP];    // 2x2 pair block diagonal   ( (flip+twist)/2 )
  double b[NP];    // 2x2 pair block off-diag   ( (flip-twist)/2 )
};

template <int D>
SKB_HD void pair_index(int k, int& p, int& q, int& r) {
  if (D == 2) {
    p = 0; q = 1; r = 0;
  } else {
    p = (k == 2) ? 1 : 0;
    q = (k == 0) ? 1 : 2;
    r = 3 - p - q;
  }
}

// Unweighted Hessian of the isotropic models in the SVD frame.
template <int D>
SKB_HD Principal<D> principal_hessian(int mat, const Vec<D>& sig, double mu, double lam) {
  constexpr int NP = Principal<D>::NP;
  Principal<D> h;
  double twist[NP], flip[NP];
  double IC = 0.0, J = 1.0;
#pragma unroll
  for (int p = 0; p < D; ++p) {
    IC = fma(sig[p], sig[p], IC);
    J *= sig[p];
  }
  switch (mat) {
    case MAT_STABLE_NEO_HOOKEAN: {
      double alpha = 1.0 + D * mu * rcp_f64((D + 1) * lam);
      double ric = rcp_f64(IC + 1.0);
      double A = mu * (1.0 - ric);
      double B = 2.0 * mu * ric * ric;
      double Dc = lam * (J - alpha);
      double ch[D];  // dJ/dsig_p = product of the other stretches
#pragma unroll
      for (int p = 0; p < D; ++p) {
        double c = 1.0;
#pragma unroll
        for (int q = 0; q < D; ++q)
          if (q != p) c *= sig[q];
        ch[p] = c;
      }
#pragma unroll
      for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) {
          double v = B * sig[p] * sig[q] + lam * ch[p] * ch[q];
          if (p == q) v += A;
          else v += Dc * (D == 2 ? 1.0 : sig[3 - p - q]);
          h.S.m[p][q] = v;
        }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        int p, q, r;
        pair_index<D>(k, p, q, r);
        double third = (D == 2) ? 1.0 : sig[r];
        twist[k] = A + Dc * third;
        flip[k] = A - Dc * third;
      }
      break;
    }
    case MAT_NEO_HOOKEAN: {
      double c1 = lam * log(J) - mu;
      double rs[D];
#pragma unroll
      for (int p = 0; p < D; ++p) rs[p] = rcp_f64(sig[p]);
#pragma unroll
      for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) {
          double inv = rs[p] * rs[q];
          h.S.m[p][q] = (p == q) ? (mu + (lam - c1) * inv) : (lam * inv);
        }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        int p, q, r;
        pair_index<D>(k, p, q, r);
        double inv = c1 * rs[p] * rs[q];
        twist[k] = mu + inv;
        flip[k] = mu - inv;
      }
      break;
    }
    case MAT_ARAP: {
      const double clampv = (D == 2) ? 1e-12 : 1e-8;  // rotation_gradient.py:38,65-67
#pragma unroll
      for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) h.S.m[p][q] = (p == q) ? mu : 0.0;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        int p, q, r;
        pair_index<D>(k, p, q, r);
        double den = fmax(sig[p] + sig[q], clampv);
        twist[k] = mu * (1.0 - 2.0 * rcp_f64(den));
        flip[k] = mu;
      }
      break;
    }
    case MAT_FCR:
    case MAT_MACKLIN_MUELLER_NEO_HOOKEAN: {
      // psi = shear(sig) + f(J):  S_pq = shear_pp d_pq + f'' ch_p ch_q + (p != q) f' sig_third,
      // pair block: twist = tw_shear + f' sig_third, flip = fl_shear - f' sig_third  (d2J/dF2 is the
      // pure off-diagonal -sig_third on each pair).  FCR: f = lam/2 (J-1)^2, shear = 2 ARAP;
      // Macklin-Mueller: f = mu (1-J) + lam/2 (1-J)^2, shear = mu/2 I_C.
      const bool fcr = (mat == MAT_FCR);
      const double f1 = fcr ? lam * (J - 1.0) : lam * (J - 1.0) - mu;
      const double sh = fcr ? 2.0 * mu : mu;
      const double clampv = (D == 2) ? 1e-12 : 1e-8;  // rotation_gradient.py:38,65-67
      double ch[D];
#pragma unroll
      for (int p = 0; p < D; ++p) {
        double c = 1.0;
#pragma unroll
        for (int q = 0; q < D; ++q)
          if (q != p) c *= sig[q];
        ch[p] = c;
      }
#pragma unroll
      for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) {
          double v = lam * ch[p] * ch[q];
          if (p == q) v += sh;
          else v += f1 * (D == 2 ? 1.0 : sig[3 - p - q]);
          h.S.m[p][q] = v;
        }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        int p, q, r;
        pair_index<D>(k, p, q, r);
        const double third = (D == 2) ? 1.0 : sig[r];
        double tw = sh;
        if (fcr) tw = sh * (1.0 - 2.0 * rcp_f64(fmax(sig[p] + sig[q], clampv)));
        twist[k] = tw + f1 * third;
        flip[k] = sh - f1 * third;
      }
      break;
    }
    default: {  // MAT_STVK
      double trE = 0.5 * (IC - D);
#pragma unroll
      for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) {
          double v = lam * sig[p] * sig[q];
          if (p == q) v += mu * (sig[p] * sig[p] - 1.0) + lam * trE + 2.0 * mu * sig[p] * sig[p];
          h.S.m[p][q] = v;
        }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        int p, q, r;
        pair_index<D>(k, p, q, r);
        double base = lam * trE + mu * (sig[p] * sig[p] + sig[q] * sig[q] - 1.0);
        double cross = mu * sig[p] * sig[q];
        twist[k] = base - cross;
        flip[k] = base + cross;
      }
      break;
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    h.a[k] = 0.5 * (flip[k] + twist[k]);
    h.b[k] = 0.5 * (flip[k] - twist[k]);
  }
  return h;
}

SKB_HD double psd_clamp(double ev, int mode) {
  if (mode == PSD_ABS_AFTER_VOL) return fabs(ev);
  return (ev < SKB_PSD_FLOOR) ? SKB_PSD_FLOOR : ev;  // NaN stays NaN
}

// Apply the quadrature weight and the eigenvalue floor in the SVD frame.
//   PSD_AFTER_VOL : eig(vol*H) floored     PSD_BEFORE_VOL : vol * floored eig(H)
template <int D>
SKB_HD void weight_and_project(Principal<D>& h, double vol, int psd_mode) {
  constexpr int NP = Principal<D>::NP;
  const double pre = (psd_mode == PSD_BEFORE_VOL) ? 1.0 : vol;
  const double post = (psd_mode == PSD_BEFORE_VOL) ? vol : 1.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    double fl = (h.a[k] + h.b[k]) * pre;
    double tw = (h.a[k] - h.b[k]) * pre;
    if (psd_mode != PSD_NONE) {
      fl = psd_clamp(fl, psd_mode);
      tw = psd_clamp(tw, psd_mode);
    }
    h.a[k] = 0.5 * (fl + tw) * post;
    h.b[k] = 0.5 * (fl - tw) * post;
  }
#pragma unroll
  for (int p = 0; p < D; ++p)
#pragma unroll
    for (int q = 0; q < D; ++q) h.S.m[p][q] *= pre;
  if (psd_mode != PSD_NONE) {
    // Cheap exact test first: S - floor*I positive definite (Cholesky pivots > 0)
    // means every eigenvalue already exceeds the floor and the projection is the
    // identity.  Otherwise eigendecompose the d x d block.
    bool need = false;
    if (psd_mode == PSD_ABS_AFTER_VOL) {
      need = true;
    } else {
      Mat<D> L = h.S;
#pragma unroll
      for (int p = 0; p < D; ++p) L.m[p][p] -= SKB_PSD_FLOOR;
#pragma unroll
      for (int p = 0; p < D; ++p) {
        double piv = L.m[p][p];
        if (!(piv > 0.0)) need = true;
        double inv = rcp_f64(piv);
#pragma unroll
        for (int q = p + 1; q < D; ++q) {
          double f = L.m[q][p] * inv;
#pragma unroll
          for (int r = q; r < D; ++r) L.m[q][r] = fma(-f, L.m[p][r], L.m[q][r]);
#pragma unroll
          for (int r = q; r < D; ++r) L.m[r][q] = L.m[q][r];
        }
      }
    }
    if (need) {
      Vec<D> w;
      Mat<D> Q;
      jacobi_eig<D>(h.S, w, Q);
#pragma unroll
      for (int p = 0; p < D; ++p) w[p] = psd_clamp(w[p], psd_mode);
      h.S = rebuild_sym(Q, w);
    }
  }
#pragma unroll
  for (int p = 0; p < D; ++p)
#pragma unroll
    for (int q = 0; q < D; ++q) h.S.m[p][q] *= post;
}

// Full d2psi/dF2 (b x b, row-major F layout, b = D*D) from the SVD-frame blocks:
//   H[(i,j),(k,l)] = sum U_ip V_jq Hhat[(p,q),(r,s)] U_kr V_ls
// Used by the element-tier entry points and psd_project-free debugging; the
// assembly kernels never materialise it.
template <int D>
SKB_HD void expand_hessian(const Principal<D>& h, const Mat<D>& U, const Mat<D>& V, double* H /* b*b */) {
  constexpr int B = D * D;
  constexpr int NP = Principal<D>::NP;
#pragma unroll
  for (int i = 0; i < B * B; ++i) H[i] = 0.0;
  // scaling block: basis d_p = vec(u_p v_p^T)
  double dvec[D][B];
#pragma unroll
  for (int p = 0; p < D; ++p)
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) dvec[p][i * D + j] = U.m[i][p] * V.m[j][p];
#pragma unroll
  for (int p = 0; p < D; ++p)
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double s = h.S.m[p][r];
      for (int x = 0; x < B; ++x)
        for (int y = 0; y < B; ++y) H[x * B + y] = fma(s * dvec[p][x], dvec[r][y], H[x * B + y]);
    }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    int p, q, r;
    pair_index<D>(k, p, q, r);
    double e1[B], e2[B];  // vec(u_p v_q^T), vec(u_q v_p^T)
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e1[i * D + j] = U.m[i][p] * V.m[j][q];
        e2[i * D + j] = U.m[i][q] * V.m[j][p];
      }
    for (int x = 0; x < B; ++x)
      for (int y = 0; y < B; ++y)
        H[x * B + y] += h.a[k] * (e1[x] * e1[y] + e2[x] * e2[y]) + h.b[k] * (e1[x] * e2[y] + e2[x] * e1[y]);
  }
}

// linear elasticity: constant Hessian  mu (I + T) + lam tr^T tr
template <int D>
SKB_HD void linear_elasticity_hessian(double mu, double lam, double* H) {
  constexpr int B = D * D;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int k = 0; k < D; ++k)
#pragma unroll
        for (int l = 0; l < D; ++l) {
          double v = 0.0;
          if (i == k && j == l) v += mu;
          if (i == l && j == k) v += mu;
          if (i == j && k == l) v += lam;
          H[(i * D + j) * B + (k * D + l)] = v;
        }
}

}  // namespace skb
