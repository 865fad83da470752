// Hot-path kernels: fused gather -> F -> energy / stress / PSD-projected stiffness
// -> J^T H J local block -> deterministic two-level reduction into the fixed
// CSR pattern (no FP64 atomics).
//
// Replaces, per call, the reference chain (e.g. energies/stable_neo_hookean.py:530-538):
//   J @ x  ->  *_hessian_element_F  ->  * vol  ->  psd_project (LAPACK eigh per block)
//   ->  scipy.sparse.block_diag  ->  J^T @ H @ J   (two SpGEMMs)
//
// Kernel bodies are SKB_HD "phase" functions so tests/host_harness.cu can replay
// them on the CPU of the GPU-less build container; the __global__ wrappers below
// only add the thread mapping and shared-memory staging.
#pragma once
#include "materials.cuh"
#include "plan.cuh"

namespace skb {

struct EvalArgs {
  int material;
  int psd_mode;
  const double* x;     // (n*dim) positions or displacements
  const double* Fbar;  // (t*dim*dim) per-element offset of the _u tier (CALLER's element order), or nullptr
  const int* eorder;   // internal element -> caller's element (plan.cuh str_element_order), or nullptr = identity
  const double* mu;
  const double* lam;
  const double* vol;
  int mu_stride, lam_stride, vol_stride;  // 0 = scalar broadcast, 1 = per element
  int want_grad, want_hess;
  double* pblocks;  // [blocks.n_ts][dim*dim] partial block records
  double* pverts;   // [verts.n_ts][dim]      partial vertex records
  double* vals;     // (nnz) CSR values in canonical order
  double* g;        // (n*dim)
};

// Doubles per partial block record: the D*D values padded to whole 32-byte sectors (12 for tets, 4 for
// triangles), so that phase 2 writes every record with 32-byte stores (STG.E.ENL2.256 on sm_100) -- one L2
// write request per sector and no partially written sector to read-modify-write.
template <int D>
struct RecStride {
  static constexpr int value = ((D * D + 3) / 4) * 4;
};
#ifndef SKB_P2_UNROLL
#define SKB_P2_UNROLL 2   // two contributions' loads in flight per lane (r02ai: 3.77 -> 3.71 ms; same summation order)
#endif
constexpr int kP2Unroll = SKB_P2_UNROLL;  // unroll factor of the phase-2 contribution loops

// Staging layout of one element's local stiffness: the K(K+1)/2 corner pairs (a <= b, row-major over the
// upper triangle) one after the other; an off-diagonal pair holds its D x D block row-major (D*D values), a
// diagonal pair the upper triangle of its symmetric block (D(D+1)/2 values).  78 values per tet, 21 per
// triangle -- the packed upper triangle of the local matrix, regrouped so that phase 2 reads a pair's values
// at compile-time offsets from one base address.
template <int K>
SKB_HD constexpr int diag_pair(int a) { return a * K - (a * (a - 1)) / 2; }  // pair index of (a, a)

template <int K>
SKB_HD constexpr int pair_of(int a, int b) { return a * K - (a * (a - 1)) / 2 + (b - a); }  // a <= b

template <int D>
SKB_HD int pair_base(int pp) {
  constexpr int K = D + 1, DD = D * D, DS = D * (D + 1) / 2;
  int nd = 0;  // diagonal pairs before pp
#pragma unroll
  for (int a = 0; a < K - 1; ++a) nd += (pp > diag_pair<K>(a)) ? 1 : 0;
  return DD * pp - (DD - DS) * nd;
}

template <int D>
SKB_HD bool pair_is_diag(int pp) {
  constexpr int K = D + 1;
  bool d = false;
#pragma unroll
  for (int a = 0; a < K; ++a) d = d || (pp == diag_pair<K>(a));
  return d;
}

// position of entry (i, k) of pair (ca, cb), ca <= cb (i <= k when ca == cb), all compile-time in the callers
template <int D>
SKB_HD int stage_idx(int ca, int cb, int i, int k) {
  constexpr int K = D + 1;
  const int base = pair_base<D>(pair_of<K>(ca, cb));
  return base + ((ca == cb) ? (i * D - (i * (i - 1)) / 2 + (k - i)) : (i * D + k));
}

template <int D>
struct Sizes {
  static constexpr int K = D + 1;
  static constexpr int NL = K * D;                // local dofs (12 / 6)
  static constexpr int NK = NL * (NL + 1) / 2;    // packed local stiffness (78 / 21)
  static constexpr int NG = NL;                   // local gradient
  static constexpr int SMEM_DOUBLES = NK + NG;    // per element
};

// Raw inputs of one element, as loaded from HBM: corner positions (gathered through T), the element
// operator D, material and weight.  Kept separate from the math so that the pipelined kernel can issue
// these loads for its NEXT tile before the shared-memory phase of the current one.
template <int D>
struct ElemRaw {
  double xs[D + 1][D];
  double Dm[D][D];
  double mu, lam, vol;
};

template <int D>
SKB_HD void load_raw(const PlanView& p, const EvalArgs& a, int e, const int* Te, ElemRaw<D>& r) {
  constexpr int K = D + 1;
#pragma unroll
  for (int c = 0; c < K; ++c)
#pragma unroll
    for (int i = 0; i < D; ++i) r.xs[c][i] = a.x[(size_t)Te[c] * D + i];
#pragma unroll
  for (int j = 0; j < D; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) r.Dm[j][c] = p.Dm[(size_t)(j * D + c) * p.t + e];
  r.mu = a.mu[(size_t)e * a.mu_stride];
  r.lam = a.lam ? a.lam[(size_t)e * a.lam_stride] : 0.0;
  r.vol = a.vol[(size_t)e * a.vol_stride];
}

// F_ij = sum_{a>=1} D[j][a] (x_a[i] - x_0[i])  (+ Fbar)      -- the "J @ x" of the reference
template <int D>
SKB_HD void build_F(const EvalArgs& a, int e, const ElemRaw<D>& r, Mat<D>& F) {
  const size_t ef = (a.Fbar && a.eorder) ? (size_t)a.eorder[e] : (size_t)e;  // Fbar is listed in the caller's order
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) F.m[i][j] = a.Fbar ? a.Fbar[ef * D * D + i * D + j] : 0.0;
#pragma unroll
  for (int c = 0; c < D; ++c) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double d = r.xs[c + 1][i] - r.xs[0][i];
#pragma unroll
      for (int j = 0; j < D; ++j) F.m[i][j] = fma(r.Dm[j][c], d, F.m[i][j]);
    }
  }
}

template <int D>
SKB_HD void load_element(const PlanView& p, const EvalArgs& a, int e, Mat<D>& F, double Dm[D][D],
                         double& mu, double& lam, double& vol) {
  ElemRaw<D> r;
  load_raw<D>(p, a, e, p.T32 + (size_t)e * (D + 1), r);
  build_F<D>(a, e, r, F);
#pragma unroll
  for (int j = 0; j < D; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) Dm[j][c] = r.Dm[j][c];
  mu = r.mu;
  lam = r.lam;
  vol = r.vol;
}

template <int D>
SKB_HD double energy_element(const PlanView& p, const EvalArgs& a, int e) {
  Mat<D> F;
  double Dm[D][D], mu, lam, vol;
  load_element<D>(p, a, e, F, Dm, mu, lam, vol);
  return vol * energy_density<D>(a.material, F, mu, lam);
}

// Per-element state carried from the register-only math (element_math) to the shared-memory
// staging (element_store).  Isotropic models: SVD frame U, principal-stretch Hessian h (weighted and
// projected), W[c][q] = sum_j D[j][c] V[j][q].  Linear elasticity: W holds d[c][j] and (cI, cT, cR) the
// coefficients of the constant block.
template <int D>
struct ElemState {
  Mat<D> U;
  Principal<D> h;
  double W[D + 1][D];
  double cI, cT, cR;
};

// Phase 1a: one thread per element, registers only: gather, F, SVD, stress, principal-stretch
// Hessian with quadrature weight and eigenvalue floor.  The local gradient (12 values) goes straight
// to its staging area sG [value][le]; everything the stiffness blocks need stays in `st`.
// MAT >= 0 fixes the material at compile time (dead constitutive branches vanish, fewer live
// registers); MAT = -1 reads it from the arguments.
struct NoHook {
  SKB_HD void operator()() const {}
};

// `before_staging` is called exactly once, by every thread, after the register-only part that never touches shared
// memory (F, SVD) and before the first store to the gradient staging area: the warp-specialised kernel waits there for
// the reducer warps to release the pair's staging memory.
//
// SCRATCH: the element operator D, the material and the weight of element `le` sit in a per-thread column of shared
// memory sS ([D*D + 3][E]: D row-major, mu, lam, vol; filled by cp.async one tile ahead) and are read twice -- for F
// and again after the hook -- instead of staying in 24 registers across the SVD, where ptxas would spill them right
// behind their loads and stall on the load latency (profiles/r02p: STL behind LDG = 7 % of the warp time).
template <int D>
SKB_HD void scratch_read(const double* sS, int le, int E, ElemRaw<D>& r) {
#pragma unroll
  for (int j = 0; j < D; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) r.Dm[j][c] = sS[(j * D + c) * E + le];
  r.mu = sS[(D * D) * E + le];
  r.lam = sS[(D * D + 1) * E + le];
  r.vol = sS[(D * D + 2) * E + le];
}

template <int D, int MAT = -1, class Hook = NoHook, bool SCRATCH = false>
SKB_HD void element_math(const EvalArgs& a, int e, ElemRaw<D>& raw, int le, int E, double* sG, ElemState<D>& st,
                         Hook before_staging = Hook(), const double* sS = nullptr) {
  constexpr int K = D + 1;
  const int material = (MAT >= 0) ? MAT : a.material;
  Mat<D> F;
  if (SCRATCH) scratch_read<D>(sS, le, E, raw);
  build_F<D>(a, e, raw, F);

  Mat<D> V;
  Vec<D> sig;
  const bool iso = (material != MAT_LINEAR_ELASTICITY);
  const bool need_svd = (a.want_hess && iso) || (a.want_grad && material_uses_rotation(material));
  if (need_svd) svd_rv(F, st.U, sig, V);
  before_staging();
  if (SCRATCH) scratch_read<D>(sS, le, E, raw);  // the hook is a barrier with a memory clobber: real loads again
  const double mu = raw.mu, lam = raw.lam, vol = raw.vol;
  const double (&Dm)[D][D] = raw.Dm;

  if (a.want_grad) {
    Mat<D> P;
    if (material == MAT_ARAP) {
      Mat<D> R = matmul_nt(st.U, V);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = mu * (F.m[i][j] - R.m[i][j]);
    } else if (material == MAT_FCR) {
      // 2 mu (F - R) + lam (J - 1) cof F, with R from the SVD already taken (fcr.py:65-125)
      Mat<D> R = matmul_nt(st.U, V);
      Mat<D> c = cofactor(F);
      const double k = lam * (det(F) - 1.0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = fma(2.0 * mu, F.m[i][j] - R.m[i][j], k * c.m[i][j]);
    } else {
      P = pk1<D>(material, F, mu, lam);
    }
    // g_a[i] = vol * sum_j P[i][j] D[j][a];  corner 0 is minus the sum
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s0 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(P.m[i][j], Dm[j][c], s);
        s *= vol;
        sG[((c + 1) * D + i) * E + le] = s;
        s0 -= s;
      }
      sG[i * E + le] = s0;
    }
  }
  if (!a.want_hess) return;

  if (iso) {
    st.h = principal_hessian<D>(material, sig, mu, lam);
    weight_and_project<D>(st.h, vol, a.psd_mode);
    // W[c][q] = sum_j D[j][c] V[j][q]   (c = corner, corner 0 = minus the sum)
#pragma unroll
    for (int q = 0; q < D; ++q) {
      double s0 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(Dm[j][c], V.m[j][q], s);
        st.W[c + 1][q] = s;
        s0 -= s;
      }
      st.W[0][q] = s0;
    }
  } else {
    // linear elasticity: constant Hessian mu (I + T) + lam tr^T tr; psd flag ignored in its
    // own module (linear_elasticity.py:199-230) but honoured through the dispatcher, where
    // the floor only lifts the exact zero modes.  K_ab[i][k] = vol (mu (d_ik da.db + da[k] db[i]) + lam da[i] db[k])
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double s0 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        st.W[c + 1][j] = Dm[j][c];
        s0 -= Dm[j][c];
      }
      st.W[0][j] = s0;
    }
    // dispatcher PSD on the constant block: eigenvalues of mu(I+T)+lam tr^T tr are
    // 2mu (sym traceless), 2mu + D*lam (trace mode), 0 (skew modes).
    double w_sym = 2.0 * mu, w_tr = 2.0 * mu + D * lam, w_skew = 0.0;
    const double pre = (a.psd_mode == PSD_BEFORE_VOL) ? 1.0 : vol;
    const double post = (a.psd_mode == PSD_BEFORE_VOL) ? vol : 1.0;
    w_sym *= pre; w_tr *= pre; w_skew *= pre;
    if (a.psd_mode != PSD_NONE) {
      w_sym = psd_clamp(w_sym, a.psd_mode);
      w_tr = psd_clamp(w_tr, a.psd_mode);
      w_skew = psd_clamp(w_skew, a.psd_mode);
    }
    w_sym *= post; w_tr *= post; w_skew *= post;
    // H = w_sym * Psym0 + w_tr * Ptr + w_skew * Pskew, with projectors
    //   Psym = (I+T)/2, Pskew = (I-T)/2, Ptr = tr^T tr / D, Psym0 = Psym - Ptr
    // => H = cI * I + cT * T + cR * tr^T tr
    st.cI = 0.5 * (w_sym + w_skew);
    st.cT = 0.5 * (w_sym - w_skew);
    st.cR = (w_tr - w_sym) / D;
  }
}

// One D x D block K_(ca,cb) of the local stiffness from the element state.  Isotropic models:
//   K = U M U^T,  M[p][p] = S_pp Wa_p Wb_p + sum_{q != p} a_pq Wa_q Wb_q,  M[p][r] = S_pr Wa_p Wb_r + b_pr Wa_r Wb_p
// linear elasticity:  K[i][k] = cI d_ik (Wa . Wb) + cT Wa[k] Wb[i] + cR Wa[i] Wb[k].
template <int D, bool LINEAR>
SKB_HD Mat<D> local_block(const ElemState<D>& st, int ca, int cb) {
  constexpr int NP = D * (D - 1) / 2;
  Mat<D> Kb;
  if (!LINEAR) {
    Mat<D> M;
#pragma unroll
    for (int pp = 0; pp < D; ++pp)
#pragma unroll
      for (int r = 0; r < D; ++r) M.m[pp][r] = st.h.S.m[pp][r] * st.W[ca][pp] * st.W[cb][r];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      int pp, q, r3;
      pair_index<D>(k, pp, q, r3);
      M.m[pp][pp] = fma(st.h.a[k] * st.W[ca][q], st.W[cb][q], M.m[pp][pp]);
      M.m[q][q] = fma(st.h.a[k] * st.W[ca][pp], st.W[cb][pp], M.m[q][q]);
      M.m[pp][q] = fma(st.h.b[k] * st.W[ca][q], st.W[cb][pp], M.m[pp][q]);
      M.m[q][pp] = fma(st.h.b[k] * st.W[ca][pp], st.W[cb][q], M.m[q][pp]);
    }
    Mat<D> UM = matmul(st.U, M);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int kk = 0; kk < D; ++kk) {
        if (ca == cb && kk < i) {
          Kb.m[i][kk] = Kb.m[kk][i];
          continue;
        }
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < D; ++r) s = fma(UM.m[i][r], st.U.m[kk][r], s);
        Kb.m[i][kk] = s;
      }
  } else {
    double dot = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) dot = fma(st.W[ca][j], st.W[cb][j], dot);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int kk = 0; kk < D; ++kk) {
        double v = st.cT * st.W[ca][kk] * st.W[cb][i] + st.cR * st.W[ca][i] * st.W[cb][kk];
        if (i == kk) v = fma(st.cI, dot, v);
        Kb.m[i][kk] = v;
      }
  }
  return Kb;
}

// Corner-0 pairs of the local stiffness from the translation invariance sum_a K_ab = 0 (W[0] = -sum_c W[c]): reads
// the element's own staged pairs among corners 1..D back and writes K_0b = -sum_{a>=1} K_ab and K_00 = -sum_b K_0b^T.
// One thread per element, conflict-free columns.
template <int D>
SKB_HD void element_sum0(int le, int E, double* sK) {
  constexpr int K = D + 1;
  {
      const volatile double* rK = sK;   // real shared-memory loads: forwarding the stores would keep 45 values live
      double k00[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int kk = 0; kk < D; ++kk) k00[i][kk] = 0.0;
#pragma unroll
      for (int cb = 1; cb < K; ++cb)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            double acc = 0.0;   // K_0b[i][kk] = -sum_{a >= 1} K_ab[i][kk],  K_ab = K_ba^T for a > b
#pragma unroll
            for (int ca = 1; ca < K; ++ca) {
              int idx;
              if (ca < cb) idx = stage_idx<D>(ca, cb, i, kk);
              else if (ca > cb) idx = stage_idx<D>(cb, ca, kk, i);
              else idx = (i <= kk) ? stage_idx<D>(ca, ca, i, kk) : stage_idx<D>(ca, ca, kk, i);
              acc -= rK[idx * E + le];
            }
            sK[stage_idx<D>(0, cb, i, kk) * E + le] = acc;
            k00[kk][i] -= acc;  // K_00 = -sum_b K_0b^T
          }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int kk = i; kk < D; ++kk) sK[stage_idx<D>(0, 0, i, kk) * E + le] = k00[i][kk];
    }
}

// Phase 1b: writes the element's packed (K*D)x(K*D) local stiffness and its local gradient to
// staging memory laid out [value][le] (stride E) so that a warp's stores hit consecutive banks.
template <int D, int MAT = -1>
SKB_HD void element_store(const EvalArgs& a, const ElemState<D>& st, int le, int E, double* sK) {
  constexpr int K = D + 1;
  constexpr int NL = K * D;
  constexpr int NP = D * (D - 1) / 2;
  const int material = (MAT >= 0) ? MAT : a.material;
  if (!a.want_hess) return;
#if defined(SKB_EXP_NOSTORE)  // timing experiment only (wrong results): one value instead of the 78
  sK[le] = st.U.m[0][0] * st.h.S.m[0][0] * st.W[0][0];
  return;
#endif
  if (material != MAT_LINEAR_ELASTICITY) {
    // block (ca, cb), ca <= cb:  K = U M U^T,
    //   M[p][p] = S_pp Wa_p Wb_p + sum_{q != p} a_pq Wa_q Wb_q
    //   M[p][r] = S_pr Wa_p Wb_r + b_pr Wa_r Wb_p
#if defined(SKB_EXP_SUM0)
    // A/B experiment: only the pairs among corners 1..D are computed; the pairs of corner 0 follow from the
    // translation invariance sum_a K_ab = 0 (W[0] = -sum_c W[c]) by reading the element's own staged values back
    constexpr int CA0 = 1;
#else
    constexpr int CA0 = 0;
#endif
#pragma unroll
    for (int ca = CA0; ca < K; ++ca)
#pragma unroll
      for (int cb = ca; cb < K; ++cb) {
        Mat<D> M;
#pragma unroll
        for (int pp = 0; pp < D; ++pp)
#pragma unroll
          for (int r = 0; r < D; ++r) M.m[pp][r] = st.h.S.m[pp][r] * st.W[ca][pp] * st.W[cb][r];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          int pp, q, r3;
          pair_index<D>(k, pp, q, r3);
          M.m[pp][pp] = fma(st.h.a[k] * st.W[ca][q], st.W[cb][q], M.m[pp][pp]);
          M.m[q][q] = fma(st.h.a[k] * st.W[ca][pp], st.W[cb][pp], M.m[q][q]);
          M.m[pp][q] = fma(st.h.b[k] * st.W[ca][q], st.W[cb][pp], M.m[pp][q]);
          M.m[q][pp] = fma(st.h.b[k] * st.W[ca][pp], st.W[cb][q], M.m[q][pp]);
        }
        Mat<D> UM = matmul(st.U, M);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            if (ca == cb && kk < i) continue;
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < D; ++r) s = fma(UM.m[i][r], st.U.m[kk][r], s);
            sK[stage_idx<D>(ca, cb, i, kk) * E + le] = s;
          }
      }
#if defined(SKB_EXP_SUM0)
    element_sum0<D>(le, E, sK);
#endif
  } else {
#pragma unroll
    for (int ca = 0; ca < K; ++ca)
#pragma unroll
      for (int cb = ca; cb < K; ++cb) {
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) dot = fma(st.W[ca][j], st.W[cb][j], dot);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            if (ca == cb && kk < i) continue;
            double v = st.cT * st.W[ca][kk] * st.W[cb][i] + st.cR * st.W[ca][i] * st.W[cb][kk];
            if (i == kk) v = fma(st.cI, dot, v);
            sK[stage_idx<D>(ca, cb, i, kk) * E + le] = v;
          }
      }
  }
}

// Phase 1 = 1a + 1b (single-tile kernel and the CPU replay)
template <int D>
SKB_HD void element_phase1(const PlanView& p, const EvalArgs& a, int e, int le, int E, double* sK, double* sG) {
  ElemRaw<D> raw;
  load_raw<D>(p, a, e, p.T32 + (size_t)e * (D + 1), raw);
  ElemState<D> st;
  element_math<D>(a, e, raw, le, E, sG, st);
  element_store<D>(a, st, le, E, sK);
}

// Phase 2 (blocks): work item = tile-slot entry of an UPPER block (row vertex <= col vertex).
// Sums the entry's contributions in their fixed order and writes one dim x dim partial record.
// Corners are sorted per element, so a contribution is always a local pair a <= b; its values sit at
// compile-time offsets from (pair base, local element) in the staging buffer, and all of a contribution's
// loads are issued before the first add.  A slot is either a vertex-vertex block (every contribution a
// diagonal pair, symmetric, 6 / 3 stored values) or an edge block (every contribution off-diagonal).
template <int D>
SKB_HD void block_phase2(const SchedEntry* ent, const uint16_t* src, int w, int E, const double* sK, double* pblocks) {
  const SchedEntry en = ent[w];
  if (en.q == 0xffffffffu) return;  // alignment padding
  constexpr int DD = D * D, DS = D * (D + 1) / 2;
  double acc[DD];
  int c = (int)(en.range & 0xffffu);
  const int c1 = (int)(en.range >> 16);
#if defined(SKB_EXP_SRCBASE)   // the schedule carries (diag flag | staging offset of the pair), see ScatterSrc
#define SKB_P2_IS_DIAG(sc) (((sc) >> 15) != 0)
#define SKB_P2_BASE(sc) ((int)(((sc) >> 8) & 0x7fu))
#else
#define SKB_P2_IS_DIAG(sc) pair_is_diag<D>((int)((sc) >> 8))
#define SKB_P2_BASE(sc) pair_base<D>((int)((sc) >> 8))
#endif
  if (SKB_P2_IS_DIAG((unsigned)src[c])) {
    double sa[DS];
#pragma unroll
    for (int k = 0; k < DS; ++k) sa[k] = 0.0;
#pragma unroll kP2Unroll
    for (; c < c1; ++c) {
      const unsigned sc = src[c];
      const double* base = sK + SKB_P2_BASE(sc) * E + (int)(sc & 0xffu);
      double v[DS];
#pragma unroll
      for (int k = 0; k < DS; ++k) v[k] = base[k * E];
#pragma unroll
      for (int k = 0; k < DS; ++k) sa[k] += v[k];
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int lo = i < k ? i : k, hi = i < k ? k : i;
        acc[i * D + k] = sa[lo * D - (lo * (lo - 1)) / 2 + (hi - lo)];
      }
  } else {
#pragma unroll
    for (int k = 0; k < DD; ++k) acc[k] = 0.0;
#pragma unroll kP2Unroll
    for (; c < c1; ++c) {
      const unsigned sc = src[c];
      const double* base = sK + SKB_P2_BASE(sc) * E + (int)(sc & 0xffu);
      double v[DD];
#pragma unroll
      for (int k = 0; k < DD; ++k) v[k] = base[k * E];
#pragma unroll
      for (int k = 0; k < DD; ++k) acc[k] += v[k];
    }
  }
#if defined(SKB_EXP_NOREC)  // timing experiment only: keep the reduction, drop the record stores
  {
    double t = 0.0;
    for (int k = 0; k < DD; ++k) t += acc[k];
    if (t != 123.456) return;
  }
#endif
  constexpr int RS = RecStride<D>::value;
  double* rec = pblocks + (size_t)en.q * RS;
#pragma unroll
  for (int k = 0; k < RS; k += 4) {
    const double v0 = (k < DD) ? acc[k] : 0.0, v1 = (k + 1 < DD) ? acc[k + 1] : 0.0;
    const double v2 = (k + 2 < DD) ? acc[k + 2] : 0.0, v3 = (k + 3 < DD) ? acc[k + 3] : 0.0;
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(rec + k), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
#else
    rec[k] = v0;
    rec[k + 1] = v1;
    rec[k + 2] = v2;
    rec[k + 3] = v3;
#endif
  }
}

// Phase 2 (vertices): work item = tile-vertex entry.
template <int D>
SKB_HD void vert_phase2(const SchedEntry* ent, const uint16_t* src, int w, int E, const double* sG, double* pverts) {
  const SchedEntry en = ent[w];
  if (en.q == 0xffffffffu) return;
  double acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.0;
  const int c1 = (int)(en.range >> 16);
#pragma unroll kP2Unroll
  for (int c = (int)(en.range & 0xffffu); c < c1; ++c) {
    const unsigned sc = src[c];
    const int le = (int)(sc & 0xffu);
    const int ca = (int)(sc >> 8);
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] += sG[(ca * D + i) * E + le];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) pverts[(size_t)en.q * D + i] = acc[i];
}

// Level 2 (blocks): item = (upper slot u, entry j = i*D+k of the block): sum entry j over the
// slot's partial records in tile order, write it into the canonical scalar-CSR layout at block
// (v, w) and, transposed, at block (w, v).  D*D consecutive items read one 72-byte (32-byte in 2D)
// record together, so the record stream is read fully coalesced.
template <int D>
SKB_HD void block_finalize(const PlanView& p, int item, const double* pblocks, double* vals) {
  const int u = item / (D * D);
  const int j = item - u * (D * D);
  constexpr int RS = RecStride<D>::value;
  double acc = 0.0;
  const int q0 = p.blocks.sp_ptr[u];
  const int q1 = p.blocks.sp_ptr[u + 1];
  const UpperPos up = p.upos[u];
  // memory-level parallelism: the first four records of the slot are requested together (slots have
  // 2.3 records on average); the summation order stays q0, q0+1, ...
  const double* r0 = pblocks + (size_t)q0 * RS + j;
  const int nq = q1 - q0;
#if defined(SKB_FIN_CS) && defined(__CUDA_ARCH__)  // A/B: streaming (evict-first) accesses, everything is touched once
#define SKB_FIN_LD(p) __ldcs(p)
#define SKB_FIN_ST(p, v) __stcs((p), (v))
#else
#define SKB_FIN_LD(p) (*(p))
#define SKB_FIN_ST(p, v) (*(p) = (v))
#endif
  const double v0 = (nq > 0) ? SKB_FIN_LD(r0) : 0.0;
  const double v1 = (nq > 1) ? SKB_FIN_LD(r0 + RS) : 0.0;
  const double v2 = (nq > 2) ? SKB_FIN_LD(r0 + 2 * RS) : 0.0;
  const double v3 = (nq > 3) ? SKB_FIN_LD(r0 + 3 * RS) : 0.0;
  acc = v0;
  if (nq > 1) acc += v1;
  if (nq > 2) acc += v2;
  if (nq > 3) acc += v3;
  for (int q = q0 + 4; q < q1; ++q) acc += SKB_FIN_LD(pblocks + (size_t)q * RS + j);
  const int i = j / D, k = j - i * D;
  SKB_FIN_ST(vals + (size_t)up.base + (size_t)i * up.stride + k, acc);
  if (up.tbase != up.base) SKB_FIN_ST(vals + (size_t)up.tbase + (size_t)k * up.tstride + i, acc);
}

// A/B experiment (-DSKB_FIN_ITEMS=N): N items per thread, `stride` apart, with the loads of every stage requested
// for all N items before the first dependent use: level 2 is a chain of dependent round trips (slot pointers and
// position -> records -> stores) and one item per thread leaves the memory system under-subscribed at 40
// registers.  Same summation order per item as block_finalize => bit-identical values.
template <int D, int N>
SKB_HD void block_finalize_multi(const PlanView& p, int item0, int stride, int n_items, const double* pblocks, double* vals) {
  constexpr int RS = RecStride<D>::value;
  int j[N], q0[N], nq[N];
  UpperPos up[N];
  bool live[N];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const int item = item0 + t * stride;
    live[t] = item < n_items;
    const int u = live[t] ? item / (D * D) : 0;
    j[t] = live[t] ? item - u * (D * D) : 0;
    q0[t] = p.blocks.sp_ptr[u];
    nq[t] = live[t] ? p.blocks.sp_ptr[u + 1] - q0[t] : 0;
    up[t] = p.upos[u];
  }
  double v[N][4];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const double* r0 = pblocks + (size_t)q0[t] * RS + j[t];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[t][k] = (nq[t] > k) ? r0[k * RS] : 0.0;
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    if (!live[t]) continue;
    double acc = v[t][0];
    if (nq[t] > 1) acc += v[t][1];
    if (nq[t] > 2) acc += v[t][2];
    if (nq[t] > 3) acc += v[t][3];
    for (int q = q0[t] + 4; q < q0[t] + nq[t]; ++q) acc += pblocks[(size_t)q * RS + j[t]];
    const int i = j[t] / D, k = j[t] - i * D;
    vals[(size_t)up[t].base + (size_t)i * up[t].stride + k] = acc;
    if (up[t].tbase != up[t].base) vals[(size_t)up[t].tbase + (size_t)k * up[t].tstride + i] = acc;
  }
}

template <int D>
SKB_HD void vert_finalize(const PlanView& p, int v, const double* pverts, double* g) {
  double acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.0;
  const int q1 = p.verts.sp_ptr[v + 1];
  for (int q = p.verts.sp_ptr[v]; q < q1; ++q) {
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] += pverts[(size_t)q * D + i];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) g[(size_t)v * D + i] = acc[i];
}

// Index of the scalar entry (row, col) in the canonical CSR value array, or -1 when the block
// (row / D, col / D) is not in the pattern: the D rows of block row v are contiguous runs of
// D * (bptr[v+1] - bptr[v]) values, block columns sorted ascending (binary search).
template <int D>
SKB_HD int csr_value_position(const int* bptr, const int* bcol, int row, int col) {
  const int v = row / D, i = row - v * D;
  const int w = col / D, j = col - w * D;
  const int b0 = bptr[v], b1 = bptr[v + 1];
  int lo = b0, hi = b1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bcol[mid] < w) lo = mid + 1; else hi = mid;
  }
  if (lo >= b1 || bcol[lo] != w) return -1;
  return b0 * (D * D) + i * (D * (b1 - b0)) + D * (lo - b0) + j;
}

// shared-memory footprint of one assembly CTA (bytes); every region starts 16-byte aligned
template <int D>
inline size_t assemble_smem_bytes(const PlanView& p) {
  constexpr int K = D + 1;
  constexpr int NP = K * (K + 1) / 2;
  const size_t E = p.tile_elems;
  size_t b = (size_t)Sizes<D>::SMEM_DOUBLES * E * sizeof(double);
  b += (size_t)p.blocks.max_entries * sizeof(SchedEntry) + E * NP * sizeof(uint16_t);
  b += (size_t)p.verts.max_entries * sizeof(SchedEntry) + E * K * sizeof(uint16_t);
  return b + 16;  // + mbarrier
}

#if defined(__CUDACC__)
// ------------------------------------------------------------ TMA helpers ---
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}

// ------------------------------------------------------------ __global__ ---
// One CTA per tile of E elements, one thread per element in phase 1.  The tile's reduction
// schedule (entries + packed sources, four contiguous ranges) is fetched by the TMA unit into
// shared memory while phase 1 computes; phase 2 then never touches global memory for indices.
template <int D>
__global__ void assemble_tile_kernel(PlanView p, EvalArgs a) {
  constexpr int K = D + 1;
  constexpr int NP = K * (K + 1) / 2;
  extern __shared__ __align__(16) double smem[];
  const int E = blockDim.x;
  double* sK = smem;
  double* sG = smem + (size_t)Sizes<D>::NK * E;
  unsigned char* sp = reinterpret_cast<unsigned char*>(smem + (size_t)Sizes<D>::SMEM_DOUBLES * E);
  SchedEntry* sBE = reinterpret_cast<SchedEntry*>(sp);
  sp += (size_t)p.blocks.max_entries * sizeof(SchedEntry);
  uint16_t* sBS = reinterpret_cast<uint16_t*>(sp);
  sp += (size_t)E * NP * sizeof(uint16_t);
  SchedEntry* sVE = reinterpret_cast<SchedEntry*>(sp);
  sp += (size_t)p.verts.max_entries * sizeof(SchedEntry);
  uint16_t* sVS = reinterpret_cast<uint16_t*>(sp);
  sp += (size_t)E * K * sizeof(uint16_t);
  const unsigned mbar = smem_u32(sp);

  const int tile = blockIdx.x;
  const int le = threadIdx.x;
  const int e = tile * E + le;
  const int nbe = a.want_hess ? p.blocks.tl_ptr[tile + 1] - p.blocks.tl_ptr[tile] : 0;
  const int nve = a.want_grad ? p.verts.tl_ptr[tile + 1] - p.verts.tl_ptr[tile] : 0;
  if (threadIdx.x == 0) mbar_init(mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned bytes = 0;
    if (nbe) bytes += nbe * (unsigned)sizeof(SchedEntry) + E * NP * (unsigned)sizeof(uint16_t);
    if (nve) bytes += nve * (unsigned)sizeof(SchedEntry) + E * K * (unsigned)sizeof(uint16_t);
    mbar_expect_tx(mbar, bytes);
    if (nbe) {
      tma_bulk_g2s(smem_u32(sBE), p.blocks.tl_ent + p.blocks.tl_ptr[tile], nbe * (unsigned)sizeof(SchedEntry), mbar);
      tma_bulk_g2s(smem_u32(sBS), p.blocks.tc_src + (size_t)tile * E * NP, E * NP * (unsigned)sizeof(uint16_t), mbar);
    }
    if (nve) {
      tma_bulk_g2s(smem_u32(sVE), p.verts.tl_ent + p.verts.tl_ptr[tile], nve * (unsigned)sizeof(SchedEntry), mbar);
      tma_bulk_g2s(smem_u32(sVS), p.verts.tc_src + (size_t)tile * E * K, E * K * (unsigned)sizeof(uint16_t), mbar);
    }
  }
  if (e < p.t) element_phase1<D>(p, a, e, le, E, sK, sG);
  mbar_wait(mbar, 0);
  __syncthreads();
  for (int w = threadIdx.x; w < nbe; w += blockDim.x) block_phase2<D>(sBE, sBS, w, E, sK, a.pblocks);
  for (int w = threadIdx.x; w < nve; w += blockDim.x) vert_phase2<D>(sVE, sVS, w, E, sG, a.pverts);
}

// ------------------------------------------------------------ pipelined --
// Persistent variant: one CTA per SM, G groups of E threads, NBUF < G staging buffers.  The
// register-only math of phase 1a needs no shared memory, so all G groups (G*E/32 warps) keep the FP64
// pipe busy while only NBUF tiles' worth of staging exists: a group takes a buffer just for
// phase 1b + phase 2 and hands it back.  Groups synchronise with named barriers (bar.sync id, E);
// buffers are handed over through shared-memory flags.  Results are identical to the single-tile
// kernel (same per-tile schedule, same summation order).
template <int D>
struct PipeSmem {
  SKB_HD static size_t sched_bytes(const PlanView& p) {
    constexpr int K = D + 1;
    constexpr int NP = K * (K + 1) / 2;
    const size_t E = p.tile_elems;
    size_t b = (size_t)p.blocks.max_entries * sizeof(SchedEntry) + E * NP * sizeof(uint16_t);
    b += (size_t)p.verts.max_entries * sizeof(SchedEntry) + E * K * sizeof(uint16_t);
    return (b + 15) & ~(size_t)15;
  }
  // a staging buffer holds the packed stiffness only; every group keeps its own gradient staging
  SKB_HD static size_t buffer_bytes(const PlanView& p) { return (size_t)Sizes<D>::NK * p.tile_elems * sizeof(double); }
  SKB_HD static size_t grad_bytes(const PlanView& p) { return (size_t)Sizes<D>::NG * p.tile_elems * sizeof(double); }
  SKB_HD static size_t total(const PlanView& p, int G, int NBUF) {
    return NBUF * buffer_bytes(p) + G * (grad_bytes(p) + sched_bytes(p)) + 8 * (size_t)G + 4 * (size_t)(NBUF + 2 * G) + 32;
  }
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#if defined(SKB_PIPE_MAXNREG)  // A/B experiments: explicit register cap instead of the launch-bounds one
#define SKB_PIPE_BOUNDS(G, E) __maxnreg__(SKB_PIPE_MAXNREG)
#else
#define SKB_PIPE_BOUNDS(G, E) __launch_bounds__((G) * (E), 1)
#endif
template <int D, int G, int NBUF, int MAT, int E = 128>
__global__ void SKB_PIPE_BOUNDS(G, E) assemble_pipelined_kernel(PlanView p, EvalArgs a) {
  constexpr int K = D + 1;
  constexpr int NP = K * (K + 1) / 2;
  extern __shared__ __align__(16) double smem[];
  const int grp = threadIdx.x / E;
  const int gt = threadIdx.x - grp * E;
  unsigned char* base = reinterpret_cast<unsigned char*>(smem);
  const size_t bufB = PipeSmem<D>::buffer_bytes(p), schB = PipeSmem<D>::sched_bytes(p) + PipeSmem<D>::grad_bytes(p);
  unsigned char* sp = base + NBUF * bufB + (size_t)grp * schB;
  double* sG = reinterpret_cast<double*>(sp);
  sp += PipeSmem<D>::grad_bytes(p);
  SchedEntry* sBE = reinterpret_cast<SchedEntry*>(sp);
  sp += (size_t)p.blocks.max_entries * sizeof(SchedEntry);
  uint16_t* sBS = reinterpret_cast<uint16_t*>(sp);
  sp += (size_t)E * NP * sizeof(uint16_t);
  SchedEntry* sVE = reinterpret_cast<SchedEntry*>(sp);
  sp += (size_t)p.verts.max_entries * sizeof(SchedEntry);
  uint16_t* sVS = reinterpret_cast<uint16_t*>(sp);
  unsigned char* tail = base + NBUF * bufB + G * schB;
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(tail);  // G mbarriers
  int* sBusy = reinterpret_cast<int*>(sBar + G);     // NBUF flags, then G buffer indices
  int* sPick = sBusy + NBUF;
  int* sNext = sPick + G;                            // per group: next phase-2 chunk
  const unsigned mbar = smem_u32(sBar + grp);

  if (threadIdx.x < NBUF) sBusy[threadIdx.x] = 0;
  if (gt == 0) mbar_init(mbar, 1);
  __syncthreads();

  unsigned parity = 0;
  const int tstep = gridDim.x * G;
  int tile = blockIdx.x * G + grp;
  // software pipeline over tiles: the corner indices of the next tile are fetched at the top of an
  // iteration and its raw inputs are issued before the shared-memory phase of the current tile, so
  // the gather latency is hidden behind phase 2 instead of stalling all four warps of the group
  int Tn[K];
  ElemRaw<D> raw;
  if (tile < p.n_tiles) {
    const int e0 = tile * E + gt;
    if (e0 < p.t) {
#pragma unroll
      for (int c = 0; c < K; ++c) Tn[c] = p.T32[(size_t)e0 * K + c];
      load_raw<D>(p, a, e0, Tn, raw);
    }
  }
  for (; tile < p.n_tiles; tile += tstep) {
    const int e = tile * E + gt;
    const int en = e + tstep * E;  // this thread's element in the group's next tile
    const bool have_next = (tile + tstep < p.n_tiles) && (en < p.t);
    const int nbe = a.want_hess ? p.blocks.tl_ptr[tile + 1] - p.blocks.tl_ptr[tile] : 0;
    const int nve = a.want_grad ? p.verts.tl_ptr[tile + 1] - p.verts.tl_ptr[tile] : 0;
    if (gt == 0) {
      unsigned bytes = 0;
      if (nbe) bytes += nbe * (unsigned)sizeof(SchedEntry) + E * NP * (unsigned)sizeof(uint16_t);
      if (nve) bytes += nve * (unsigned)sizeof(SchedEntry) + E * K * (unsigned)sizeof(uint16_t);
      mbar_expect_tx(mbar, bytes);
      if (nbe) {
        tma_bulk_g2s(smem_u32(sBE), p.blocks.tl_ent + p.blocks.tl_ptr[tile], nbe * (unsigned)sizeof(SchedEntry), mbar);
        tma_bulk_g2s(smem_u32(sBS), p.blocks.tc_src + (size_t)tile * E * NP, E * NP * (unsigned)sizeof(uint16_t), mbar);
      }
      if (nve) {
        tma_bulk_g2s(smem_u32(sVE), p.verts.tl_ent + p.verts.tl_ptr[tile], nve * (unsigned)sizeof(SchedEntry), mbar);
        tma_bulk_g2s(smem_u32(sVS), p.verts.tc_src + (size_t)tile * E * K, E * K * (unsigned)sizeof(uint16_t), mbar);
      }
    }
    if (have_next) {
#pragma unroll
      for (int c = 0; c < K; ++c) Tn[c] = p.T32[(size_t)en * K + c];
    }
    ElemState<D> st;
    if (e < p.t) element_math<D, MAT>(a, e, raw, gt, E, sG, st);
    // take a staging buffer
#if defined(SKB_EXP_NOSYNC)
    if (gt == 0) sPick[grp] = grp % NBUF;
#else
    if (gt == 0) {
      int b = -1;
      while (b < 0) {
#pragma unroll
        for (int k = 0; k < NBUF; ++k)
          if (b < 0 && atomicCAS(&sBusy[k], 0, 1) == 0) b = k;
        if (b < 0) __nanosleep(64);
      }
      sPick[grp] = b;
    }
#endif
    if (gt == 0) sNext[grp] = 0;
    group_barrier(grp + 1, E);
    const int b = sPick[grp];
    double* sK = reinterpret_cast<double*>(base + (size_t)b * bufB);
    if (e < p.t) element_store<D, MAT>(a, st, gt, E, sK);
    mbar_wait(mbar, parity);
    parity ^= 1u;
    group_barrier(grp + 1, E);
    if (have_next) load_raw<D>(p, a, en, Tn, raw);  // in flight during phase 2
#if !defined(SKB_EXP_NOPHASE2)
#if defined(SKB_P2_STATIC)
    for (int w = gt; w < nbe; w += E) block_phase2<D>(sBE, sBS, w, E, sK, a.pblocks);
    for (int w = gt; w < nve; w += E) vert_phase2<D>(sVE, sVS, w, E, sG, a.pverts);
#else
    // The entries are sorted by descending contribution count, so chunk 0 (32 entries) takes several
    // times longer than the last one.  The group's warps therefore pull 32-entry chunks from a shared
    // counter (greedy longest-first scheduling) instead of striding through them; the vertex chunks are
    // slotted in after the first block chunks.  Which warp sums a slot does not change its value.
    {
      const int lane = gt & 31;
      const int nbc = (nbe + 31) >> 5, nvc = (nve + 31) >> 5;
      const int nb_first = nbc < 4 ? nbc : 4;
      for (;;) {
        int ch = 0;
        if (lane == 0) ch = atomicAdd(&sNext[grp], 1);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= nbc + nvc) break;
        if (ch >= nb_first && ch < nb_first + nvc) {
          const int w = ((ch - nb_first) << 5) + lane;
          if (w < nve) vert_phase2<D>(sVE, sVS, w, E, sG, a.pverts);
        } else {
          const int w = ((ch < nb_first ? ch : ch - nvc) << 5) + lane;
          if (w < nbe) block_phase2<D>(sBE, sBS, w, E, sK, a.pblocks);
        }
      }
    }
#endif
#endif
    group_barrier(grp + 1, E);
    if (gt == 0) {
      __threadfence_block();
      atomicExch(&sBusy[b], 0);
    }
  }
}

// ------------------------------------------------------------ warp-specialised --
// The pipelined kernel above is latency-bound with the register file full: its 12 warps hold 168 registers each
// through phase 2 (a shared-memory reduction that needs 50) and through the waits for a staging buffer, ptxas spills
// the next tile's prefetched inputs right behind their loads (a stall of a full DRAM latency per tile), and only ~43 %
// of the warp time is math (profiles/r02p_assemble_ncu_summary.txt).  Here the roles are split the Blackwell way:
// a CTA is four warpgroups, two COMPUTE groups that only ever run the per-element math (phase 1a + 1b) at 200
// registers (setmaxnreg.inc: no spills), and two REDUCER groups at 56 registers (setmaxnreg.dec) that run phase 2 and
// the TMA prefetch of the reduction schedule.  2 x 128 x 200 + 2 x 128 x 56 = 65,536 registers: the whole file.
//   * Each compute group owns one staging area (packed stiffness + gradient) and hands it to the reducers and back
//     with two named barriers (bar.arrive / bar.sync, the PTX producer-consumer idiom): FULL (compute arrives, reducers
//     wait) and EMPTY (reducers arrive, compute waits -- inside element_math, after the SVD and before its first
//     staging store, so the wait is normally over before it is reached).
//   * The reducers are POOLED: all 256 reducer threads work on one tile at a time, alternating between the two
//     compute groups' areas, which halves the time an area stays with the reducers (4.46 -> 4.21 ms at C5).
//   * Inputs of a compute group's NEXT tile never pass through long-lived registers: the corner indices, the element
//     operator, material and weight are copied global -> shared with cp.async (LDGSTS) one tile ahead; only the 12
//     gathered coordinates (the indirect loads) sit in registers, and only during the K-block math.  The operator is
//     read from that scratch twice (for F and after the SVD) instead of living across the SVD.
//   * The schedule of a group's next tile is requested (TMA bulk copies on an mbarrier) as soon as its last phase 2 ends.
// Same per-tile schedule and summation order as the other two kernels: results agree to the compiler's FMA contraction
// and are bit-identical run to run (tests/test_gpu_kernel_variants.py).  Measured at C5 (profiles/r02t_*): pipelined
// 4.79 ms, this kernel 3.79 ms; compute groups alone 3.26 ms, reducers alone 2.62 ms (before the sweep loops of the
// SVD / eigen-solve were re-rolled, which cut the instruction footprint from 99 KB to 52 KB and the kernel by 10 %).
template <int D>
struct WsSmem {
  static constexpr int P = 2;  // (compute, reducer) pairs per CTA
  // per compute group: [D*D + 3][E] doubles (element operator, mu, lam, vol) + [K][E] corner indices of the NEXT tile
  SKB_HD static size_t scratch_bytes(const PlanView& p) {
    return (size_t)p.tile_elems * ((D * D + 3) * sizeof(double) + (D + 1) * sizeof(int));
  }
  SKB_HD static size_t pair_bytes(const PlanView& p) {
    return PipeSmem<D>::buffer_bytes(p) + PipeSmem<D>::grad_bytes(p) + PipeSmem<D>::sched_bytes(p) + scratch_bytes(p);
  }
  SKB_HD static size_t total(const PlanView& p) { return P * pair_bytes(p) + 8 * (size_t)P + 4 * (size_t)P + 32; }
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// Ampere-style asynchronous copies global -> shared (SASS LDGSTS): no register is involved, so nothing can be spilled
// behind the load
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifndef SKB_WS_COMPUTE_REGS
#define SKB_WS_COMPUTE_REGS 200
#endif
#ifndef SKB_WS_REDUCER_REGS
#define SKB_WS_REDUCER_REGS 56
#endif
#ifndef SKB_WS_POOLED   // 1: both reducer groups work on every tile; 0 (A/B): reducer group c serves compute group c
#define SKB_WS_POOLED 1
#endif

template <int D, int MAT>
__global__ void __launch_bounds__(512, 1) assemble_ws_kernel(PlanView p, EvalArgs a) {
  constexpr int K = D + 1;
  constexpr int NP = K * (K + 1) / 2;
  constexpr int DD = D * D;
  constexpr int E = 128;  // a warpgroup = one tile
  constexpr int P = WsSmem<D>::P;
  constexpr int BAR_FULL = 1, BAR_EMPTY = 1 + P, BAR_RG = 1 + 2 * P;  // named barriers 1 .. 3P (0 = __syncthreads)
  constexpr bool POOL = SKB_WS_POOLED != 0;
  constexpr int RT = POOL ? P * E : E;  // reducer threads that work on one tile
  constexpr int HAND = E + RT;          // threads on a FULL / EMPTY barrier: one compute group + its reducers
  extern __shared__ __align__(16) double smem[];
  const int wg = threadIdx.x >> 7;
  const int gt = threadIdx.x & 127;
  const bool is_compute = wg < P;
  const int pair = is_compute ? wg : wg - P;
  unsigned char* base = reinterpret_cast<unsigned char*>(smem);
  const size_t bufB = PipeSmem<D>::buffer_bytes(p), gradB = PipeSmem<D>::grad_bytes(p), schB = PipeSmem<D>::sched_bytes(p);
  unsigned char* pb = base + (size_t)pair * WsSmem<D>::pair_bytes(p);
  double* sK = reinterpret_cast<double*>(pb);
  double* sG = reinterpret_cast<double*>(pb + bufB);
  unsigned char* sched = pb + bufB + gradB;
  double* sS = reinterpret_cast<double*>(pb + bufB + gradB + schB);  // [DD + 3][E]
  int* sT = reinterpret_cast<int*>(sS + (DD + 3) * E);                // [K][E]
  unsigned char* tail = base + (size_t)P * WsSmem<D>::pair_bytes(p);
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(tail);  // [P] mbarriers of the schedule copies
  int* sNext = reinterpret_cast<int*>(sBar + P);                           // [P] next phase-2 chunk

  if (threadIdx.x < P) {
    mbar_init(smem_u32(sBar + threadIdx.x), 1);
    sNext[threadIdx.x] = 0;
  }
  __syncthreads();

  const int tstep = gridDim.x * P;
  int tile = blockIdx.x * P + pair;

  if (is_compute) {
    reg_alloc<SKB_WS_COMPUTE_REGS>();
    // Threads past the end of the mesh run the math on the last element (no divergence around the barriers); their
    // staging column is not referenced by the tile's schedule.
    const int last = p.t - 1;
    auto prefetch_T = [&](int en) {  // corner indices -> sT
#pragma unroll
      for (int c = 0; c < K; ++c) cp_async4(sT + c * E + gt, p.T32 + (size_t)en * K + c);
      cp_async_commit();
    };
    auto prefetch_S = [&](int en) {  // element operator, material, weight -> sS
#pragma unroll
      for (int j = 0; j < DD; ++j) cp_async8(sS + j * E + gt, p.Dm + (size_t)j * p.t + en);
      cp_async8(sS + DD * E + gt, a.mu + (size_t)en * a.mu_stride);
      if (a.lam) cp_async8(sS + (DD + 1) * E + gt, a.lam + (size_t)en * a.lam_stride);
      else sS[(DD + 1) * E + gt] = 0.0;
      cp_async8(sS + (DD + 2) * E + gt, a.vol + (size_t)en * a.vol_stride);
      cp_async_commit();
    };
    ElemRaw<D> raw;
    auto gather_x = [&]() {  // the only indirect loads; 2 D (D + 1) registers in flight during the K-block math
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int v = sT[c * E + gt];
#pragma unroll
        for (int i = 0; i < D; ++i) raw.xs[c][i] = a.x[(size_t)v * D + i];
      }
    };
    if (tile < p.n_tiles) {
      const int e0 = min(tile * E + gt, last);
      prefetch_T(e0);
      prefetch_S(e0);
      cp_async_wait<0>();
      gather_x();
    }
    for (; tile < p.n_tiles; tile += tstep) {
      const int e = min(tile * E + gt, last);
      const bool have_next = tile + tstep < p.n_tiles;
      const int en = min((tile + tstep) * E + gt, last);
      if (have_next) {
        prefetch_T(en);
        cp_async_wait<1>();  // everything but the copy group just committed: this tile's sS
      } else {
        cp_async_wait<0>();
      }
      ElemState<D> st;
      auto wait_empty = [&]() { named_bar_sync(BAR_EMPTY + pair, HAND); };
#if defined(SKB_WS_NOMATH)  // timing experiment only (wrong results): the reducers alone
      wait_empty();
#else
      element_math<D, MAT, decltype(wait_empty), true>(a, e, raw, gt, E, sG, st, wait_empty, sS);
#endif
      if (have_next) {
        cp_async_wait<0>();  // the next tile's corner indices (requested before the math)
        gather_x();          // in flight during the K-block math
        prefetch_S(en);      // this tile's sS was last read right after the hook
      }
#if !defined(SKB_WS_NOMATH)
      element_store<D, MAT>(a, st, gt, E, sK);
#endif
      __threadfence_block();
      named_bar_arrive(BAR_FULL + pair, HAND);
    }
  } else {
    reg_dealloc<SKB_WS_REDUCER_REGS>();
    // Pooled: all P * E reducer threads walk the CTA's tiles in order (pair 0, pair 1, pair 0, ...); otherwise reducer
    // group c walks the tiles of compute group c.
    const int rt = POOL ? (int)threadIdx.x - P * E : gt;
    const int rg = POOL ? 0 : pair;  // reducer-group barrier / chunk counter
    unsigned parbits = 0;
    auto issue = [&](int tl, int pr) {  // one thread: the tile's reduction schedule -> pair pr's schedule area (TMA unit)
      unsigned char* sc = base + (size_t)pr * WsSmem<D>::pair_bytes(p) + bufB + gradB;
      const unsigned mb = smem_u32(sBar + pr);
      const int nbe = a.want_hess ? p.blocks.tl_ptr[tl + 1] - p.blocks.tl_ptr[tl] : 0;
      const int nve = a.want_grad ? p.verts.tl_ptr[tl + 1] - p.verts.tl_ptr[tl] : 0;
      unsigned bytes = 0;
      if (nbe) bytes += nbe * (unsigned)sizeof(SchedEntry) + E * NP * (unsigned)sizeof(uint16_t);
      if (nve) bytes += nve * (unsigned)sizeof(SchedEntry) + E * K * (unsigned)sizeof(uint16_t);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy reads of the last tile are done
      mbar_expect_tx(mb, bytes);
      unsigned char* q = sc;
      if (nbe) tma_bulk_g2s(smem_u32(q), p.blocks.tl_ent + p.blocks.tl_ptr[tl], nbe * (unsigned)sizeof(SchedEntry), mb);
      q += (size_t)p.blocks.max_entries * sizeof(SchedEntry);
      if (nbe) tma_bulk_g2s(smem_u32(q), p.blocks.tc_src + (size_t)tl * E * NP, E * NP * (unsigned)sizeof(uint16_t), mb);
      q += (size_t)E * NP * sizeof(uint16_t);
      if (nve) tma_bulk_g2s(smem_u32(q), p.verts.tl_ent + p.verts.tl_ptr[tl], nve * (unsigned)sizeof(SchedEntry), mb);
      q += (size_t)p.verts.max_entries * sizeof(SchedEntry);
      if (nve) tma_bulk_g2s(smem_u32(q), p.verts.tc_src + (size_t)tl * E * K, E * K * (unsigned)sizeof(uint16_t), mb);
    };
    // every staging area starts empty; its first schedule is requested right away
#pragma unroll
    for (int pr = 0; pr < P; ++pr) {
      if (!POOL && pr != pair) continue;
      const int t0 = blockIdx.x * P + pr;
      if (t0 < p.n_tiles) {
        if (rt == 0) issue(t0, pr);
        named_bar_arrive(BAR_EMPTY + pr, HAND);
      }
    }
    for (int s = 0;; ++s) {
      const int pr = POOL ? s % P : pair;
      const int tl = blockIdx.x * P + pr + (POOL ? s / P : s) * tstep;  // increasing in s
      if (tl >= p.n_tiles) break;
      const bool have_next = tl + tstep < p.n_tiles;
      const int nbe = a.want_hess ? p.blocks.tl_ptr[tl + 1] - p.blocks.tl_ptr[tl] : 0;
      const int nve = a.want_grad ? p.verts.tl_ptr[tl + 1] - p.verts.tl_ptr[tl] : 0;
      unsigned char* ppb = base + (size_t)pr * WsSmem<D>::pair_bytes(p);
      const double* rK = reinterpret_cast<const double*>(ppb);
      const double* rG = reinterpret_cast<const double*>(ppb + bufB);
      const unsigned char* sc = ppb + bufB + gradB;
      const SchedEntry* sBE = reinterpret_cast<const SchedEntry*>(sc);
      sc += (size_t)p.blocks.max_entries * sizeof(SchedEntry);
      const uint16_t* sBS = reinterpret_cast<const uint16_t*>(sc);
      sc += (size_t)E * NP * sizeof(uint16_t);
      const SchedEntry* sVE = reinterpret_cast<const SchedEntry*>(sc);
      sc += (size_t)p.verts.max_entries * sizeof(SchedEntry);
      const uint16_t* sVS = reinterpret_cast<const uint16_t*>(sc);
      named_bar_sync(BAR_FULL + pr, HAND);                         // compute group pr has staged the tile
      mbar_wait(smem_u32(sBar + pr), (parbits >> pr) & 1u);       // its schedule has landed (requested a tile ago)
      parbits ^= 1u << pr;
#if !defined(SKB_WS_NOP2)  // (timing experiment only when defined: the compute groups alone)
      {
        // greedy longest-first chunks of 32 entries from a shared counter, as in the pipelined kernel
        const int lane = rt & 31;
        const int nbc = (nbe + 31) >> 5, nvc = (nve + 31) >> 5;
        const int nb_first = nbc < 4 ? nbc : 4;
        for (;;) {
          int ch = 0;
          if (lane == 0) ch = atomicAdd(&sNext[rg], 1);
          ch = __shfl_sync(0xffffffffu, ch, 0);
          if (ch >= nbc + nvc) break;
          if (ch >= nb_first && ch < nb_first + nvc) {
            const int w = ((ch - nb_first) << 5) + lane;
            if (w < nve) vert_phase2<D>(sVE, sVS, w, E, rG, a.pverts);
          } else {
            const int w = ((ch < nb_first ? ch : ch - nvc) << 5) + lane;
            if (w < nbe) block_phase2<D>(sBE, sBS, w, E, rK, a.pblocks);
          }
        }
      }
#endif
      named_bar_sync(BAR_RG + rg, RT);  // every reducer thread is done with the staging area and the schedule
      if (rt == 0) {
        sNext[rg] = 0;
        if (have_next) issue(tl + tstep, pr);
      }
      if (have_next) named_bar_arrive(BAR_EMPTY + pr, HAND);
    }
  }
}

#ifndef SKB_FIN_THREADS
#define SKB_FIN_THREADS 128
#endif
#if defined(SKB_FIN_ITEMS)
#define SKB_FIN_PER_THREAD SKB_FIN_ITEMS
#else
#define SKB_FIN_PER_THREAD 1
#endif
template <int D>
__global__ void finalize_blocks_kernel(PlanView p, const double* pblocks, double* vals) {
#if defined(SKB_FIN_ITEMS)
  block_finalize_multi<D, SKB_FIN_ITEMS>(p, blockIdx.x * (blockDim.x * SKB_FIN_ITEMS) + threadIdx.x, blockDim.x,
                                         p.nu * (D * D), pblocks, vals);
#else
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item < p.nu * (D * D)) block_finalize<D>(p, item, pblocks, vals);
#endif
}

template <int D>
__global__ void finalize_verts_kernel(PlanView p, const double* pverts, double* g) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < p.n) vert_finalize<D>(p, v, pverts, g);
}

// fixed-shape block reduction: deterministic for a fixed launch configuration
__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;  // valid in thread 0
}

// count <= p.t: the elements that take part in the sum (a shard that also evaluates its lower neighbour's interface
// elements counts only its own, which the plan lists first)
template <int D>
__global__ void energy_kernel(PlanView p, EvalArgs a, int count, double* block_sums) {
  __shared__ double sh[32];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double v = (e < count) ? energy_element<D>(p, a, e) : 0.0;
  v = block_reduce_sum(v, sh);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
}

// single-block fixed-order final reduction of `n` partial sums; out[0] = total
static __global__ void reduce_final_kernel(const double* in, int n, double* out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += in[i];
  v = block_reduce_sum(v, sh);
  if (threadIdx.x == 0) out[0] = v;
}
#endif  // __CUDACC__

}  // namespace skb
