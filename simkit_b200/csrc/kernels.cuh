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
  const double* Fbar;  // (t*dim*dim) per-element offset of the _u tier, or nullptr
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

// packed upper-triangular index of an N x N symmetric matrix
SKB_HD int sym_idx(int N, int r, int c) {
  if (r > c) {
    int tmp = r;
    r = c;
    c = tmp;
  }
  return r * N - (r * (r - 1)) / 2 + (c - r);
}

template <int D>
struct Sizes {
  static constexpr int K = D + 1;
  static constexpr int NL = K * D;                // local dofs (12 / 6)
  static constexpr int NK = NL * (NL + 1) / 2;    // packed local stiffness (78 / 21)
  static constexpr int NG = NL;                   // local gradient
  static constexpr int SMEM_DOUBLES = NK + NG;    // per element
};

// F_ij = sum_{a>=1} D[j][a] (x_a[i] - x_0[i])  (+ Fbar)      -- the "J @ x" of the reference
template <int D>
SKB_HD void load_element(const PlanView& p, const EvalArgs& a, int e, Mat<D>& F, double Dm[D][D],
                         double& mu, double& lam, double& vol) {
  constexpr int K = D + 1;
  const int* Te = p.T32 + (size_t)e * K;
  double x0[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x0[i] = a.x[(size_t)Te[0] * D + i];
#pragma unroll
  for (int j = 0; j < D; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) Dm[j][c] = p.Dm[(size_t)(j * D + c) * p.t + e];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) F.m[i][j] = a.Fbar ? a.Fbar[(size_t)e * D * D + i * D + j] : 0.0;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const size_t v = (size_t)Te[c + 1] * D;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double d = a.x[v + i] - x0[i];
#pragma unroll
      for (int j = 0; j < D; ++j) F.m[i][j] = fma(Dm[j][c], d, F.m[i][j]);
    }
  }
  mu = a.mu[(size_t)e * a.mu_stride];
  lam = a.lam ? a.lam[(size_t)e * a.lam_stride] : 0.0;
  vol = a.vol[(size_t)e * a.vol_stride];
}

template <int D>
SKB_HD double energy_element(const PlanView& p, const EvalArgs& a, int e) {
  Mat<D> F;
  double Dm[D][D], mu, lam, vol;
  load_element<D>(p, a, e, F, Dm, mu, lam, vol);
  return vol * energy_density<D>(a.material, F, mu, lam);
}

// Phase 1: one thread per element.  Writes the element's packed (K*D)x(K*D)
// local stiffness and its local gradient to staging memory laid out [value][le]
// (stride E) so that a warp's stores hit consecutive banks.
template <int D>
SKB_HD void element_phase1(const PlanView& p, const EvalArgs& a, int e, int le, int E, double* sK, double* sG) {
  constexpr int K = D + 1;
  constexpr int NL = K * D;
  constexpr int NP = D * (D - 1) / 2;
  Mat<D> F;
  double Dm[D][D], mu, lam, vol;
  load_element<D>(p, a, e, F, Dm, mu, lam, vol);

  Mat<D> U, V;
  Vec<D> sig;
  const bool iso = (a.material != MAT_LINEAR_ELASTICITY);
  const bool need_svd = (a.want_hess && iso) || (a.want_grad && a.material == MAT_ARAP);
  if (need_svd) svd_rv(F, U, sig, V);

  if (a.want_grad) {
    Mat<D> P;
    if (a.material == MAT_ARAP) {
      Mat<D> R = matmul_nt(U, V);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P.m[i][j] = mu * (F.m[i][j] - R.m[i][j]);
    } else {
      P = pk1<D>(a.material, F, mu, lam);
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
    Principal<D> h = principal_hessian<D>(a.material, sig, mu, lam);
    weight_and_project<D>(h, vol, a.psd_mode);
    // W[c][q] = sum_j D[j][c] V[j][q]   (c = corner, corner 0 = minus the sum)
    double W[K][D];
#pragma unroll
    for (int q = 0; q < D; ++q) {
      double s0 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(Dm[j][c], V.m[j][q], s);
        W[c + 1][q] = s;
        s0 -= s;
      }
      W[0][q] = s0;
    }
    // block (ca, cb), ca <= cb:  K = U M U^T,
    //   M[p][p] = S_pp Wa_p Wb_p + sum_{q != p} a_pq Wa_q Wb_q
    //   M[p][r] = S_pr Wa_p Wb_r + b_pr Wa_r Wb_p
#pragma unroll
    for (int ca = 0; ca < K; ++ca)
#pragma unroll
      for (int cb = ca; cb < K; ++cb) {
        Mat<D> M;
#pragma unroll
        for (int pp = 0; pp < D; ++pp)
#pragma unroll
          for (int r = 0; r < D; ++r) M.m[pp][r] = h.S.m[pp][r] * W[ca][pp] * W[cb][r];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          int pp, q, r3;
          pair_index<D>(k, pp, q, r3);
          M.m[pp][pp] = fma(h.a[k] * W[ca][q], W[cb][q], M.m[pp][pp]);
          M.m[q][q] = fma(h.a[k] * W[ca][pp], W[cb][pp], M.m[q][q]);
          M.m[pp][q] = fma(h.b[k] * W[ca][q], W[cb][pp], M.m[pp][q]);
          M.m[q][pp] = fma(h.b[k] * W[ca][pp], W[cb][q], M.m[q][pp]);
        }
        Mat<D> UM = matmul(U, M);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            if (ca == cb && kk < i) continue;
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < D; ++r) s = fma(UM.m[i][r], U.m[kk][r], s);
            sK[sym_idx(NL, ca * D + i, cb * D + kk) * E + le] = s;
          }
      }
  } else {
    // linear elasticity: constant Hessian mu (I + T) + lam tr^T tr; psd flag ignored in its
    // own module (linear_elasticity.py:199-230) but honoured through the dispatcher, where
    // the floor only lifts the exact zero modes.  K_ab[i][k] = vol (mu (d_ik da.db + da[k] db[i]) + lam da[i] db[k])
    double d[K][D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double s0 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        d[c + 1][j] = Dm[j][c];
        s0 -= Dm[j][c];
      }
      d[0][j] = s0;
    }
    // dispatcher PSD on the constant block: eigenvalues of mu(I+T)+lam tr^T tr are
    // 2mu (sym traceless, and twist -> 0), 2mu + D*lam (trace mode), 0 (skew modes).
    // Floor / weight handling: skew (zero) modes become `fl0`.
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
    const double cI = 0.5 * (w_sym + w_skew), cT = 0.5 * (w_sym - w_skew), cR = (w_tr - w_sym) / D;
#pragma unroll
    for (int ca = 0; ca < K; ++ca)
#pragma unroll
      for (int cb = ca; cb < K; ++cb) {
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) dot = fma(d[ca][j], d[cb][j], dot);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            if (ca == cb && kk < i) continue;
            double v = cT * d[ca][kk] * d[cb][i] + cR * d[ca][i] * d[cb][kk];
            if (i == kk) v = fma(cI, dot, v);
            sK[sym_idx(NL, ca * D + i, cb * D + kk) * E + le] = v;
          }
      }
  }
}

// Phase 2 (blocks): work item = (tile-slot entry, block row i).  Sums the
// entry's contributions in their fixed order and writes one partial-record row.
template <int D>
SKB_HD void block_phase2(const ReduceSchedView& s, int tile, int w, int E, const double* sK, double* pblocks) {
  constexpr int K = D + 1;
  constexpr int NL = K * D;
  const int entry = s.tl_ptr[tile] + w / D;
  const int i = w - (w / D) * D;
  const int q = s.tl_q[entry];
  double acc[D];
#pragma unroll
  for (int k = 0; k < D; ++k) acc[k] = 0.0;
  const int c1 = s.tl_cptr[entry + 1];
  for (int c = s.tl_cptr[entry]; c < c1; ++c) {
    const int src = s.tc_src[c];
    const int le = src / (K * K);
    const int ab = src - le * (K * K);
    const int ca = ab / K, cb = ab - ca * K;
    const int r = ca * D + i;
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] += sK[sym_idx(NL, r, cb * D + k) * E + le];
  }
#pragma unroll
  for (int k = 0; k < D; ++k) pblocks[(size_t)q * (D * D) + i * D + k] = acc[k];
}

// Phase 2 (vertices): work item = tile-vertex entry.
template <int D>
SKB_HD void vert_phase2(const ReduceSchedView& s, int tile, int w, int E, const double* sG, double* pverts) {
  constexpr int K = D + 1;
  const int entry = s.tl_ptr[tile] + w;
  const int q = s.tl_q[entry];
  double acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.0;
  const int c1 = s.tl_cptr[entry + 1];
  for (int c = s.tl_cptr[entry]; c < c1; ++c) {
    const int src = s.tc_src[c];
    const int le = src / K;
    const int ca = src - le * K;
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] += sG[(ca * D + i) * E + le];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) pverts[(size_t)q * D + i] = acc[i];
}

// Level 2 (blocks): item = (slot, row i): sum the slot's partial records in
// tile order, write row i of the block into the canonical scalar-CSR layout.
template <int D>
SKB_HD void block_finalize(const PlanView& p, int item, const double* pblocks, double* vals) {
  const int s = item / D;
  const int i = item - s * D;
  double acc[D];
#pragma unroll
  for (int k = 0; k < D; ++k) acc[k] = 0.0;
  const int q1 = p.blocks.sp_ptr[s + 1];
  for (int q = p.blocks.sp_ptr[s]; q < q1; ++q) {
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] += pblocks[(size_t)q * (D * D) + i * D + k];
  }
  const int v = p.brow[s];
  const int b0 = p.bptr[v];
  const int nb = p.bptr[v + 1] - b0;
  const size_t pos = (size_t)b0 * (D * D) + (size_t)i * nb * D + (size_t)(s - b0) * D;
#pragma unroll
  for (int k = 0; k < D; ++k) vals[pos + k] = acc[k];
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

#if defined(__CUDACC__)
// ------------------------------------------------------------ __global__ ---
template <int D>
__global__ void assemble_tile_kernel(PlanView p, EvalArgs a) {
  extern __shared__ double smem[];
  const int E = blockDim.x;
  double* sK = smem;
  double* sG = smem + (size_t)Sizes<D>::NK * E;
  const int tile = blockIdx.x;
  const int le = threadIdx.x;
  const int e = tile * E + le;
  if (e < p.t) element_phase1<D>(p, a, e, le, E, sK, sG);
  __syncthreads();
  if (a.want_hess) {
    const int nitems = (p.blocks.tl_ptr[tile + 1] - p.blocks.tl_ptr[tile]) * D;
    for (int w = threadIdx.x; w < nitems; w += blockDim.x) block_phase2<D>(p.blocks, tile, w, E, sK, a.pblocks);
  }
  if (a.want_grad) {
    const int nitems = p.verts.tl_ptr[tile + 1] - p.verts.tl_ptr[tile];
    for (int w = threadIdx.x; w < nitems; w += blockDim.x) vert_phase2<D>(p.verts, tile, w, E, sG, a.pverts);
  }
}

template <int D>
__global__ void finalize_blocks_kernel(PlanView p, const double* pblocks, double* vals) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item < p.nnzb * D) block_finalize<D>(p, item, pblocks, vals);
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

template <int D>
__global__ void energy_kernel(PlanView p, EvalArgs a, double* block_sums) {
  __shared__ double sh[32];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double v = (e < p.t) ? energy_element<D>(p, a, e) : 0.0;
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
