// TEST INFRASTRUCTURE ONLY -- CPU replay of the kernel phase functions.
//
// The build container has no GPU.  The kernel bodies in simkit_b200/csrc are
// SKB_HD functions and the plan builder is backend-templated, so this harness
// runs them on the host (thrust::host, plain loops standing in for the thread
// grid) to debug index logic and element math before spending GPU time.  It is
// compiled into tests/_build/ by tests/hostsim.py, is never imported by
// simkit_b200/, and proves nothing about the GPU path: the -m gpu tests do that.
#include <stdint.h>
#include <vector>

#include "../simkit_b200/csrc/kernels.cuh"

using namespace skb;

template <int D>
static void run_assemble(const PlanView& p, const EvalArgs& a, std::vector<double>& esum) {
  const int E = p.tile_elems;
  std::vector<double> sK((size_t)Sizes<D>::NK * E), sG((size_t)Sizes<D>::NG * E);
  for (int tile = 0; tile < p.n_tiles; ++tile) {
    for (int le = 0; le < E; ++le) {
      int e = tile * E + le;
      if (e < p.t) element_phase1<D>(p, a, e, le, E, sK.data(), sG.data());
    }
    constexpr int K = D + 1, NP = K * (K + 1) / 2;
    if (a.want_hess) {
      int nitems = p.blocks.tl_ptr[tile + 1] - p.blocks.tl_ptr[tile];
      for (int w = 0; w < nitems; ++w)
        block_phase2<D>(p.blocks.tl_ent + p.blocks.tl_ptr[tile], p.blocks.tc_src + (size_t)tile * E * NP, w, E,
                        sK.data(), a.pblocks);
    }
    if (a.want_grad) {
      int nitems = p.verts.tl_ptr[tile + 1] - p.verts.tl_ptr[tile];
      for (int w = 0; w < nitems; ++w)
        vert_phase2<D>(p.verts.tl_ent + p.verts.tl_ptr[tile], p.verts.tc_src + (size_t)tile * E * K, w, E, sG.data(),
                       a.pverts);
    }
  }
#if defined(SKB_FIN_ITEMS)
  if (a.want_hess) {   // the A/B variant's thread grid: blocks of 128 threads, SKB_FIN_ITEMS items per thread
    const int n_items = p.nu * D * D, T = 128;
    for (int blk = 0; blk * T * SKB_FIN_ITEMS < n_items; ++blk)
      for (int th = 0; th < T; ++th)
        block_finalize_multi<D, SKB_FIN_ITEMS>(p, blk * (T * SKB_FIN_ITEMS) + th, T, n_items, a.pblocks, a.vals);
  }
#else
  if (a.want_hess)
    for (int item = 0; item < p.nu * D * D; ++item) block_finalize<D>(p, item, a.pblocks, a.vals);
#endif
  if (a.want_grad)
    for (int v = 0; v < p.n; ++v) vert_finalize<D>(p, v, a.pverts, a.g);
  double s = 0.0;
  for (int e = 0; e < p.t; ++e) s += energy_element<D>(p, a, e);
  esum.assign(1, s);
}

template <int D>
static void elem_hess(int material, int psd_mode, const double* F, double mu, double lam, double vol, double* H) {
  Mat<D> f, u, v; Vec<D> s;
  for (int i = 0; i < D * D; ++i) f.m[i / D][i % D] = F[i];
  if (material == MAT_LINEAR_ELASTICITY) {
    linear_elasticity_hessian<D>(mu * vol, lam * vol, H);
    return;
  }
  svd_rv(f, u, s, v);
  Principal<D> h = principal_hessian<D>(material, s, mu, lam);
  weight_and_project<D>(h, vol, psd_mode);
  expand_hessian<D>(h, u, v, H);
}

extern "C" {

// info: nnzb, n_block_partials, n_vertex_partials.  Outputs may be null (size query).
int hs_run(const double* X, const int64_t* T, int64_t n, int64_t t, int64_t t_active, int dim, int tile_elems, int material,
           int psd_mode, const double* x, const double* Fbar, const double* mu, int64_t mu_n,
           const double* lam, int64_t lam_n, const double* vol, int64_t vol_n, int64_t* info, int32_t* bptr,
           int32_t* bcol, int32_t* bslot, double* Dm_out, double* vol_out, double* g, double* vals,
           double* energy, int reorder) {
  const int K = dim + 1;
  thrust::host_vector<double> Xh(X, X + n * dim);
  thrust::host_vector<int> Th((size_t)t * K);
  for (int64_t i = 0; i < t * K; ++i) Th[i] = (int)T[i];
  PlanData<HostBackend> pd;
  // reorder != 0: the plan lists the active elements in its own spatial order (plan.cuh str_element_order), as
  // skb_plan_create does; per-element inputs and outputs below stay in the caller's order
  if (reorder) apply_element_order<HostBackend>(pd, Th, Xh, (int)n, (int)t_active, dim, tile_elems);
  if (!build_plan<HostBackend>(pd, Th, (int)n, (int)t_active, dim, tile_elems, (int)t)) return -1;
  t = t_active;
  set_geometry_from_X<HostBackend>(pd, Xh);
  info[0] = pd.nnzb;
  info[1] = pd.blocks.n_ts;
  info[2] = pd.verts.n_ts;
  if (!bptr) return 0;
  for (int i = 0; i <= n; ++i) bptr[i] = pd.bptr[i];
  for (int i = 0; i < pd.nnzb; ++i) bcol[i] = pd.bcol[i];
  (void)bslot;
  auto caller = [&](int i) { return pd.eorder.empty() ? i : (int)pd.eorder[i]; };
  for (int k = 0; k < dim * dim; ++k)
    for (int i = 0; i < t; ++i) Dm_out[(size_t)k * t + caller(i)] = pd.Dm[(size_t)k * t + i];
  for (int i = 0; i < t; ++i) vol_out[caller(i)] = pd.vol0[i];
  PlanView p = pd.view();
  std::vector<double> mu_i, lam_i, vol_i;  // per-element inputs in the internal order
  if (reorder) {
    auto gather = [&](const double*& src, int64_t cnt, std::vector<double>& dst) {
      if (!src || cnt <= 1) return;
      dst.resize(t);
      for (int i = 0; i < t; ++i) dst[i] = src[caller(i)];
      src = dst.data();
    };
    gather(mu, mu_n, mu_i);
    gather(lam, lam_n, lam_i);
    gather(vol, vol_n, vol_i);
  }
  std::vector<double> pb((size_t)pd.blocks.n_ts * (dim == 3 ? RecStride<3>::value : RecStride<2>::value)), pv((size_t)pd.verts.n_ts * dim);
  EvalArgs a;
  a.material = material;
  a.psd_mode = psd_mode;
  a.x = x;
  a.Fbar = Fbar;
  a.eorder = p.eorder;
  a.mu = mu;
  a.lam = lam;
  a.vol = vol ? vol : p.vol0;
  a.mu_stride = mu_n > 1;
  a.lam_stride = lam_n > 1;
  a.vol_stride = vol ? (vol_n > 1) : 1;
  a.want_grad = g != nullptr;
  a.want_hess = vals != nullptr;
  a.pblocks = pb.data();
  a.pverts = pv.data();
  a.g = g;
  a.vals = vals;
  std::vector<double> es;
  if (dim == 3) run_assemble<3>(p, a, es);
  else run_assemble<2>(p, a, es);
  if (energy) *energy = es[0];
  return 0;
}

// csr_value_position (kernels.cuh): where scalar entries (rows[k], cols[k]) sit in the canonical CSR values
int hs_value_positions(int dim, const int32_t* bptr, const int32_t* bcol, int64_t nq, const int32_t* rows,
                       const int32_t* cols, int32_t* out) {
  for (int64_t k = 0; k < nq; ++k)
    out[k] = dim == 3 ? csr_value_position<3>(bptr, bcol, rows[k], cols[k]) : csr_value_position<2>(bptr, bcol, rows[k], cols[k]);
  return 0;
}

// element-level checks of the principal-stretch machinery
int hs_svd(int dim, int64_t t, const double* F, double* U, double* S, double* V) {
  for (int64_t e = 0; e < t; ++e) {
    if (dim == 3) {
      Mat<3> f, u, v; Vec<3> s;
      for (int i = 0; i < 9; ++i) f.m[i / 3][i % 3] = F[e * 9 + i];
      svd_rv(f, u, s, v);
      for (int i = 0; i < 9; ++i) { U[e * 9 + i] = u.m[i / 3][i % 3]; V[e * 9 + i] = v.m[i / 3][i % 3]; }
      for (int i = 0; i < 3; ++i) S[e * 3 + i] = s[i];
    } else {
      Mat<2> f, u, v; Vec<2> s;
      for (int i = 0; i < 4; ++i) f.m[i / 2][i % 2] = F[e * 4 + i];
      svd_rv(f, u, s, v);
      for (int i = 0; i < 4; ++i) { U[e * 4 + i] = u.m[i / 2][i % 2]; V[e * 4 + i] = v.m[i / 2][i % 2]; }
      for (int i = 0; i < 2; ++i) S[e * 2 + i] = s[i];
    }
  }
  return 0;
}

int hs_element_hessian(int material, int psd_mode, int dim, int64_t t, const double* F, const double* mu,
                       const double* lam, const double* vol, double* H) {
  const int b = dim * dim;
  for (int64_t e = 0; e < t; ++e) {
    if (dim == 3) elem_hess<3>(material, psd_mode, F + e * b, mu[e], lam[e], vol[e], H + e * b * b);
    else elem_hess<2>(material, psd_mode, F + e * b, mu[e], lam[e], vol[e], H + e * b * b);
  }
  return 0;
}

int hs_element_energy_gradient(int material, int dim, int64_t t, const double* F, const double* mu,
                               const double* lam, double* psi, double* P) {
  const int b = dim * dim;
  for (int64_t e = 0; e < t; ++e) {
    if (dim == 3) {
      Mat<3> f;
      for (int i = 0; i < 9; ++i) f.m[i / 3][i % 3] = F[e * 9 + i];
      psi[e] = energy_density<3>(material, f, mu[e], lam[e]);
      Mat<3> p = pk1<3>(material, f, mu[e], lam[e]);
      for (int i = 0; i < 9; ++i) P[e * b + i] = p.m[i / 3][i % 3];
    } else {
      Mat<2> f;
      for (int i = 0; i < 4; ++i) f.m[i / 2][i % 2] = F[e * 4 + i];
      psi[e] = energy_density<2>(material, f, mu[e], lam[e]);
      Mat<2> p = pk1<2>(material, f, mu[e], lam[e]);
      for (int i = 0; i < 4; ++i) P[e * b + i] = p.m[i / 2][i % 2];
    }
  }
  return 0;
}

int hs_jacobi_dyn(int n, const double* A, double* w, double* V) {
  std::vector<double> a(A, A + n * n);
  jacobi_eig_dyn(a.data(), w, V, n);
  return 0;
}
}
