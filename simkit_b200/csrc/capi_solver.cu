// C ABI, part 3: SpMV, block-Jacobi PCG and the device-resident Newton step.
#include <vector>

#include "capi_common.cuh"
#include "solver.cuh"
#include "coarse.cuh"

namespace skb {

void coarse_destroy(CoarseSpace* c) { delete c; }

// Per solve: coarse matrix of the current Newton system and its dense inverse (Cholesky).
template <int D>
static int coarse_setup(skb_plan* pl, const double* vals, const double* dadd, cudaStream_t st) {
  CoarseSpace& c = *pl->coarse;
  int rc = coarse_assemble_launch<D>(pl, vals, dadd, raw(c.Ac), st);
  if (rc) return rc;
  return coarse_invert(pl, raw(c.Ac), CoarseDim<D>::NC * c.n_agg, st);
}

static CoarseView coarse_view(CoarseSpace& c, int ncper) {
  CoarseView v;
  v.n_agg = c.n_agg;
  v.nc = ncper * c.n_agg;
  v.agg = raw(c.agg);
  v.xrel = raw(c.xrel);
  v.vord = raw(c.vord);
  v.aptr = raw(c.aptr);
  v.Ainv = raw(c.Ac);
  v.rc = raw(c.rc);
  v.zc = raw(c.zc);
  return v;
}

static int pcg_grid(const skb_plan* pl) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
  // fixed grid: a multiple of the SM count, no more CTAs than there is work
  const int per_cta_rows = PCG_THREADS / SPMV_GROUP;
  int want = (pl->d.n + per_cta_rows - 1) / per_cta_rows;
  // 8 CTAs of 256 threads per SM = full occupancy: the PCG kernels are latency-bound streams (ncu: long-scoreboard
  // stalls, 4.3 TB/s at 4 CTAs per SM), bytes in flight scale with resident warps
  int grid = sms * PCG_CTAS_PER_SM;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  if (grid > PCG_MAX_GRID) grid = PCG_MAX_GRID;
  return grid;
}

struct PcgWork {
  double *r, *z, *p, *q, *dinv, *part;  // part: 3 * grid doubles
  PcgScalars* sc;                        // 2 ping-pong slots
};

static void ensure(dvec<double>& v, size_t n) {
  if (v.size() < n) v.resize(n);
}

// PCG driver shared by the plan (block) path and the generic CSR path.
//   spmv_dot(pvec, q, part_pq, sc) launches  q = A pvec  + partial sums of pvec.q
//   nb = number of D x D diagonal blocks, dinv already filled
template <int D, class SpmvDot>
static int pcg_loop(skb_plan* pl, int nb, int grid, SpmvDot spmv_dot, const double* dinv, const double* rhs,
                    double rtol, int max_iter, double* x, double* r, double* z, double* pv, double* q,
                    double* red, int* iters, double* relres, cudaStream_t st, const CoarseView* cv = nullptr) {
  double* part_pq = red;
  double* part_rz = part_pq + PCG_MAX_GRID;
  double* part_rr = part_rz + PCG_MAX_GRID;
  PcgScalars* sc = reinterpret_cast<PcgScalars*>(part_rr + PCG_MAX_GRID);  // two ping-pong slots
  // coarse correction of the two-level preconditioner: z += P Ainv P^T r, r.z partials recomputed
  auto coarse_apply = [&](double* p_or_null, const PcgScalars* s) {
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_restrict_kernel<D><<<cv->n_agg, 256, 0, st>>>(*cv, r, s));
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_gemv_kernel<<<(cv->nc * 32 + 255) / 256, 256, 0, st>>>(*cv, s));
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st,
               coarse_add_kernel<D><<<grid, PCG_THREADS, 0, st>>>(*cv, 0, nb, r, z, p_or_null, part_rz, s));
  };
  SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st,
             pcg_init_kernel<D><<<grid, PCG_THREADS, 0, st>>>(nb, rhs, dinv, x, r, z, pv, part_rz, part_rr));
  if (cv) coarse_apply(pv, nullptr);
  SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg_init_scalars_kernel<<<1, PCG_THREADS, 0, st>>>(part_rz, part_rr, grid, rtol, sc));
  int cur = 0;
  PcgScalars h;
  const int check_every = 25;
  int it = 0;
  bool done = false;
  while (it < max_iter && !done) {
    const int batch = (max_iter - it < check_every) ? (max_iter - it) : check_every;
    for (int b = 0; b < batch; ++b) {
      spmv_dot(pv, q, part_pq, sc + cur);
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st,
                 pcg_update_kernel<D><<<grid, PCG_THREADS, 0, st>>>(nb, dinv, pv, q, x, r, z, part_pq, grid, part_rz,
                                                                    part_rr, sc + cur));
      if (cv) coarse_apply(nullptr, sc + cur);
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st,
                 pcg_direction_kernel<D><<<grid, PCG_THREADS, 0, st>>>(nb, z, pv, part_rz, part_rr, grid, rtol, sc + cur,
                                                                       sc + (cur ^ 1)));
      cur ^= 1;
    }
    it += batch;
    SKB_CUDA(cudaMemcpyAsync(&h, sc + cur, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    SKB_CUDA(cudaStreamSynchronize(st));
    done = h.done != 0;
  }
  if (it == 0) {
    SKB_CUDA(cudaMemcpyAsync(&h, sc + cur, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    SKB_CUDA(cudaStreamSynchronize(st));
  }
  SKB_CUDA(cudaGetLastError());
  if (iters) *iters = h.iters;
  if (relres) *relres = (h.bb > 0.0) ? sqrt(h.rr / h.bb) : 0.0;
  return SKB_OK;
}

template <int D>
static int pcg_run(skb_plan* pl, const double* vals, const double* dadd, const double* rhs, double rtol,
                   int max_iter, double* x, int* iters, double* relres, cudaStream_t st) {
  const PlanView p = pl->view();
  const size_t nd = (size_t)p.n * D;
  const int grid = pcg_grid(pl);
  ensure(pl->w_r, nd);
  ensure(pl->w_z, nd);
  ensure(pl->w_p, nd);
  ensure(pl->w_q, nd);
  ensure(pl->w_dinv, (size_t)p.n * D * D);
  ensure(pl->w_red, 3 * PCG_MAX_GRID + 64);
  double* dinv = raw(pl->w_dinv);
  SKB_LAUNCH(pl, SKB_K_OTHER, st, block_jacobi_kernel<D><<<(p.n + 127) / 128, 128, 0, st>>>(p, vals, dadd, dinv));
  auto spmv_dot = [&](const double* pvec, double* q, double* part_pq, const PcgScalars* sc) {
    SKB_LAUNCH(pl, SKB_K_SPMV, st, pcg_spmv_dot_kernel<D><<<grid, PCG_THREADS, 0, st>>>(p, vals, dadd, pvec, q, part_pq, sc));
  };
  CoarseView cv;
  const bool two_level = pl->coarse != nullptr && pl->coarse->n_agg > 0;
  bool two_level_ok = two_level;
  if (two_level) {
    int rc = coarse_setup<D>(pl, vals, dadd, st);
    if (rc == SKB_ENOTSPD) two_level_ok = false;  // degenerate aggregate: block-Jacobi only for this solve
    else if (rc) return rc;
    else cv = coarse_view(*pl->coarse, CoarseDim<D>::NC);
  }
  return pcg_loop<D>(pl, p.n, grid, spmv_dot, dinv, rhs, rtol, max_iter, x, raw(pl->w_r), raw(pl->w_z),
                     raw(pl->w_p), raw(pl->w_q), raw(pl->w_red), iters, relres, st, two_level_ok ? &cv : nullptr);
}

// Builds the plan-side data of the coarse space from the vertex -> aggregate map.
int coarse_build(skb_plan* pl, int n_agg, const int* agg_h, const double* xrel_h, int v0, int v1) {
  if (pl->coarse) {
    coarse_destroy(pl->coarse);
    pl->coarse = nullptr;
  }
  if (n_agg <= 0) return SKB_OK;
  const PlanView p = pl->view();
  const int n = p.n, D = p.dim;
  const int ncper = D == 3 ? 6 : 3;
  for (int v = 0; v < n; ++v)
    if (agg_h[v] < 0 || agg_h[v] >= n_agg) return fail(SKB_EINVAL, "aggregate id out of range");
  CoarseSpace* c = new CoarseSpace();
  pl->coarse = c;
  c->n_agg = n_agg;
  c->v0 = v0;
  c->v1 = v1;
  c->agg.assign(agg_h, agg_h + n);
  c->xrel.assign(xrel_h, xrel_h + (size_t)n * D);
  // vertices sorted by aggregate
  {
    dvec<int> key(n);
    c->vord.resize(n);
    thrust::sequence(thrust::device, c->vord.begin(), c->vord.end());
    thrust::transform(thrust::device, c->vord.begin(), c->vord.end(), key.begin(), CoarseVertexKey{raw(c->agg), n_agg, v0, v1});
    thrust::stable_sort_by_key(thrust::device, key.begin(), key.end(), c->vord.begin());
    c->aptr.resize(n_agg + 1);
    thrust::counting_iterator<int> it0(0);
    thrust::lower_bound(thrust::device, key.begin(), key.end(), it0, it0 + n_agg + 1, c->aptr.begin());
  }
  // fine blocks sorted by coarse block (I, J)
  {
    const int nnzb = p.nnzb;
    dvec<uint64_t> key(nnzb);
    c->fb.resize(nnzb);
    thrust::sequence(thrust::device, c->fb.begin(), c->fb.end());
    thrust::transform(thrust::device, c->fb.begin(), c->fb.end(), key.begin(),
                      CoarseKeyOf{p.brow, p.bcol, raw(c->agg), n_agg, v0, v1});
    thrust::stable_sort_by_key(thrust::device, key.begin(), key.end(), c->fb.begin());
    const int nvalid = (int)(thrust::lower_bound(thrust::device, key.begin(), key.end(), ~0ull) - key.begin());
    dvec<uint64_t> ukey(nvalid);
    auto uend = thrust::unique_copy(thrust::device, key.begin(), key.begin() + nvalid, ukey.begin());
    c->n_cb = (int)(uend - ukey.begin());
    ukey.resize(c->n_cb);
    c->cb_ptr.resize(c->n_cb + 1);
    thrust::lower_bound(thrust::device, key.begin(), key.begin() + nvalid, ukey.begin(), ukey.end(), c->cb_ptr.begin());
    c->cb_ptr[c->n_cb] = nvalid;
    c->cb_I.resize(c->n_cb);
    c->cb_J.resize(c->n_cb);
    thrust::transform(thrust::device, ukey.begin(), ukey.end(), c->cb_I.begin(), CoarseKeyI{n_agg});
    thrust::transform(thrust::device, ukey.begin(), ukey.end(), c->cb_J.begin(), CoarseKeyJ{n_agg});
  }
  const size_t nc = (size_t)ncper * n_agg;
  c->Ac.resize(nc * nc);
  c->rc.resize(nc);
  c->zc.resize(nc);
  return SKB_OK;
}

int pcg_solve(skb_plan* pl, const double* vals, const double* dadd, const double* rhs, double rtol, int max_iter,
              double* x, int* iters, double* relres, cudaStream_t st) {
  return pl->d.dim == 3 ? pcg_run<3>(pl, vals, dadd, rhs, rtol, max_iter, x, iters, relres, st)
                        : pcg_run<2>(pl, vals, dadd, rhs, rtol, max_iter, x, iters, relres, st);
}

template <int D>
static int csr_pcg_run(int64_t n, const int32_t* indptr_h, const int32_t* indices_h, const double* vals_h,
                       const double* rhs_h, double rtol, int max_iter, double* x_h, int* iters, double* relres) {
  const int nb = (int)(n / D);
  const int64_t nnz = indptr_h[n];
  dvec<int> indptr(indptr_h, indptr_h + n + 1), indices(indices_h, indices_h + nnz);
  dvec<double> vals(vals_h, vals_h + nnz), rhs(rhs_h, rhs_h + n), x(n), r(n), z(n), pv(n), q(n),
      dinv((size_t)nb * D * D), red(3 * PCG_MAX_GRID + 64);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * 4;
  const int want = (int)((n + (PCG_THREADS / SPMV_GROUP) - 1) / (PCG_THREADS / SPMV_GROUP));
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  if (grid > PCG_MAX_GRID) grid = PCG_MAX_GRID;
  cudaStream_t st = 0;
  csr_block_jacobi_kernel<D><<<(nb + 127) / 128, 128, 0, st>>>(nb, raw(indptr), raw(indices), raw(vals), raw(dinv));
  const int* ip = raw(indptr);
  const int* ix = raw(indices);
  const double* vp = raw(vals);
  const int nn = (int)n;
  auto spmv_dot = [&](const double* pvec, double* qq, double* part_pq, const PcgScalars* sc) {
    csr_pcg_spmv_dot_kernel<<<grid, PCG_THREADS, 0, st>>>(nn, ip, ix, vp, pvec, qq, part_pq, sc);
  };
  int rc = pcg_loop<D>(nullptr, nb, grid, spmv_dot, raw(dinv), raw(rhs), rtol, max_iter, raw(x), raw(r), raw(z),
                       raw(pv), raw(q), raw(red), iters, relres, st);
  if (rc) return rc;
  SKB_CUDA(cudaMemcpy(x_h, raw(x), n * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
}

}  // namespace skb

using namespace skb;

extern "C" {

int skb_pcg_set_coarse(skb_plan* pl, int64_t n_agg, const int32_t* agg, const double* xrel) {
  if (!pl) return fail(SKB_EINVAL, "null argument");
  if (n_agg > 0 && (!agg || !xrel)) return fail(SKB_EINVAL, "null argument");
  if (n_agg > 2048) return fail(SKB_EINVAL, "at most 2048 aggregates (the coarse system is inverted densely)");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  return coarse_build(pl, (int)n_agg, agg, xrel, 0, pl->d.n);
  SKB_CATCH
}

int skb_newton_set_contact_plane(skb_plan* pl, double k, const double* p, const double* n, const double* weights) {
  if (!pl) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  if (!(k > 0.0) || !p || !n) {
    pl->contact_on = false;
    pl->contact_w.clear();
    return SKB_OK;
  }
  pl->contact_on = true;
  pl->contact_k = k;
  for (int i = 0; i < 3; ++i) {
    pl->contact_p[i] = i < pl->d.dim ? p[i] : 0.0;
    pl->contact_n[i] = i < pl->d.dim ? n[i] : 0.0;
  }
  if (weights) pl->contact_w.assign(weights, weights + pl->d.n);
  else pl->contact_w.clear();
  return SKB_OK;
  SKB_CATCH
}

static int contact_eval(int kind, int dim, int64_t nv, const double* X, double k, const double* p, const double* n, double r,
                        const double* weights, double* energy, double* grad, double* blocks, int32_t* under);

int skb_contact_springs_plane(int dim, int64_t nv, const double* X, double k, const double* p, const double* n,
                              const double* weights, double* energy, double* grad, double* blocks, int32_t* under) {
  if (!n) return fail(SKB_EINVAL, "null argument");
  return contact_eval(0, dim, nv, X, k, p, n, 0.0, weights, energy, grad, blocks, under);
}

int skb_contact_springs_sphere(int dim, int64_t nv, const double* X, double k, const double* p, double r,
                               const double* weights, double* energy, double* grad, double* blocks, int32_t* under) {
  return contact_eval(1, dim, nv, X, k, p, nullptr, r, weights, energy, grad, blocks, under);
}

int skb_newton_set_contact_sphere(skb_plan* pl, double k, const double* p, double r, const double* weights) {
  if (!pl) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  if (!(k > 0.0) || !p) {
    pl->sphere_on = false;
    pl->sphere_w.clear();
    return SKB_OK;
  }
  pl->sphere_on = true;
  pl->sphere_k = k;
  pl->sphere_r = r;
  for (int i = 0; i < 3; ++i) pl->sphere_p[i] = i < pl->d.dim ? p[i] : 0.0;
  if (weights) pl->sphere_w.assign(weights, weights + pl->d.n);
  else pl->sphere_w.clear();
  return SKB_OK;
  SKB_CATCH
}

static int contact_eval(int kind, int dim, int64_t nv, const double* X, double k, const double* p, const double* n, double r,
                        const double* weights, double* energy, double* grad, double* blocks, int32_t* under) {
  if (!X || !p) return fail(SKB_EINVAL, "null argument");
  if (dim != 2 && dim != 3) return fail(SKB_EINVAL, "Only dim == 2 or 3 are supported");
  if (nv <= 0 || nv >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "bad vertex count");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  dvec<double> Xd(X, X + nv * dim), gd, bd, wd, part(PCG_MAX_GRID), out(1);
  dvec<int> ud;
  ContactPlaneArgs c;
  c.k = k;
  c.kind = kind;
  c.r = r;
  for (int i = 0; i < 3; ++i) {
    c.p[i] = i < dim ? p[i] : 0.0;
    c.n[i] = (n && i < dim) ? n[i] : 0.0;
  }
  if (weights) wd.assign(weights, weights + nv);
  c.w = weights ? raw(wd) : nullptr;
  if (grad) gd.assign((size_t)nv * dim, 0.0);
  if (blocks) bd.resize((size_t)nv * dim * dim);
  if (under) ud.resize(nv);
  int grid = (int)((nv + PCG_THREADS - 1) / PCG_THREADS);
  if (grid > 1024) grid = 1024;
  if (dim == 3)
    contact_plane_kernel<3><<<grid, PCG_THREADS>>>((int)nv, raw(Xd), c, grad ? raw(gd) : nullptr, blocks ? raw(bd) : nullptr, nullptr, nullptr, raw(part), under ? raw(ud) : nullptr);
  else
    contact_plane_kernel<2><<<grid, PCG_THREADS>>>((int)nv, raw(Xd), c, grad ? raw(gd) : nullptr, blocks ? raw(bd) : nullptr, nullptr, nullptr, raw(part), under ? raw(ud) : nullptr);
  reduce_final_kernel<<<1, PCG_THREADS>>>(raw(part), grid, raw(out));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  if (energy) SKB_CUDA(cudaMemcpy(energy, raw(out), sizeof(double), cudaMemcpyDeviceToHost));
  if (grad) SKB_CUDA(cudaMemcpy(grad, raw(gd), gd.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (blocks) SKB_CUDA(cudaMemcpy(blocks, raw(bd), bd.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (under) SKB_CUDA(cudaMemcpy(under, raw(ud), ud.size() * sizeof(int), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

static int quad_grid(int64_t n) {
  int64_t grid = (n + PCG_THREADS - 1) / PCG_THREADS;
  if (grid > 1024) grid = 1024;
  return grid < 1 ? 1 : (int)grid;
}

int skb_quadratic(int64_t n, const int32_t* indptr, const int32_t* indices, const double* vals, const double* b,
                  const double* x, double* energy, double* grad) {
  if (!indptr || !x || (!energy && !grad)) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || n >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "bad size");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  const int64_t nnz = indptr[n];
  if (nnz < 0 || (nnz > 0 && (!indices || !vals))) return fail(SKB_EINVAL, "null argument");
  SKB_TRY
  dvec<int> dp(indptr, indptr + n + 1), dc;
  dvec<double> dv, db, gd, part(PCG_MAX_GRID), out(1);
  if (nnz > 0) {
    dc.assign(indices, indices + nnz);
    dv.assign(vals, vals + nnz);
  }
  dvec<double> dx(x, x + n);
  if (b) db.assign(b, b + n);
  if (grad) gd.assign((size_t)n, 0.0);
  const int grid = quad_grid(n);
  quad_term_kernel<<<grid, PCG_THREADS>>>((int)n, raw(dp), raw(dc), raw(dv), b ? raw(db) : nullptr, raw(dx),
                                          grad ? raw(gd) : nullptr, raw(part));
  reduce_final_kernel<<<1, PCG_THREADS>>>(raw(part), grid, raw(out));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  if (energy) SKB_CUDA(cudaMemcpy(energy, raw(out), sizeof(double), cudaMemcpyDeviceToHost));
  if (grad) SKB_CUDA(cudaMemcpy(grad, raw(gd), (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

int skb_newton_set_quadratic(skb_plan* pl, const int32_t* indptr, const int32_t* indices, const double* vals,
                             const double* b) {
  if (!pl) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->quad_on = false;
  if (!indptr) {
    pl->quad_ptr.clear();
    pl->quad_col.clear();
    pl->quad_pos.clear();
    pl->quad_val.clear();
    pl->quad_b.clear();
    return SKB_OK;
  }
  const int nd = (int)pl->ndof();
  const int64_t nnz = indptr[nd];
  if (nnz < 0 || nnz >= ((int64_t)1 << 31) || (nnz > 0 && (!indices || !vals))) return fail(SKB_EINVAL, "bad quadratic term");
  for (int64_t k = 0; k < nnz; ++k)
    if (indices[k] < 0 || indices[k] >= nd) return fail(SKB_EINVAL, "quadratic term: column index out of range");
  pl->quad_ptr.assign(indptr, indptr + nd + 1);
  pl->quad_pos.assign((size_t)(nnz > 0 ? nnz : 1), 0);
  if (nnz > 0) {
    pl->quad_col.assign(indices, indices + nnz);
    pl->quad_val.assign(vals, vals + nnz);
  } else {
    pl->quad_col.assign(1, 0);
    pl->quad_val.assign(1, 0.0);
  }
  if (b) pl->quad_b.assign(b, b + nd);
  else pl->quad_b.clear();
  dvec<int> bad(1, 0);
  const PlanView pv = pl->view();
  const int grid = (nd + PCG_THREADS - 1) / PCG_THREADS;
  if (pl->d.dim == 3)
    quad_positions_kernel<3><<<grid, PCG_THREADS, 0, pl->stream>>>(pv, nd, raw(pl->quad_ptr), raw(pl->quad_col), raw(pl->quad_pos), raw(bad));
  else
    quad_positions_kernel<2><<<grid, PCG_THREADS, 0, pl->stream>>>(pv, nd, raw(pl->quad_ptr), raw(pl->quad_col), raw(pl->quad_pos), raw(bad));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  const int nbad = bad[0];
  if (nbad != 0) return fail(SKB_EINVAL, "quadratic term: " + std::to_string(nbad) + " entries lie outside the mesh's CSR pattern");
  pl->quad_on = true;
  return SKB_OK;
  SKB_CATCH
}

int skb_spmv_dev(skb_plan* pl, const double* vals, const double* diag_add, const double* x, double* y,
                 void* stream) {
  if (!pl || !vals || !x || !y) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  const PlanView p = pl->view();
  const int grid = pcg_grid(pl);
  if (p.dim == 3)
    SKB_LAUNCH(pl, SKB_K_SPMV, (cudaStream_t)stream,
               spmv_kernel<3><<<grid, PCG_THREADS, 0, (cudaStream_t)stream>>>(p, vals, diag_add, x, y));
  else
    SKB_LAUNCH(pl, SKB_K_SPMV, (cudaStream_t)stream,
               spmv_kernel<2><<<grid, PCG_THREADS, 0, (cudaStream_t)stream>>>(p, vals, diag_add, x, y));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_pcg_dev(skb_plan* pl, const double* vals, const double* diag_add, const double* rhs, double rtol,
                int max_iter, double* x, int* iters, double* relres, void* stream) {
  if (!pl || !vals || !rhs || !x) return fail(SKB_EINVAL, "null argument");
  if (!(rtol >= 0.0) || max_iter < 0) return fail(SKB_EINVAL, "bad tolerance / max_iter");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  return pcg_solve(pl, vals, diag_add, rhs, rtol, max_iter, x, iters, relres, (cudaStream_t)stream);
  SKB_CATCH
}

int skb_pcg(skb_plan* pl, const double* vals, const double* diag_add, const double* rhs, double rtol,
            int max_iter, double* x, int* iters, double* relres) {
  if (!pl || !vals || !rhs || !x) return fail(SKB_EINVAL, "null argument");
  if (!(rtol >= 0.0) || max_iter < 0) return fail(SKB_EINVAL, "bad tolerance / max_iter");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  const size_t nd = pl->ndof();
  pl->vals.resize(pl->nnz());
  ensure(pl->w_dx, nd);
  ensure(pl->g, nd);
  SKB_CUDA(cudaMemcpyAsync(raw(pl->vals), vals, pl->nnz() * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  SKB_CUDA(cudaMemcpyAsync(raw(pl->g), rhs, nd * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  const double* dadd = nullptr;
  if (diag_add) {
    ensure(pl->w_diag, nd);
    SKB_CUDA(cudaMemcpyAsync(raw(pl->w_diag), diag_add, nd * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    dadd = raw(pl->w_diag);
  }
  int rc = pcg_solve(pl, raw(pl->vals), dadd, raw(pl->g), rtol, max_iter, raw(pl->w_dx), iters, relres, pl->stream);
  if (rc) return rc;
  SKB_CUDA(cudaMemcpyAsync(x, raw(pl->w_dx), nd * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

// The same solve on CSR values that are already resident on the device (the lazy Hessian the drop-in `*_hessian_x`
// functions return, simkit_b200/device_csr.py): nothing but the right-hand side and the solution cross PCIe.
int skb_pcg_vals_dev(skb_plan* pl, const double* vals_dev, const double* diag_add, const double* rhs, double rtol,
                int max_iter, double* x, int* iters, double* relres) {
  if (!pl || !vals_dev || !rhs || !x) return fail(SKB_EINVAL, "null argument");
  if (!(rtol >= 0.0) || max_iter < 0) return fail(SKB_EINVAL, "bad tolerance / max_iter");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  const size_t nd = pl->ndof();
  ensure(pl->w_dx, nd);
  ensure(pl->g, nd);
  SKB_CUDA(cudaMemcpyAsync(raw(pl->g), rhs, nd * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  const double* dadd = nullptr;
  if (diag_add) {
    ensure(pl->w_diag, nd);
    SKB_CUDA(cudaMemcpyAsync(raw(pl->w_diag), diag_add, nd * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    dadd = raw(pl->w_diag);
  }
  int rc = pcg_solve(pl, vals_dev, dadd, raw(pl->g), rtol, max_iter, raw(pl->w_dx), iters, relres, pl->stream);
  if (rc) return rc;
  SKB_CUDA(cudaMemcpyAsync(x, raw(pl->w_dx), nd * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"

template <int D>
static __global__ void value_positions_kernel(const int* bptr, const int* bcol, int64_t count, const int32_t* rows,
                                              const int32_t* cols, int32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = csr_value_position<D>(bptr, bcol, rows[i], cols[i]);
}

extern "C" {

// pos[i] = index of the scalar entry (rows[i], cols[i]) in the plan's canonical CSR value array, -1 if the entry is
// outside the pattern (device binary search over the block columns; host arrays in and out)
int skb_plan_value_positions(skb_plan* pl, int64_t count, const int32_t* rows, const int32_t* cols, int32_t* pos) {
  if (!pl || (count > 0 && (!rows || !cols || !pos))) return fail(SKB_EINVAL, "null argument");
  if (count <= 0) return SKB_OK;
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  dvec<int32_t> r(rows, rows + count), c(cols, cols + count), o(count);
  const unsigned grid = (unsigned)((count + 255) / 256);
  if (pl->d.dim == 3)
    value_positions_kernel<3><<<grid, 256, 0, pl->stream>>>(raw(pl->d.bptr), raw(pl->d.bcol), count, raw(r), raw(c), raw(o));
  else
    value_positions_kernel<2><<<grid, 256, 0, pl->stream>>>(raw(pl->d.bptr), raw(pl->d.bcol), count, raw(r), raw(c), raw(o));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaMemcpyAsync(pos, raw(o), count * sizeof(int32_t), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

int skb_csr_pcg(int64_t n, const int32_t* indptr, const int32_t* indices, const double* vals, int block,
                const double* rhs, double rtol, int max_iter, double* x, int* iters, double* relres) {
  if (!indptr || !indices || !vals || !rhs || !x) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || n >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "bad matrix size");
  if (block < 1 || block > 3 || n % block) return fail(SKB_EINVAL, "block must be 1, 2 or 3 and divide n");
  if (!(rtol >= 0.0) || max_iter < 0) return fail(SKB_EINVAL, "bad tolerance / max_iter");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  switch (block) {
    case 1: return csr_pcg_run<1>(n, indptr, indices, vals, rhs, rtol, max_iter, x, iters, relres);
    case 2: return csr_pcg_run<2>(n, indptr, indices, vals, rhs, rtol, max_iter, x, iters, relres);
    default: return csr_pcg_run<3>(n, indptr, indices, vals, rhs, rtol, max_iter, x, iters, relres);
  }
  SKB_CATCH
}

int skb_dense_solve(int64_t n, const double* A, const double* b, double* x) {
  if (!A || !b || !x) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || n > 8192) return fail(SKB_EINVAL, "dense solve supports 1 <= n <= 8192");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  dvec<double> Ad(A, A + n * n), bd(b, b + n), xd(n);
  dvec<int> status(1, 0);
  dense_solve_kernel<<<1, 1024>>>((int)n, raw(Ad), raw(bd), raw(xd), raw(status));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  int st = status[0];
  if (st) return fail(SKB_EINVAL, "Matrix is singular.");
  SKB_CUDA(cudaMemcpy(x, raw(xd), n * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

// One implicit step: Newton iterations with PCG and Armijo backtracking, all vectors resident.
// Mirrors solvers/newton.py:42-70 and backtracking_line_search.py:56-66.
int skb_newton(skb_plan* pl, const skb_newton_opts* o, const double* x0, const double* x_tilde,
               const double* mass, double kin_scale, const double* f_ext, const double* pin_k,
               const double* pin_target, double* x_out, skb_newton_info* info) {
  if (!pl || !o || !x0 || !x_out) return fail(SKB_EINVAL, "null argument");
  if (!pl->have_materials) return fail(SKB_EINVAL, "materials not set (skb_set_materials)");
  if (o->do_line_search) {
    // backtracking_line_search.py:52-53 (AssertionError in the reference)
    if (!(o->ls_alpha > 0 && o->ls_alpha <= 0.5) || !(o->ls_beta > 0 && o->ls_beta < 1))
      return fail(SKB_EINVAL, "line search needs 0 < alpha <= 0.5 and 0 < beta < 1");
  }
  if ((pin_k == nullptr) != (pin_target == nullptr)) return fail(SKB_EINVAL, "pin_k and pin_target go together");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  cudaStream_t st = pl->stream;
  const int nd = (int)pl->ndof();
  const size_t bytes = (size_t)nd * sizeof(double);
  ensure(pl->w_x, nd);
  ensure(pl->w_xtrial, nd);
  ensure(pl->w_dx, nd);
  ensure(pl->g, nd);
  ensure(pl->x, nd);  // rhs
  ensure(pl->w_diag, nd);
  pl->vals.resize(pl->nnz());
  ensure(pl->esums, 3 * PCG_MAX_GRID + 8);
  double* x = raw(pl->w_x);
  double* xtrial = raw(pl->w_xtrial);
  double* dx = raw(pl->w_dx);
  double* g = raw(pl->g);
  double* rhs = raw(pl->x);
  double* dadd = raw(pl->w_diag);
  SKB_CUDA(cudaMemcpyAsync(x, x0, bytes, cudaMemcpyHostToDevice, st));
  const double *d_xt = nullptr, *d_mass = nullptr, *d_f = nullptr, *d_pk = nullptr, *d_pt = nullptr;
  auto up = [&](dvec<double>& buf, const double* src) -> const double* {
    if (!src) return nullptr;
    ensure(buf, nd);
    cudaMemcpyAsync(raw(buf), src, bytes, cudaMemcpyHostToDevice, st);
    return raw(buf);
  };
  d_xt = up(pl->w_xtilde, x_tilde);
  d_mass = up(pl->w_mass, mass);
  d_f = up(pl->w_fext, f_ext);
  dvec<double> pinbuf;
  if (pin_k) {
    pinbuf.resize(2 * (size_t)nd);
    SKB_CUDA(cudaMemcpyAsync(raw(pinbuf), pin_k, bytes, cudaMemcpyHostToDevice, st));
    SKB_CUDA(cudaMemcpyAsync(raw(pinbuf) + nd, pin_target, bytes, cudaMemcpyHostToDevice, st));
    d_pk = raw(pinbuf);
    d_pt = raw(pinbuf) + nd;
  }
  const int vgrid = pcg_grid(pl);
  dvec<double> parts(3 * PCG_MAX_GRID + 8);
  double* part_e = raw(parts);
  double* part_g = part_e + PCG_MAX_GRID;
  double* part_d = part_g + PCG_MAX_GRID;
  double* red = part_d + PCG_MAX_GRID;  // 3 sums + elastic energy + up to 2 contact energies
  double hred[7];
  ContactPlaneArgs contacts[2];
  int n_contacts = 0;
  dvec<double> part_c;
  dvec<PlanView> pview_d;
  if (pl->contact_on) {
    ContactPlaneArgs& c = contacts[n_contacts++];
    c.kind = 0;
    c.r = 0.0;
    c.k = pl->contact_k;
    for (int i = 0; i < 3; ++i) {
      c.p[i] = pl->contact_p[i];
      c.n[i] = pl->contact_n[i];
    }
    c.w = pl->contact_w.empty() ? nullptr : raw(pl->contact_w);
  }
  if (pl->sphere_on) {
    ContactPlaneArgs& c = contacts[n_contacts++];
    c.kind = 1;
    c.r = pl->sphere_r;
    c.k = pl->sphere_k;
    for (int i = 0; i < 3; ++i) {
      c.p[i] = pl->sphere_p[i];
      c.n[i] = 0.0;
    }
    c.w = pl->sphere_w.empty() ? nullptr : raw(pl->sphere_w);
  }
  if (n_contacts) {
    part_c.resize(PCG_MAX_GRID);
    const PlanView hv = pl->view();
    pview_d.assign(&hv, &hv + 1);
  }
  const int nverts = pl->d.n;
  const int dim_ = pl->d.dim;
  // general sparse quadratic term (skb_newton_set_quadratic)
  dvec<double> part_q;
  const int qgrid = quad_grid(nd);
  const int qnnz = pl->quad_on ? (int)pl->quad_val.size() : 0;
  const double* quad_b = (pl->quad_on && !pl->quad_b.empty()) ? raw(pl->quad_b) : nullptr;
  if (pl->quad_on) part_q.resize(PCG_MAX_GRID);

  // total energy at x + s*dx (also leaves the trial point in xtrial)
  auto total_energy = [&](double s, const double* dxp, bool with_gdx, double& e_tot, double& gdx, double& dx2) -> int {
    SKB_LAUNCH(pl, SKB_K_OTHER, st,
               newton_energy_terms_kernel<<<vgrid, PCG_THREADS, 0, st>>>(nd, x, dxp, s, d_f, d_mass, d_xt, kin_scale, d_pk,
                                                                         d_pt, with_gdx ? g : nullptr, xtrial, part_e,
                                                                         part_g, part_d));
    SKB_LAUNCH(pl, SKB_K_OTHER, st, reduce3_kernel<<<1, PCG_THREADS, 0, st>>>(part_e, part_g, part_d, vgrid, red));
    EvalArgs a;
    int rc = make_args(pl, o->material, PSD_NONE, xtrial, nullptr, nullptr, nullptr, a);
    if (rc) return rc;
    rc = launch_energy(pl, a, red + 3, st);
    if (rc) return rc;
    hred[4] = hred[5] = 0.0;
    for (int ci = 0; ci < n_contacts; ++ci) {
      if (dim_ == 3)
        SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<3><<<vgrid, PCG_THREADS, 0, st>>>(nverts, xtrial, contacts[ci], nullptr, nullptr, nullptr, nullptr, raw(part_c), nullptr));
      else
        SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<2><<<vgrid, PCG_THREADS, 0, st>>>(nverts, xtrial, contacts[ci], nullptr, nullptr, nullptr, nullptr, raw(part_c), nullptr));
      SKB_LAUNCH(pl, SKB_K_OTHER, st, reduce_final_kernel<<<1, PCG_THREADS, 0, st>>>(raw(part_c), vgrid, red + 4 + ci));
    }
    hred[6] = 0.0;
    if (pl->quad_on) {
      SKB_LAUNCH(pl, SKB_K_OTHER, st,
                 quad_term_kernel<<<qgrid, PCG_THREADS, 0, st>>>(nd, raw(pl->quad_ptr), raw(pl->quad_col), raw(pl->quad_val), quad_b,
                                                                  xtrial, nullptr, raw(part_q)));
      SKB_LAUNCH(pl, SKB_K_OTHER, st, reduce_final_kernel<<<1, PCG_THREADS, 0, st>>>(raw(part_q), qgrid, red + 6));
    }
    SKB_CUDA(cudaMemcpyAsync(hred, red, (pl->quad_on ? 7 : 4 + n_contacts) * sizeof(double), cudaMemcpyDeviceToHost, st));
    SKB_CUDA(cudaStreamSynchronize(st));
    if (!pl->quad_on) hred[6] = 0.0;
    if (n_contacts < 2) hred[5] = 0.0;
    if (n_contacts < 1) hred[4] = 0.0;
    e_tot = hred[0] + hred[3] + hred[4] + hred[5] + hred[6];
    gdx = hred[1];
    dx2 = hred[2];
    return SKB_OK;
  };

  if (info) {
    memset(info, 0, sizeof(*info));
    info->iters = -1;
  }
  for (int it = 0; it < o->max_iter; ++it) {
    EvalArgs a;
    int rc = make_args(pl, o->material, o->psd_mode, x, nullptr, g, raw(pl->vals), a);
    if (rc) return rc;
    rc = launch_assemble(pl, a, st);
    if (rc) return rc;
    for (int ci = 0; ci < n_contacts; ++ci) {
      // contact springs: gradient into g (before rhs = -g is formed), Hessian blocks into the diagonal blocks of vals
      if (dim_ == 3)
        SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<3><<<vgrid, PCG_THREADS, 0, st>>>(nverts, x, contacts[ci], g, nullptr, raw(pview_d), raw(pl->vals), nullptr, nullptr));
      else
        SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<2><<<vgrid, PCG_THREADS, 0, st>>>(nverts, x, contacts[ci], g, nullptr, raw(pview_d), raw(pl->vals), nullptr, nullptr));
    }
    if (pl->quad_on) {
      // gradient Q x + b into g, Hessian Q into its slots of the CSR values
      SKB_LAUNCH(pl, SKB_K_OTHER, st,
                 quad_term_kernel<<<qgrid, PCG_THREADS, 0, st>>>(nd, raw(pl->quad_ptr), raw(pl->quad_col), raw(pl->quad_val), quad_b,
                                                                  x, g, nullptr));
      SKB_LAUNCH(pl, SKB_K_OTHER, st,
                 add_at_kernel<<<quad_grid(qnnz), PCG_THREADS, 0, st>>>(raw(pl->vals), raw(pl->quad_pos), qnnz, raw(pl->quad_val)));
    }
    SKB_LAUNCH(pl, SKB_K_OTHER, st,
               newton_gradient_kernel<<<vgrid, PCG_THREADS, 0, st>>>(nd, x, d_f, d_mass, d_xt, kin_scale, d_pk, d_pt, g, rhs,
                                                                     dadd));
    int pit = 0;
    double relres = 0.0;
    rc = pcg_solve(pl, raw(pl->vals), dadd, rhs, o->pcg_rtol, o->pcg_max_iter, dx, &pit, &relres, st);
    if (rc) return rc;
    double alpha = 1.0, e0 = 0, gdx = 0, dx2 = 0, e1 = 0, t1, t2;
    if (o->do_line_search) {
      rc = total_energy(0.0, dx, true, e0, gdx, dx2);
      if (rc) return rc;
      double tstep = 1.0;
      bool ok = false;
      for (int ls = 0; ls < o->ls_max_iter; ++ls) {
        rc = total_energy(tstep, dx, false, e1, t1, t2);
        if (rc) return rc;
        if (e1 <= e0 + o->ls_alpha * tstep * gdx + o->ls_threshold) {
          ok = true;
          break;
        }
        tstep *= o->ls_beta;
      }
      alpha = ok ? tstep : 0.0;
    } else {
      rc = total_energy(1.0, dx, false, e1, t1, dx2);  // leaves x + dx in xtrial, |dx|^2 in dx2
      if (rc) return rc;
    }
    if (alpha == 1.0 || !o->do_line_search) {
      SKB_CUDA(cudaMemcpyAsync(x, xtrial, bytes, cudaMemcpyDeviceToDevice, st));
    } else if (alpha > 0.0) {
      // xtrial currently holds x + alpha*dx (the last accepted trial)
      SKB_CUDA(cudaMemcpyAsync(x, xtrial, bytes, cudaMemcpyDeviceToDevice, st));
    }
    const double step_norm = alpha * sqrt(dx2);
    if (info) {
      info->iters = it;
      info->pcg_iters_total += pit;
      info->last_alpha = alpha;
      info->last_step_norm = step_norm;
      info->last_pcg_relres = relres;
      if (it < 64) info->alphas[it] = alpha;
    }
    if (step_norm < o->tolerance) break;
  }
  SKB_CUDA(cudaMemcpyAsync(x_out, x, bytes, cudaMemcpyDeviceToHost, st));
  SKB_CUDA(cudaStreamSynchronize(st));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"
