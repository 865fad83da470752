// C ABI, part 5: building blocks of the DISTRIBUTED block-Jacobi PCG (one rank of an element-sharded
// mesh, simkit_b200/sharding.py).  A rank owns the vertex rows [v0, v1) of its local numbering; its
// rows of the matrix are globally complete after the interface exchange.  Every call below is one
// stream-ordered step on device data; the scalars they produce / consume live in a small device array
// so that the host can all-reduce them (NCCL) between the steps without ever reading them back:
//
//   init      : dinv = inv(diag blocks), x = 0, r = rhs, z = dinv r, p = z     -> s[0] = r.z, s[1] = r.r   (owned)
//   spmv_dot  : q = (A + diag) p on the owned rows                              -> s[2] = p.q              (owned)
//   update    : alpha = s[0] / s[2]; x += alpha p; r -= alpha q; z = dinv r     -> s[3] = r.z, s[4] = r.r
//   direction : beta = s[3] / s[0]; p = z + beta p; then s[0] = s[3], s[1] = s[4]
//
// Reductions use a fixed grid and a fixed tree: bitwise reproducible for a fixed rank count.
#include "capi_common.cuh"
#include "solver.cuh"
#include "coarse.cuh"

namespace skb {

constexpr int DIST_GRID = 148 * PCG_CTAS_PER_SM;  // full occupancy on a B200; fixed so that the summation tree is fixed

template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) dist_init_kernel(PlanView p, const double* vals, const double* dadd, int v0, int v1, const double* rhs,
                                 double* dinv, double* x, double* r, double* z, double* pv, double* part) {
  __shared__ double sh[32];
  double rz = 0.0, rr = 0.0;
  for (int v = v0 + blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += gridDim.x * blockDim.x) {
    const int b0 = p.bptr[v];
    const int nb = p.bptr[v + 1] - b0;
    int jd = -1;
    for (int j = 0; j < nb; ++j)
      if (p.bcol[b0 + j] == v) jd = j;
    Mat<D> A;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double a = (jd >= 0) ? vals[(size_t)b0 * (D * D) + (size_t)i * nb * D + (size_t)jd * D + k] : 0.0;
        if (i == k && dadd) a += dadd[(size_t)v * D + i];
        A.m[i][k] = a;
      }
    const double inv = 1.0 / det(A);
    Mat<D> c = cofactor(A);
    double rl[D], zl[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      rl[i] = rhs[(size_t)v * D + i];
#pragma unroll
      for (int k = 0; k < D; ++k) dinv[(size_t)v * (D * D) + i * D + k] = c.m[k][i] * inv;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(c.m[k][i] * inv, rl[k], s);
      zl[i] = s;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const size_t k = (size_t)v * D + i;
      x[k] = 0.0;
      r[k] = rl[i];
      z[k] = zl[i];
      pv[k] = zl[i];
      rz = fma(rl[i], zl[i], rz);
      rr = fma(rl[i], rl[i], rr);
    }
  }
  rz = block_reduce_sum(rz, sh);
  rr = block_reduce_sum(rr, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = rz;
    part[gridDim.x + blockIdx.x] = rr;
  }
}

// y = (A + diag) x for the block rows [v0, v1); per-thread partial of x.y
template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) dist_spmv_dot_kernel(PlanView p, const double* __restrict__ vals, const double* __restrict__ dadd,
                                     const double* __restrict__ x, double* __restrict__ y, int v0, int v1, double* part) {
  __shared__ double sh[32];
  constexpr int GW = 32 / SPMV_GROUP;
  const int lane = threadIdx.x & (SPMV_GROUP - 1);
  const int gw = (threadIdx.x & 31) / SPMV_GROUP;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int vb = v0 + warp * GW; vb < v1; vb += nwarps * GW) {
    const int v = vb + gw;
    const bool valid = v < v1;
    const int b0 = valid ? p.bptr[v] : 0;
    const int nb = valid ? p.bptr[v + 1] - b0 : 0;
    const int ncol = nb * D;
    const double* rowbase = vals + (size_t)b0 * (D * D);
    double acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.0;
    for (int idx = lane; idx < ncol; idx += SPMV_GROUP) {
      const int j = idx / D;
      const int k = idx - j * D;
      const double xv = x[(size_t)p.bcol[b0 + j] * D + k];
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] = fma(rowbase[(size_t)i * ncol + idx], xv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int o = SPMV_GROUP / 2; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o, SPMV_GROUP);
    }
    if (lane == 0 && valid) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const size_t r = (size_t)v * D + i;
        double yi = acc[i];
        if (dadd) yi = fma(dadd[r], x[r], yi);
        y[r] = yi;
        dot = fma(x[r], yi, dot);
      }
    }
  }
  dot = block_reduce_sum(dot, sh);
  if (threadIdx.x == 0 && part) part[blockIdx.x] = dot;
}

template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) dist_update_kernel(int v0, int v1, const double* dinv, const double* pv, const double* q, double* x,
                                   double* r, double* z, const double* s, double* part) {
  __shared__ double sh[32];
  const double alpha = (s[2] != 0.0) ? s[0] / s[2] : 0.0;  // p == 0: already converged exactly
  double rz = 0.0, rr = 0.0;
  for (int v = v0 + blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += gridDim.x * blockDim.x) {
    double rl[D], zl[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const size_t k = (size_t)v * D + i;
      x[k] = fma(alpha, pv[k], x[k]);
      rl[i] = fma(-alpha, q[k], r[k]);
      r[k] = rl[i];
    }
    apply_dinv<D>(dinv, v, rl, zl);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      z[(size_t)v * D + i] = zl[i];
      rz = fma(rl[i], zl[i], rz);
      rr = fma(rl[i], rl[i], rr);
    }
  }
  rz = block_reduce_sum(rz, sh);
  rr = block_reduce_sum(rr, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = rz;
    part[gridDim.x + blockIdx.x] = rr;
  }
}

// out[o0] = sum part[0..n), out[o1] = sum part[n..2n)   (single CTA, fixed order); o1 < 0: one sum only
static __global__ void dist_reduce_kernel(const double* part, int n, double* out, int o0, int o1) {
  __shared__ double sh[32];
  const double a = reduce_partials(part, n, sh);
  double b = 0.0;
  if (o1 >= 0) b = reduce_partials(part + n, n, sh);
  if (threadIdx.x == 0) {
    out[o0] = a;
    if (o1 >= 0) out[o1] = b;
  }
}

// p = z + (s[3]/s[0]) p on the owned dofs; afterwards s[0] = s[3], s[1] = s[4]  (done by the last CTA to finish
// reading: a separate tiny kernel keeps it simple and race-free)
static __global__ void dist_direction_kernel(int k0, int k1, const double* z, double* pv, const double* s) {
  const double beta = (s[0] != 0.0) ? s[3] / s[0] : 0.0;
  for (int k = k0 + blockIdx.x * blockDim.x + threadIdx.x; k < k1; k += gridDim.x * blockDim.x)
    pv[k] = fma(beta, pv[k], z[k]);
}
static __global__ void dist_roll_kernel(double* s) {
  s[0] = s[3];
  s[1] = s[4];
}

// total gradient / rhs / diagonal of the implicit step on the owned dofs (same formula as newton_gradient_kernel)
// and the non-elastic energy terms + g.dx + |dx|^2 partial sums over the owned dofs
static __global__ void dist_axpy_kernel(int k0, int k1, double s, const double* dx, const double* x, double* out) {
  for (int k = k0 + blockIdx.x * blockDim.x + threadIdx.x; k < k1; k += gridDim.x * blockDim.x)
    out[k] = fma(s, dx[k], x[k]);
}

}  // namespace skb

using namespace skb;

extern "C" {

#define DIST_CHECK(pl)                                              \
  if (!(pl)) return fail(SKB_EINVAL, "null plan");                  \
  if (v0 < 0 || v1 > (pl)->d.n || v0 > v1) return fail(SKB_EINVAL, "bad owned row range"); \
  SKB_CUDA(cudaSetDevice((pl)->device));

int skb_dist_pcg_init_dev(skb_plan* pl, const double* vals, const double* diag_add, int v0, int v1, const double* rhs,
                          double* dinv, double* x, double* r, double* z, double* p, double* scalars, double* work,
                          void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  const PlanView pv = pl->view();
  if (pv.dim == 3)
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, dist_init_kernel<3><<<DIST_GRID, PCG_THREADS, 0, st>>>(pv, vals, diag_add, v0, v1, rhs, dinv, x, r, z, p, work));
  else
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, dist_init_kernel<2><<<DIST_GRID, PCG_THREADS, 0, st>>>(pv, vals, diag_add, v0, v1, rhs, dinv, x, r, z, p, work));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(work, DIST_GRID, scalars, 0, 1));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_dist_spmv_dot_dev(skb_plan* pl, const double* vals, const double* diag_add, int v0, int v1, const double* p,
                          double* q, double* scalars, double* work, void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  const PlanView pv = pl->view();
  if (pv.dim == 3)
    SKB_LAUNCH(pl, SKB_K_SPMV, st, dist_spmv_dot_kernel<3><<<DIST_GRID, PCG_THREADS, 0, st>>>(pv, vals, diag_add, p, q, v0, v1, work));
  else
    SKB_LAUNCH(pl, SKB_K_SPMV, st, dist_spmv_dot_kernel<2><<<DIST_GRID, PCG_THREADS, 0, st>>>(pv, vals, diag_add, p, q, v0, v1, work));
  if (scalars) SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(work, DIST_GRID, scalars, 2, -1));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_dist_pcg_update_dev(skb_plan* pl, int v0, int v1, const double* dinv, const double* p, const double* q, double* x,
                            double* r, double* z, double* scalars, double* work, void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->d.dim == 3)
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, dist_update_kernel<3><<<DIST_GRID, PCG_THREADS, 0, st>>>(v0, v1, dinv, p, q, x, r, z, scalars, work));
  else
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, dist_update_kernel<2><<<DIST_GRID, PCG_THREADS, 0, st>>>(v0, v1, dinv, p, q, x, r, z, scalars, work));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(work, DIST_GRID, scalars, 3, 4));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_dist_pcg_direction_dev(skb_plan* pl, int v0, int v1, const double* z, double* p, double* scalars, void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  const int D = pl->d.dim;
  SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, dist_direction_kernel<<<DIST_GRID, PCG_THREADS, 0, st>>>(v0 * D, v1 * D, z, p, scalars));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_roll_kernel<<<1, 1, 0, st>>>(scalars));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// ---- two-level preconditioner on a sharded mesh (csrc/coarse.cuh) -------------------------------------------
// The coarse space is global (same aggregates on every rank); a rank contributes the fine blocks of its OWNED
// rows to the coarse matrix and its owned vertices to the restriction, the host side all-reduces both (the
// coarse matrix once per solve, the restricted residual once per iteration) and every rank applies the same
// dense inverse.  Buffers (Ac, rc, zc) belong to the caller.
int skb_dist_coarse_set(skb_plan* pl, int64_t n_agg, const int32_t* agg, const double* xrel, int v0, int v1) {
  DIST_CHECK(pl)
  if (n_agg > 0 && (!agg || !xrel)) return fail(SKB_EINVAL, "null argument");
  if (n_agg > 2048) return fail(SKB_EINVAL, "at most 2048 aggregates (the coarse system is inverted densely)");
  SKB_TRY
  return coarse_build(pl, (int)n_agg, agg, xrel, v0, v1);
  SKB_CATCH
}

int skb_dist_coarse_assemble_dev(skb_plan* pl, const double* vals, const double* diag_add, double* Ac, void* stream) {
  if (!pl || !pl->coarse || !vals || !Ac) return fail(SKB_EINVAL, "null argument / no coarse space");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  return pl->d.dim == 3 ? coarse_assemble_launch<3>(pl, vals, diag_add, Ac, (cudaStream_t)stream)
                        : coarse_assemble_launch<2>(pl, vals, diag_add, Ac, (cudaStream_t)stream);
  SKB_CATCH
}

int skb_dist_coarse_invert_dev(skb_plan* pl, double* Ac, void* stream) {
  if (!pl || !pl->coarse || !Ac) return fail(SKB_EINVAL, "null argument / no coarse space");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  return coarse_invert(pl, Ac, (pl->d.dim == 3 ? 6 : 3) * pl->coarse->n_agg, (cudaStream_t)stream);
  SKB_CATCH
}

static CoarseView dist_coarse_view(skb_plan* pl, const double* Ainv, double* rc, double* zc) {
  CoarseSpace& c = *pl->coarse;
  CoarseView v;
  v.n_agg = c.n_agg;
  v.nc = (pl->d.dim == 3 ? 6 : 3) * c.n_agg;
  v.agg = raw(c.agg);
  v.xrel = raw(c.xrel);
  v.vord = raw(c.vord);
  v.aptr = raw(c.aptr);
  v.Ainv = Ainv;
  v.rc = rc;
  v.zc = zc;
  return v;
}

// rc = sum over the owned vertices of P_v^T r_v  (to be all-reduced by the caller)
int skb_dist_coarse_restrict_dev(skb_plan* pl, const double* r, double* rc, void* stream) {
  if (!pl || !pl->coarse || !r || !rc) return fail(SKB_EINVAL, "null argument / no coarse space");
  SKB_CUDA(cudaSetDevice(pl->device));
  cudaStream_t st = (cudaStream_t)stream;
  const CoarseView cv = dist_coarse_view(pl, nullptr, rc, nullptr);
  if (pl->d.dim == 3)
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_restrict_kernel<3><<<cv.n_agg, 256, 0, st>>>(cv, r, nullptr));
  else
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_restrict_kernel<2><<<cv.n_agg, 256, 0, st>>>(cv, r, nullptr));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// zc = Ainv rc;  z += P zc on the owned vertices (p = z too when p != NULL);  scalars[slot] = r.z over them
int skb_dist_coarse_correct_dev(skb_plan* pl, int v0, int v1, const double* Ainv, double* rc, double* zc, const double* r,
                                double* z, double* p, double* scalars, int slot, double* work, void* stream) {
  DIST_CHECK(pl)
  if (!pl->coarse || !Ainv || !rc || !zc || !r || !z || !scalars || !work) return fail(SKB_EINVAL, "null argument / no coarse space");
  cudaStream_t st = (cudaStream_t)stream;
  const CoarseView cv = dist_coarse_view(pl, Ainv, rc, zc);
  SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_gemv_kernel<<<(cv.nc * 32 + 255) / 256, 256, 0, st>>>(cv, nullptr));
  if (pl->d.dim == 3)
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_add_kernel<3><<<DIST_GRID, PCG_THREADS, 0, st>>>(cv, v0, v1, r, z, p, work, nullptr));
  else
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_add_kernel<2><<<DIST_GRID, PCG_THREADS, 0, st>>>(cv, v0, v1, r, z, p, work, nullptr));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(work, DIST_GRID, scalars, slot, -1));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// Newton-step vector pieces on the owned dofs [v0*dim, v1*dim):
//   skb_dist_newton_rhs_dev    g += -f_ext + kin*mass*(x - x_tilde) + pin_k*(x - pin_t); rhs = -g; diag = kin*mass + pin_k
//   skb_dist_newton_terms_dev  xtrial = x + s*dx (owned dofs); out[0] = non-elastic energy at xtrial, out[1] = g.dx,
//                              out[2] = |dx|^2, all restricted to the owned dofs (the caller all-reduces them)
int skb_dist_newton_rhs_dev(skb_plan* pl, int v0, int v1, const double* x, const double* f_ext, const double* mass,
                            const double* x_tilde, double kin_scale, const double* pin_k, const double* pin_t, double* g,
                            double* rhs, double* diag, void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  const int D = pl->d.dim;
  const int off = v0 * D, n = (v1 - v0) * D;
  auto o = [&](const double* p_) { return p_ ? p_ + off : nullptr; };
  SKB_LAUNCH(pl, SKB_K_OTHER, st,
             newton_gradient_kernel<<<DIST_GRID, PCG_THREADS, 0, st>>>(n, x + off, o(f_ext), o(mass), o(x_tilde), kin_scale,
                                                                      o(pin_k), o(pin_t), g + off, rhs + off, diag + off));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// Contact springs (energies/contact_springs_plane.py:245-388, contact_springs_sphere.py) on the owned vertices [v0, v1):
// kind 0 = plane (point p, normal nrm), 1 = sphere (centre p, radius r); w = per-vertex weights (device, local
// numbering) or NULL.  g / vals non-NULL: gradient added into g, k m_v n n^T into the diagonal blocks of vals.
// energy_out non-NULL: this rank's part of the energy (the caller all-reduces it).
int skb_dist_contact_dev(skb_plan* pl, int v0, int v1, const double* x, int kind, double k, const double* p, const double* nrm,
                         double r, const double* w, double* g, double* vals, double* energy_out, double* work, void* stream) {
  DIST_CHECK(pl)
  if (!x || !p || (kind == 0 && !nrm) || (energy_out && !work)) return fail(SKB_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  SKB_TRY
  const int D = pl->d.dim;
  ContactPlaneArgs c;
  memset(&c, 0, sizeof(c));
  c.k = k;
  for (int i = 0; i < D; ++i) {
    c.p[i] = p[i];
    c.n[i] = nrm ? nrm[i] : 0.0;
  }
  c.w = w;
  c.kind = kind;
  c.r = r;
  const PlanView* pvd = nullptr;
  if (vals) {
    if (pl->pview_dev.size() != 1) pl->pview_dev.resize(1);
    const PlanView hv = pl->view();
    SKB_CUDA(cudaMemcpyAsync(raw(pl->pview_dev), &hv, sizeof(PlanView), cudaMemcpyHostToDevice, st));
    pvd = raw(pl->pview_dev);
  }
  if (D == 3)
    SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<3><<<DIST_GRID, PCG_THREADS, 0, st>>>(v1, x, c, g, nullptr, pvd, vals, energy_out ? work : nullptr, nullptr, v0));
  else
    SKB_LAUNCH(pl, SKB_K_OTHER, st, contact_plane_kernel<2><<<DIST_GRID, PCG_THREADS, 0, st>>>(v1, x, c, g, nullptr, pvd, vals, energy_out ? work : nullptr, nullptr, v0));
  if (energy_out) SKB_LAUNCH(pl, SKB_K_OTHER, st, dist_reduce_kernel<<<1, PCG_THREADS, 0, st>>>(work, DIST_GRID, energy_out, 0, -1));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
  SKB_CATCH
}

int skb_dist_newton_terms_dev(skb_plan* pl, int v0, int v1, const double* x, const double* dx, double s,
                              const double* f_ext, const double* mass, const double* x_tilde, double kin_scale,
                              const double* pin_k, const double* pin_t, const double* g, double* xtrial, double* out,
                              double* work, void* stream) {
  DIST_CHECK(pl)
  cudaStream_t st = (cudaStream_t)stream;
  const int D = pl->d.dim;
  const int off = v0 * D, n = (v1 - v0) * D;
  auto o = [&](const double* p_) { return p_ ? p_ + off : nullptr; };
  SKB_LAUNCH(pl, SKB_K_OTHER, st,
             newton_energy_terms_kernel<<<DIST_GRID, PCG_THREADS, 0, st>>>(n, x + off, dx ? dx + off : nullptr, s, o(f_ext),
                                                                          o(mass), o(x_tilde), kin_scale, o(pin_k), o(pin_t),
                                                                          g ? g + off : nullptr, xtrial ? xtrial + off : nullptr,
                                                                          work, work + DIST_GRID, work + 2 * DIST_GRID));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, reduce3_kernel<<<1, PCG_THREADS, 0, st>>>(work, work + DIST_GRID, work + 2 * DIST_GRID, DIST_GRID, out));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

}  // extern "C"
