// Linear solve and Newton-step kernels.
//
// The matrix lives in the plan's canonical scalar-CSR value layout; the kernels
// use the block structure (bptr, bcol) so that every dim x dim block costs one
// int32 column index:  bytes per block = dim*dim*8 + 4  (SURVEY §8d "PCG iteration").
//
// PCG iteration = 3 kernels, no host round trip:
//   pcg_spmv_dot : q = (A + diag) p,  partial sums of p.q per CTA
//   pcg_update   : alpha from the partials; x += alpha p; r -= alpha q; z = Dinv r;
//                  partial sums of r.z and r.r per CTA
//   pcg_direction: beta from the partials; p = z + beta p; CTA 0 publishes the scalars
// Reductions use a fixed grid and a fixed summation tree => bitwise reproducible.
#pragma once
#include "kernels.cuh"

namespace skb {

constexpr int PCG_THREADS = 256;
#ifndef SKB_PCG_CTAS_PER_SM
#define SKB_PCG_CTAS_PER_SM 8
#endif
constexpr int PCG_MAX_GRID = 2048;  // size of each per-CTA partial-sum array
constexpr int PCG_CTAS_PER_SM = SKB_PCG_CTAS_PER_SM;  // resident CTAs per SM the PCG kernels are compiled and launched for
#ifndef SKB_SPMV_GROUP
#define SKB_SPMV_GROUP 16
#endif
constexpr int SPMV_GROUP = SKB_SPMV_GROUP;  // lanes cooperating on one block row

// scalars kept on the device between the kernels of an iteration
struct PcgScalars {
  double rz;      // r.z of the current iterate
  double rr;      // r.r
  double bb;      // rhs.rhs
  double alpha;
  double beta;
  int done;       // set once rr <= rtol^2 * bb
  int iters;
};

#if defined(__CUDACC__)

// every CTA reduces the same n partials in the same order => same bits everywhere
__device__ __forceinline__ double reduce_partials(const double* part, int n, double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = block_reduce_sum(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = v;
  __syncthreads();
  v = bc;
  __syncthreads();
  return v;
}

// y = (A + diag(dadd)) x for the block rows [grid-stride]; returns per-thread partial of x.y
template <int D>
__device__ __forceinline__ double spmv_rows(const PlanView& p, const double* __restrict__ vals,
                                            const double* __restrict__ dadd, const double* __restrict__ x,
                                            double* __restrict__ y) {
  // all lanes of a warp run the same number of outer iterations (full-mask shuffles below)
  constexpr int GW = 32 / SPMV_GROUP;
  const int lane = threadIdx.x & (SPMV_GROUP - 1);
  const int gw = (threadIdx.x & 31) / SPMV_GROUP;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int v0 = warp * GW; v0 < p.n; v0 += nwarps * GW) {
    const int v = v0 + gw;
    const bool valid = v < p.n;
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
  return dot;
}

template <int D>
__global__ void spmv_kernel(PlanView p, const double* vals, const double* dadd, const double* x, double* y) {
  spmv_rows<D>(p, vals, dadd, x, y);
}

template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) pcg_spmv_dot_kernel(PlanView p, const double* vals, const double* dadd, const double* pvec,
                                    double* q, double* part_pq, const PcgScalars* sc) {
  __shared__ double sh[32];
  if (sc->done) return;
  double dot = spmv_rows<D>(p, vals, dadd, pvec, q);
  dot = block_reduce_sum(dot, sh);
  if (threadIdx.x == 0) part_pq[blockIdx.x] = dot;
}

// Dinv[v] = inverse of the diagonal block of (A + diag(dadd))
template <int D>
__global__ void block_jacobi_kernel(PlanView p, const double* vals, const double* dadd, double* dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.n) return;
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
  const double dt = det(A);
  Mat<D> c = cofactor(A);
  const double inv = 1.0 / dt;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int k = 0; k < D; ++k) dinv[(size_t)v * (D * D) + i * D + k] = c.m[k][i] * inv;  // adj = cof^T
}

template <int D>
__device__ __forceinline__ void apply_dinv(const double* dinv, int v, const double* r, double* z) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s = fma(dinv[(size_t)v * (D * D) + i * D + k], r[k], s);
    z[i] = s;
  }
}

// r = rhs (x0 = 0), z = Dinv r, p = z; partials of r.z and r.r
template <int D>
__global__ void pcg_init_kernel(int nb, const double* rhs, const double* dinv, double* x, double* r,
                                double* z, double* pv, double* part_rz, double* part_rr) {
  __shared__ double sh[32];
  double rz = 0.0, rr = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nb; v += gridDim.x * blockDim.x) {
    double rl[D], zl[D];
#pragma unroll
    for (int i = 0; i < D; ++i) rl[i] = rhs[(size_t)v * D + i];
    apply_dinv<D>(dinv, v, rl, zl);
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
    part_rz[blockIdx.x] = rz;
    part_rr[blockIdx.x] = rr;
  }
}

static __global__ void pcg_init_scalars_kernel(const double* part_rz, const double* part_rr, int nparts, double rtol,
                                        PcgScalars* sc) {
  __shared__ double sh[32];
  const double rz = reduce_partials(part_rz, nparts, sh);
  const double rr = reduce_partials(part_rr, nparts, sh);
  if (threadIdx.x == 0) {
    sc->rz = rz;
    sc->rr = rr;
    sc->bb = rr;
    sc->alpha = 0.0;
    sc->beta = 0.0;
    sc->iters = 0;
    sc->done = !(rr > 0.0) ? 1 : 0;
  }
}

template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) pcg_update_kernel(int nb, const double* dinv, const double* pv, const double* q, double* x,
                                  double* r, double* z, const double* part_pq, int nparts, double* part_rz,
                                  double* part_rr, const PcgScalars* sc) {
  __shared__ double sh[32];
  if (sc->done) return;
  const double pq = reduce_partials(part_pq, nparts, sh);
  const double alpha = sc->rz / pq;
  double rz = 0.0, rr = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nb; v += gridDim.x * blockDim.x) {
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
    part_rz[blockIdx.x] = rz;
    part_rr[blockIdx.x] = rr;
  }
}

template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM) pcg_direction_kernel(int nb, const double* z, double* pv, const double* part_rz,
                                     const double* part_rr, int nparts, double rtol, PcgScalars* sc,
                                     PcgScalars* sc_next) {
  __shared__ double sh[32];
  if (sc->done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *sc_next = *sc;
    return;
  }
  const double rz_new = reduce_partials(part_rz, nparts, sh);
  const double rr_new = reduce_partials(part_rr, nparts, sh);
  const double beta = rz_new / sc->rz;
  const int n = nb * D;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    pv[k] = fma(beta, pv[k], z[k]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    PcgScalars s = *sc;
    s.rz = rz_new;
    s.rr = rr_new;
    s.beta = beta;
    s.iters = sc->iters + 1;
    s.done = (!(rr_new > rtol * rtol * s.bb)) ? 1 : 0;  // also stops on NaN
    *sc_next = s;
  }
}

// ---- Newton-step vector kernels -------------------------------------------
// total gradient  g += -f_ext + kin_scale * mass * (x - x_tilde) + pin_k * (x - pin_t);
// rhs = -g;  partial sums of nothing (kept simple)
static __global__ void newton_gradient_kernel(int n, const double* x, const double* fext, const double* mass,
                                       const double* xt, double kin, const double* pin_k, const double* pin_t,
                                       double* g, double* rhs, double* dadd) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double gv = g[k];
    double da = 0.0;
    if (fext) gv -= fext[k];
    if (mass && xt) {
      gv = fma(kin * mass[k], x[k] - xt[k], gv);
      da += kin * mass[k];
    }
    if (pin_k) {
      gv = fma(pin_k[k], x[k] - pin_t[k], gv);
      da += pin_k[k];
    }
    g[k] = gv;
    rhs[k] = -gv;
    dadd[k] = da;
  }
}

// partial sums of the non-elastic energy terms at x (+ optional step: x + s*dx) and of g.dx
static __global__ void newton_energy_terms_kernel(int n, const double* x, const double* dx, double s, const double* fext,
                                           const double* mass, const double* xt, double kin, const double* pin_k,
                                           const double* pin_t, const double* g, double* xtrial, double* part_e,
                                           double* part_gdx, double* part_dx2) {
  __shared__ double sh[32];
  double e = 0.0, gd = 0.0, d2 = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double d = dx ? dx[k] : 0.0;
    const double xv = fma(s, d, x[k]);
    if (xtrial) xtrial[k] = xv;
    if (fext) e = fma(-fext[k], xv, e);
    if (mass && xt) {
      const double dd = xv - xt[k];
      e = fma(0.5 * kin * mass[k] * dd, dd, e);
    }
    if (pin_k) {
      const double dd = xv - pin_t[k];
      e = fma(0.5 * pin_k[k] * dd, dd, e);
    }
    if (g) gd = fma(g[k], d, gd);
    d2 = fma(d, d, d2);
  }
  e = block_reduce_sum(e, sh);
  gd = block_reduce_sum(gd, sh);
  d2 = block_reduce_sum(d2, sh);
  if (threadIdx.x == 0) {
    part_e[blockIdx.x] = e;
    part_gdx[blockIdx.x] = gd;
    part_dx2[blockIdx.x] = d2;
  }
}

// ---- plane contact springs (energies/contact_springs_plane.py:245-388) --------------------------------------
// E = k/2 sum_{v under the plane} m_v (n . (x_v - p))^2 ; per contacting vertex: gradient k m_v off n, Hessian block
// k m_v n n^T.  One thread per vertex; any of the outputs may be null.  `vals` (with the plan view) receives the
// Hessian blocks straight in the diagonal blocks of the scalar-CSR value array.
// kind 1: sphere (energies/contact_springs_sphere.py:242-360): vertices with |x_v - p| < r, per-vertex outward normal
// n_v = (x_v - p)/|x_v - p| held fixed in the derivatives, offset |x_v - p| - r; same energy / gradient / block forms.
struct ContactPlaneArgs {
  double k;
  double p[3];
  double n[3];
  const double* w;  // per-vertex weights m_v or nullptr (1)
  int kind;         // 0 plane, 1 sphere
  double r;         // sphere radius
};

// vertices [vbeg, nv): a shard passes its owned range (every vertex is then handled by exactly one rank)
template <int D>
__global__ void contact_plane_kernel(int nv, const double* x, ContactPlaneArgs c, double* g_add, double* blocks,
                                     const PlanView* pv, double* vals, double* part_e, int* under, int vbeg = 0) {
  __shared__ double sh[32];
  double e = 0.0;
  for (int v = vbeg + blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    double off = 0.0;
    double nv_[D];
    if (c.kind == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        nv_[i] = c.n[i];
        off = fma(c.n[i], x[(size_t)v * D + i] - c.p[i], off);
      }
    } else {
      double len2 = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        nv_[i] = x[(size_t)v * D + i] - c.p[i];
        len2 = fma(nv_[i], nv_[i], len2);
      }
      const double len = sqrt(len2);
#pragma unroll
      for (int i = 0; i < D; ++i) nv_[i] /= len;
      off = len - c.r;
    }
    const bool in = off < 0.0;
    if (under) under[v] = in ? 1 : 0;
    const double m = c.w ? c.w[v] : 1.0;
    const double km = in ? c.k * m : 0.0;
    e = fma(0.5 * km * off, off, e);
    if (g_add) {
#pragma unroll
      for (int i = 0; i < D; ++i) g_add[(size_t)v * D + i] += km * off * nv_[i];
    }
    if (blocks) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) blocks[((size_t)v * D + i) * D + j] = km * nv_[i] * nv_[j];
    }
    if (vals && in) {
      // diagonal block (v, v) of the block row: binary search of v among the sorted block columns
      const int b0 = pv->bptr[v], b1 = pv->bptr[v + 1];
      int lo = b0, hi = b1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pv->bcol[mid] < v) lo = mid + 1; else hi = mid;
      }
      const int ncol = D * (b1 - b0);
      double* base = vals + (size_t)b0 * (D * D) + (size_t)D * (lo - b0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) base[(size_t)i * ncol + j] += km * nv_[i] * nv_[j];
    }
  }
  if (part_e) {
    e = block_reduce_sum(e, sh);
    if (threadIdx.x == 0) part_e[blockIdx.x] = e;
  }
}

// ---- general sparse quadratic term 1/2 x^T Q x + b^T x (energies/quadratic.py:15-70) -----------------------
// One thread per row of the CSR matrix Q (rows of penalty / regulariser matrices are short): y_i = sum_j Q_ij x_j,
// g_add[i] += y_i + b_i (quadratic_gradient, Q symmetric as the reference assumes), per-CTA partial of
// x_i (y_i / 2 + b_i) (quadratic_energy).  b, g_add and part_e may be null.  Fixed order => deterministic.
static __global__ void quad_term_kernel(int n, const int* __restrict__ qptr, const int* __restrict__ qcol,
                                        const double* __restrict__ qv, const double* __restrict__ b,
                                        const double* __restrict__ x, double* g_add, double* part_e) {
  __shared__ double sh[32];
  double e = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    double y = 0.0;
    const int k1 = qptr[r + 1];
    for (int k = qptr[r]; k < k1; ++k) y = fma(qv[k], x[qcol[k]], y);
    const double bv = b ? b[r] : 0.0;
    if (g_add) g_add[r] += y + bv;
    e = fma(x[r], fma(0.5, y, bv), e);
  }
  if (part_e) {
    e = block_reduce_sum(e, sh);
    if (threadIdx.x == 0) part_e[blockIdx.x] = e;
  }
}

// where every entry of Q sits in the plan's CSR values (quadratic_hessian = Q is then one scatter-add per Newton
// iteration); *bad counts the entries outside the pattern
template <int D>
__global__ void quad_positions_kernel(PlanView p, int n, const int* qptr, const int* qcol, int* pos, int* bad) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int k = qptr[r]; k < qptr[r + 1]; ++k) {
    const int q = csr_value_position<D>(p.bptr, p.bcol, r, qcol[k]);
    pos[k] = q < 0 ? 0 : q;
    if (q < 0) atomicAdd(bad, 1);
  }
}

static __global__ void add_at_kernel(double* dst, const int* __restrict__ pos, int n, const double* __restrict__ src) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst[pos[k]] += src[k];
}

// out[0..2] = sums of three partial arrays (single CTA)
static __global__ void reduce3_kernel(const double* a, const double* b, const double* c, int n, double* out) {
  __shared__ double sh[32];
  const double va = reduce_partials(a, n, sh);
  const double vb = reduce_partials(b, n, sh);
  const double vc = reduce_partials(c, n, sh);
  if (threadIdx.x == 0) {
    out[0] = va;
    out[1] = vb;
    out[2] = vc;
  }
}
// ---- generic scalar-CSR operators (any sparse SPD matrix handed to newton_solver) ----------
// y = A x, SPMV_GROUP lanes per row; returns the per-thread partial of x.y
__device__ __forceinline__ double csr_spmv_rows(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                const double* __restrict__ vals, const double* __restrict__ x,
                                                double* __restrict__ y) {
  constexpr int GW = 32 / SPMV_GROUP;
  const int lane = threadIdx.x & (SPMV_GROUP - 1);
  const int gw = (threadIdx.x & 31) / SPMV_GROUP;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int r0 = warp * GW; r0 < n; r0 += nwarps * GW) {
    const int r = r0 + gw;
    const bool valid = r < n;
    const int k0 = valid ? indptr[r] : 0, k1 = valid ? indptr[r + 1] : 0;
    double acc = 0.0;
    for (int k = k0 + lane; k < k1; k += SPMV_GROUP) acc = fma(vals[k], x[indices[k]], acc);
#pragma unroll
    for (int o = SPMV_GROUP / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, SPMV_GROUP);
    if (lane == 0 && valid) {
      y[r] = acc;
      dot = fma(x[r], acc, dot);
    }
  }
  return dot;
}

static __global__ void csr_pcg_spmv_dot_kernel(int n, const int* indptr, const int* indices, const double* vals,
                                               const double* pvec, double* q, double* part_pq, const PcgScalars* sc) {
  __shared__ double sh[32];
  if (sc->done) return;
  double dot = csr_spmv_rows(n, indptr, indices, vals, pvec, q);
  dot = block_reduce_sum(dot, sh);
  if (threadIdx.x == 0) part_pq[blockIdx.x] = dot;
}

// inverse of the D x D diagonal blocks of a scalar CSR matrix (n = nb * D rows)
template <int D>
__global__ void csr_block_jacobi_kernel(int nb, const int* indptr, const int* indices, const double* vals, double* dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nb) return;
  Mat<D> A;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int k = 0; k < D; ++k) A.m[i][k] = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int r = v * D + i;
    for (int k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int c = indices[k] - v * D;
      if (c >= 0 && c < D) {
#pragma unroll
        for (int kk = 0; kk < D; ++kk)
          if (kk == c) A.m[i][kk] += vals[k];
      }
    }
  }
  if (D == 1) {
    dinv[v] = 1.0 / A.m[0][0];
    return;
  }
  const double dt = det(A);
  Mat<D> c = cofactor(A);
  const double inv = 1.0 / dt;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int k = 0; k < D; ++k) dinv[(size_t)v * (D * D) + i * D + k] = c.m[k][i] * inv;
}

// dense LU with partial pivoting, one CTA, A (n x n row-major) and b overwritten; x = solution
static __global__ void dense_solve_kernel(int n, double* A, double* b, double* x, int* status) {
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ int piv;
  for (int k = 0; k < n; ++k) {
    // pivot search
    double best = -1.0;
    int bi = k;
    for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
      const double a = fabs(A[(size_t)i * n + k]);
      if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sval[threadIdx.x >> 5] = best; sidx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double bb = sval[0];
      int ii = sidx[0];
      for (int w = 1; w < (blockDim.x + 31) / 32; ++w)
        if (sval[w] > bb || (sval[w] == bb && sidx[w] < ii)) { bb = sval[w]; ii = sidx[w]; }
      piv = ii;
      if (!(bb > 0.0)) *status = 1;
    }
    __syncthreads();
    const int p = piv;
    if (p != k) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double tmp = A[(size_t)k * n + j];
        A[(size_t)k * n + j] = A[(size_t)p * n + j];
        A[(size_t)p * n + j] = tmp;
      }
      if (threadIdx.x == 0) { const double tmp = b[k]; b[k] = b[p]; b[p] = tmp; }
    }
    __syncthreads();
    const double inv = 1.0 / A[(size_t)k * n + k];
    for (int i = k + 1 + (threadIdx.x >> 5); i < n; i += (blockDim.x >> 5)) {
      const double f = A[(size_t)i * n + k] * inv;
      for (int j = k + 1 + (threadIdx.x & 31); j < n; j += 32) A[(size_t)i * n + j] = fma(-f, A[(size_t)k * n + j], A[(size_t)i * n + j]);
      if ((threadIdx.x & 31) == 0) b[i] = fma(-f, b[k], b[i]);
    }
    __syncthreads();
  }
  // back substitution
  for (int k = n - 1; k >= 0; --k) {
    if (threadIdx.x == 0) x[k] = b[k] / A[(size_t)k * n + k];
    __syncthreads();
    const double xk = x[k];
    for (int i = threadIdx.x; i < k; i += blockDim.x) b[i] = fma(-A[(size_t)i * n + k], xk, b[i]);
    __syncthreads();
  }
}
#endif  // __CUDACC__


}  // namespace skb
