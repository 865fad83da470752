// Two-level preconditioner for the Newton system: block-Jacobi plus a coarse correction on the rigid-body
// modes of vertex aggregates,
//
//     M^-1 = D^-1 + P (P^T A P)^-1 P^T ,     P_v = [ I | -[x_v - c_I]_x ]   (3 x 6 per vertex of aggregate I;
//                                                                               2 x 3 in 2D)
//
// (additive, symmetric positive definite, so plain PCG applies).  The reference solves the Newton system
// directly (solvers/newton.py:52, scipy spsolve); any preconditioner only changes how many CG iterations reach
// the same solution to the requested tolerance.  Block-Jacobi alone needs O(cells per side) iterations (709 at
// the 139^3-cell workload); the coarse space removes the smooth, nearly rigid error modes that cause it.
//
// Everything is deterministic: aggregates own fixed vertex lists, the coarse matrix is assembled by one CTA per
// coarse block over a fixed list of fine blocks with a fixed reduction tree, and the dense inverse is computed
// once per Newton iteration by cuSOLVER (potrf + potri on <= 6,000 unknowns: a plain library factorization off
// the hot path; the per-iteration work -- restriction, dense GEMV, prolongation -- is the kernels below).
#pragma once
#include <cusolverDn.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "capi_common.cuh"
#include "solver.cuh"

namespace skb {

template <int D>
struct CoarseDim {
  static constexpr int NC = (D == 3) ? 6 : 3;  // rigid modes per aggregate
};

struct CoarseView {
  int n_agg, nc;        // aggregates, coarse unknowns = NC * n_agg
  const int* agg;       // [n]
  const double* xrel;   // [n][D] vertex position relative to its aggregate's centre
  const int* vord;      // [n] vertices sorted by aggregate
  const int* aptr;      // [n_agg + 1]
  const double* Ainv;   // [nc][nc]
  double* rc;           // [nc]
  double* zc;           // [nc]
};

// P_v^T y (y a D-vector at vertex v): forces and the torque about the aggregate centre
template <int D>
__device__ __forceinline__ void coarse_Pt(const double* x, const double* y, double* out) {
  if (D == 3) {
    out[0] = y[0]; out[1] = y[1]; out[2] = y[2];
    out[3] = x[1] * y[2] - x[2] * y[1];
    out[4] = x[2] * y[0] - x[0] * y[2];
    out[5] = x[0] * y[1] - x[1] * y[0];
  } else {
    out[0] = y[0]; out[1] = y[1];
    out[2] = x[0] * y[1] - x[1] * y[0];
  }
}
// P_v c (c the NC coarse values of the vertex's aggregate): u + w x x
template <int D>
__device__ __forceinline__ void coarse_P(const double* x, const double* c, double* out) {
  if (D == 3) {
    out[0] = c[0] + (c[4] * x[2] - c[5] * x[1]);
    out[1] = c[1] + (c[5] * x[0] - c[3] * x[2]);
    out[2] = c[2] + (c[3] * x[1] - c[4] * x[0]);
  } else {
    out[0] = c[0] - c[2] * x[1];
    out[1] = c[1] + c[2] * x[0];
  }
}

// r_c[I] = sum_{v in I} P_v^T r_v : one CTA per aggregate, fixed vertex order and reduction tree
template <int D>
__global__ void coarse_restrict_kernel(CoarseView c, const double* r, const PcgScalars* sc) {
  constexpr int NC = CoarseDim<D>::NC;
  __shared__ double sh[32];
  if (sc && sc->done) return;
  const int I = blockIdx.x;
  double acc[NC];
#pragma unroll
  for (int a = 0; a < NC; ++a) acc[a] = 0.0;
  for (int k = c.aptr[I] + threadIdx.x; k < c.aptr[I + 1]; k += blockDim.x) {
    const int v = c.vord[k];
    double y[D], x[D], o[NC];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      y[i] = r[(size_t)v * D + i];
      x[i] = c.xrel[(size_t)v * D + i];
    }
    coarse_Pt<D>(x, y, o);
#pragma unroll
    for (int a = 0; a < NC; ++a) acc[a] += o[a];
  }
#pragma unroll
  for (int a = 0; a < NC; ++a) {
    const double s = block_reduce_sum(acc[a], sh);
    if (threadIdx.x == 0) c.rc[I * NC + a] = s;
  }
}

// z_c = Ainv r_c : one warp per row of the dense inverse
static __global__ void coarse_gemv_kernel(CoarseView c, const PcgScalars* sc) {
  if (sc && sc->done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= c.nc) return;
  const double* a = c.Ainv + (size_t)row * c.nc;
  double s = 0.0;
  for (int k = lane; k < c.nc; k += 32) s = fma(a[k], c.rc[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) c.zc[row] = s;
}

// z += P z_c ; partial sums of r.z (replace the block-Jacobi ones); optionally p = z (initialisation)
template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM)
coarse_add_kernel(CoarseView c, int v0, int v1, const double* r, double* z, double* pv, double* part_rz,
                  const PcgScalars* sc) {
  constexpr int NC = CoarseDim<D>::NC;
  __shared__ double sh[32];
  if (sc && sc->done) return;
  double rz = 0.0;
  for (int v = v0 + blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += gridDim.x * blockDim.x) {
    const int I = c.agg[v];
    double x[D], cc[NC], o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = c.xrel[(size_t)v * D + i];
#pragma unroll
    for (int a = 0; a < NC; ++a) cc[a] = c.zc[I * NC + a];
    coarse_P<D>(x, cc, o);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const size_t k = (size_t)v * D + i;
      const double zi = z[k] + o[i];
      z[k] = zi;
      if (pv) pv[k] = zi;
      rz = fma(r[k], zi, rz);
    }
  }
  rz = block_reduce_sum(rz, sh);
  if (threadIdx.x == 0) part_rz[blockIdx.x] = rz;
}

// A_c[I][J] = sum over the fine blocks (v in I, w in J) of P_v^T (A_vw + diag) P_w : one CTA per coarse block,
// fine blocks in a fixed (sorted) order, fixed reduction tree.  Fine block b lives in the scalar-CSR value array
// at  vals[bptr[v]*D*D + i*ncol + D*j + k],  j = b - bptr[v],  ncol = D * (bptr[v+1] - bptr[v]).
template <int D>
__global__ void coarse_assemble_kernel(PlanView p, const double* vals, const double* dadd, int n_agg, const int* agg,
                                       const double* xrel, const int* cb_ptr, const int* cb_I, const int* cb_J,
                                       const int* fb, double* Ac) {
  constexpr int NC = CoarseDim<D>::NC;
  __shared__ double sh[32];
  const int cb = blockIdx.x;
  double acc[NC * NC];
#pragma unroll
  for (int a = 0; a < NC * NC; ++a) acc[a] = 0.0;
  for (int k = cb_ptr[cb] + threadIdx.x; k < cb_ptr[cb + 1]; k += blockDim.x) {
    const int b = fb[k];
    const int v = p.brow[b], w = p.bcol[b];
    const int b0 = p.bptr[v];
    const int ncol = D * (p.bptr[v + 1] - b0);
    const double* base = vals + (size_t)b0 * (D * D) + (size_t)D * (b - b0);
    double A[D][D], xv[D], xw[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int kk = 0; kk < D; ++kk) A[i][kk] = base[(size_t)i * ncol + kk];
      if (v == w && dadd) A[i][i] += dadd[(size_t)v * D + i];
      xv[i] = xrel[(size_t)v * D + i];
      xw[i] = xrel[(size_t)w * D + i];
    }
    // T = A P_w  (D x NC): column a of P_w is P_w e_a
    double T[D][NC];
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      double e[NC], col[D];
#pragma unroll
      for (int q = 0; q < NC; ++q) e[q] = (q == a) ? 1.0 : 0.0;
      coarse_P<D>(xw, e, col);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < D; ++kk) s = fma(A[i][kk], col[kk], s);
        T[i][a] = s;
      }
    }
    // acc += P_v^T T
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      double y[D], o[NC];
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = T[i][a];
      coarse_Pt<D>(xv, y, o);
#pragma unroll
      for (int q = 0; q < NC; ++q) acc[q * NC + a] += o[q];
    }
  }
  const int I = cb_I[cb], J = cb_J[cb];
  const size_t nc = (size_t)NC * n_agg;
#pragma unroll
  for (int a = 0; a < NC * NC; ++a) {
    const double s = block_reduce_sum(acc[a], sh);
    if (threadIdx.x == 0) Ac[((size_t)I * NC + a / NC) * nc + (size_t)J * NC + (a % NC)] = s;
  }
}

// full symmetric matrix from the triangle potri leaves (row-major view: potri on the column-major "lower"
// triangle fills entries with column >= row of the row-major array)
static __global__ void coarse_symmetrize_kernel(int n, double* A) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int row = (int)(idx / n), col = (int)(idx - (size_t)row * n);
  if (row > col) A[idx] = A[(size_t)col * n + row];
}

struct CoarseKeyOf {
  const int* brow;
  const int* bcol;
  const int* agg;
  int n_agg, v0, v1;  // only fine blocks of the owned rows [v0, v1) take part (others sort to the end)
  __host__ __device__ uint64_t operator()(int b) const {
    const int v = brow[b];
    if (v < v0 || v >= v1) return ~0ull;
    return (uint64_t)agg[v] * (uint64_t)n_agg + (uint64_t)agg[bcol[b]];
  }
};
struct CoarseVertexKey {
  const int* agg;
  int n_agg, v0, v1;
  __host__ __device__ int operator()(int v) const { return (v < v0 || v >= v1) ? n_agg : agg[v]; }
};
struct CoarseKeyI {
  int n_agg;
  __host__ __device__ int operator()(uint64_t k) const { return (int)(k / (uint64_t)n_agg); }
};
struct CoarseKeyJ {
  int n_agg;
  __host__ __device__ int operator()(uint64_t k) const { return (int)(k % (uint64_t)n_agg); }
};

// plan-side data of the coarse space (built once by skb_pcg_set_coarse)
struct CoarseSpace {
  int n_agg = 0, n_cb = 0, v0 = 0, v1 = 0;
  dvec<int> agg, vord, aptr, cb_ptr, cb_I, cb_J, fb;
  dvec<double> xrel, Ac, rc, zc, work;
  dvec<int> info;
  cusolverDnHandle_t handle = nullptr;
  int lwork = 0;
  ~CoarseSpace() {
    if (handle) cusolverDnDestroy(handle);
  }
};


#define SKB_ENOTSPD (-5)  // internal: the callers fall back to block-Jacobi for this solve

#define SKB_CUSOLVER(call)                                                                       \
  do {                                                                                           \
    cusolverStatus_t _s = (call);                                                                \
    if (_s != CUSOLVER_STATUS_SUCCESS) return fail(SKB_ECUDA, "cuSOLVER call failed: " #call);   \
  } while (0)

// Ac (nc x nc, zeroed here) += P^T (A + diag) P over the fine blocks of the owned rows
template <int D>
inline int coarse_assemble_launch(skb_plan* pl, const double* vals, const double* dadd, double* Ac, cudaStream_t st) {
  CoarseSpace& c = *pl->coarse;
  const size_t nc = (size_t)CoarseDim<D>::NC * c.n_agg;
  const PlanView p = pl->view();
  SKB_CUDA(cudaMemsetAsync(Ac, 0, nc * nc * sizeof(double), st));
  if (c.n_cb > 0)
    SKB_LAUNCH(pl, SKB_K_OTHER, st,
               coarse_assemble_kernel<D><<<c.n_cb, 256, 0, st>>>(p, vals, dadd, c.n_agg, raw(c.agg), raw(c.xrel), raw(c.cb_ptr),
                                                                raw(c.cb_I), raw(c.cb_J), raw(c.fb), Ac));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// Ac := Ac^-1 (dense Cholesky, cuSOLVER potrf + potri, then the full symmetric matrix)
inline int coarse_invert(skb_plan* pl, double* Ac, int nc, cudaStream_t st) {
  CoarseSpace& c = *pl->coarse;
  if (!c.handle) SKB_CUSOLVER(cusolverDnCreate(&c.handle));
  SKB_CUSOLVER(cusolverDnSetStream(c.handle, st));
  int lw1 = 0, lw2 = 0;
  SKB_CUSOLVER(cusolverDnDpotrf_bufferSize(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, &lw1));
  SKB_CUSOLVER(cusolverDnDpotri_bufferSize(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, &lw2));
  const int lw = lw1 > lw2 ? lw1 : lw2;
  if ((int)c.work.size() < lw) c.work.resize(lw);
  if (c.info.size() < 2) c.info.resize(2);
  SKB_CUSOLVER(cusolverDnDpotrf(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, raw(c.work), lw, raw(c.info)));
  SKB_CUSOLVER(cusolverDnDpotri(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, raw(c.work), lw, raw(c.info) + 1));
  int hinfo[2] = {0, 0};
  SKB_CUDA(cudaMemcpyAsync(hinfo, raw(c.info), 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  SKB_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0] != 0 || hinfo[1] != 0)  // e.g. an aggregate of collinear vertices: its rotations are dependent
    return fail(SKB_ENOTSPD, "coarse matrix of the two-level preconditioner is not positive definite");
  coarse_symmetrize_kernel<<<(unsigned)(((size_t)nc * nc + 255) / 256), 256, 0, st>>>(nc, Ac);
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// unit columns c0 .. c0 + w of the identity, column-major nc x w
static __global__ void coarse_unit_columns_kernel(int nc, int c0, int w, double* X) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nc * w) return;
  const int col = (int)(idx / nc), row = (int)(idx - (size_t)col * nc);
  X[idx] = (row == c0 + col) ? 1.0 : 0.0;
}

// Cholesky factor of Ac in place (potrf) and the columns [c0, c0 + w) of its inverse into Xs (nc x w, column-major;
// potrs with unit right-hand sides): what ONE rank of a sharded solve needs of the dense inverse.  No potri.
inline int coarse_factor_slice(skb_plan* pl, double* Ac, int nc, int c0, int w, double* Xs, cudaStream_t st) {
  CoarseSpace& c = *pl->coarse;
  if (!c.handle) SKB_CUSOLVER(cusolverDnCreate(&c.handle));
  SKB_CUSOLVER(cusolverDnSetStream(c.handle, st));
  int lw = 0;
  SKB_CUSOLVER(cusolverDnDpotrf_bufferSize(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, &lw));
  if ((int)c.work.size() < lw) c.work.resize(lw);
  if (c.info.size() < 2) c.info.resize(2);
  SKB_CUSOLVER(cusolverDnDpotrf(c.handle, CUBLAS_FILL_MODE_LOWER, nc, Ac, nc, raw(c.work), lw, raw(c.info)));
  coarse_unit_columns_kernel<<<(unsigned)(((size_t)nc * w + 255) / 256), 256, 0, st>>>(nc, c0, w, Xs);
  SKB_CUDA(cudaGetLastError());
  SKB_CUSOLVER(cusolverDnDpotrs(c.handle, CUBLAS_FILL_MODE_LOWER, nc, w, Ac, nc, Xs, nc, raw(c.info) + 1));
  int hinfo[2] = {0, 0};
  SKB_CUDA(cudaMemcpyAsync(hinfo, raw(c.info), 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  SKB_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0] != 0 || hinfo[1] != 0)
    return fail(SKB_ENOTSPD, "coarse matrix of the two-level preconditioner is not positive definite");
  return SKB_OK;
}

}  // namespace skb
