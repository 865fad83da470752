// C ABI, part 7: the distributed Newton-system solve as ONE collective per iteration.
//
// The reference solves H dx = -g directly on one CPU (solvers/newton.py:52, scipy spsolve); on an element-sharded
// mesh the solve is a preconditioned CG whose cost at 8 GPUs is not the 0.07 ms SpMV but the
// collectives and launches around it: the textbook loop of capi_nccl.cu needs three all-reduces (p.q; the restricted
// residual of the two-level preconditioner; r.z and r.r), a halo exchange and ~16 launches per iteration.
//
// This file is the single-reduction form (Chronopoulos & Gear 1989; prototype and proof of equivalence in
// oracle/elasticity.py::block_jacobi_cg_single_reduction):
//
//     u = M^-1 r,  w = A u,   gamma = r.u,  delta = w.u,  rr = r.r             <- ONE all-reduce
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p,  s = w + beta s (= A p),  x += alpha p,  r -= alpha s
//
// and the coarse correction of the two-level preconditioner M^-1 = D^-1 + P Ac^-1 P^T (coarse.cuh) rides in the same
// all-reduce: the restricted residual obeys the same recurrence as r,
//
//     P^T r_{k+1} = P^T r_k - alpha (P^T w_k + beta P^T s_{k-1}),
//
// so every rank adds the restriction of ITS rows of w to the reduction buffer and no second collective is needed.
// Per iteration: 7 kernels, one grouped ncclSend/ncclRecv (halo of u) and one ncclAllReduce of 4 + 6 n_agg doubles;
// chunks of iterations are captured once into a CUDA graph (kernels and NCCL calls) and replayed, the host reads the
// device-side convergence flag between chunks.  All reductions have a fixed shape: the result is reproducible for
// a fixed number of ranks.
#include "nccl_api.cuh"
#include "solver.cuh"
#include "coarse.cuh"

#include <chrono>

using namespace skb;

namespace skb {

constexpr int PCG2_GRID = 148 * PCG_CTAS_PER_SM;

struct Pcg2Scalars {
  double gamma_old, alpha_old, alpha, beta, bb, rr, rtol2;
  int stage;   // 0: bootstrap pass (alpha = beta = 0: computes u_0, w_0 and the first reduction), then counts up
  int done;    // 1: converged (rr <= rtol^2 bb), 2: breakdown (delta <= 0 or NaN), 3: zero right-hand side
  int iters;   // updates of x applied
  int pad;
  unsigned long long seq;   // peer-memory transport: passes issued since the plan's first solve (never reset; identical
                            // on every rank because every rank runs the same passes) -- the value the flags carry
};

// ---- peer-memory transport (NVLink loads/stores into the neighbours' HBM through CUDA IPC mappings) -------------------
// Every rank owns one slab, mapped by all the others:
//   G      double [2][world][PEER_RED_CAP]   reduction partials: rank q writes its 4 + nc values of pass j into slot
//                                            [j & 1][q] of EVERY rank's slab, so each rank sums the world partials
//                                            itself in rank order (same bits everywhere, no collective call)
//   fred   u64 [world]                       fred[q] = seq of the last pass whose partials rank q delivered
//   fhalo  u64 [world]                       fhalo[q] = seq of the last pass whose halo values rank q delivered
//   rbuf   double [nr]                       halo values of u, written by the owners of those vertices
constexpr int PEER_MAXW = 16;
constexpr int PEER_RED_CAP = 4 + 6 * 2048 + 4;   // 12296 doubles: 4 scalars + the largest coarse space, 32-byte multiple
constexpr long long PEER_SPIN_LIMIT = 6000000000ll;   // clock64 ticks (~3 s) before a wait gives up and flags an error

struct PeerSlabLayout {
  size_t off_G, off_fred, off_fhalo, off_zc, off_fzc, off_rbuf, bytes;
  __host__ __device__ static PeerSlabLayout make(int world, int64_t nr) {
    PeerSlabLayout L;
    L.off_G = 0;
    L.off_fred = (size_t)2 * world * PEER_RED_CAP * sizeof(double);
    L.off_fhalo = L.off_fred + (size_t)PEER_MAXW * sizeof(unsigned long long);
    L.off_zc = L.off_fhalo + (size_t)PEER_MAXW * sizeof(unsigned long long) + 256;   // zc: the coarse solution, every rank
    L.off_fzc = L.off_zc + (size_t)PEER_RED_CAP * sizeof(double);                    // writes its slice into every slab
    L.off_rbuf = L.off_fzc + (size_t)PEER_MAXW * sizeof(unsigned long long) + 256;
    L.bytes = L.off_rbuf + (size_t)(nr > 0 ? nr : 1) * sizeof(double) + 256;
    return L;
  }
};

struct PeerRedTable {      // kernel argument: where my partials go on every rank (myself included)
  int world, me;
  double* G[PEER_MAXW];                 // base of rank q's G array
  unsigned long long* fred[PEER_MAXW];  // rank q's fred array
};
struct PeerZcTable {       // kernel argument: where my slice of the coarse solution goes on every rank
  int world, me;
  double* zc[PEER_MAXW];
  unsigned long long* fzc[PEER_MAXW];
};
struct PeerHaloTable {     // kernel argument: where my halo values go
  int np;
  int64_t soff[9];                      // my send list is the concatenation of the peers' lists
  double* rbuf[8];                      // peer's receive area for me
  unsigned long long* fhalo[8];         // peer's flag for me
  int peer[8];                          // peer ranks (the flags I wait for are my own fhalo[peer])
};

__device__ __forceinline__ bool peer_wait(const volatile unsigned long long* flag, unsigned long long want) {
  const long long t0 = clock64();
  while (*flag < want) {
    if (clock64() - t0 > PEER_SPIN_LIMIT) return false;
    __nanosleep(20);
  }
  return true;
}

// Waits until the first n flags (n <= 32) reach this pass's sequence number.  ONE small CTA polls: when every CTA of a
// large kernel polled the same flag line, the reads saturated that L2 slice and delayed the very store they were waiting
// for (8 GPUs: +15 us per synchronisation point); the consumer kernel that follows in the stream needs no wait of its own.
static __global__ void pcg3_wait_kernel(Pcg2Scalars* sc, const volatile unsigned long long* flags, int n, int use_peer_index,
                                        PeerHaloTable t) {
  if (sc->done) return;
  const unsigned long long seq = sc->seq;
  if ((int)threadIdx.x < n) {
    const volatile unsigned long long* f = use_peer_index ? flags + t.peer[threadIdx.x] : flags + threadIdx.x;
    if (!peer_wait(f, seq)) sc->done = 5;
  }
  __threadfence_system();
}

// halo of u: my owned values the neighbours need go straight into their receive areas; the last CTA raises the flags
static __global__ void pcg3_pack_push_kernel(const Pcg2Scalars* sc, const double* __restrict__ v,
                                             const int32_t* __restrict__ idx, PeerHaloTable t, unsigned int* counter) {
  if (sc->done) return;
  const int64_t n = t.soff[t.np];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int pi = 0;
    while (pi + 1 < t.np && i >= t.soff[pi + 1]) ++pi;
    t.rbuf[pi][i - t.soff[pi]] = v[idx[i]];
  }
  __threadfence_system();
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0;
    __threadfence_system();
    if ((int)threadIdx.x < t.np) *(volatile unsigned long long*)t.fhalo[threadIdx.x] = sc->seq;
  }
}

// copies the neighbours' halo values of this pass (already waited for by pcg3_wait_kernel) into the non-owned entries of u
static __global__ void pcg3_unpack_kernel(const Pcg2Scalars* sc, double* __restrict__ v, const int32_t* __restrict__ idx,
                                          int64_t n, const double* rbuf) {
  if (sc->done) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[idx[i]] = __ldcg(rbuf + i);
}

// ---- iteration kernels ------------------------------------------------------------------------------------------
// scalars of the iteration from the reduced sums red = [gamma, delta, rr, -, P^T w ...], then the coarse recurrences
//   rcs = P^T w + beta rcs,   rc -= alpha rcs      (one CTA; nc <= 12288 values)
// With the peer-memory transport (G != NULL) the kernel first waits for the partials of the previous pass from every
// rank and sums them into `red` in rank order (what the all-reduce does for the NCCL transport).
static __global__ void pcg2_scalars_kernel(Pcg2Scalars* sc, double* red, int nc, double* rc, double* rcs, int world,
                                           const double* G, const volatile unsigned long long* fred) {
  __shared__ double sab[2];
  __shared__ int sdone;
  __shared__ int ok;
  if (G && !sc->done && sc->stage > 0) {      // bootstrap pass: nothing was reduced yet (red is zero)
    const unsigned long long seq = sc->seq;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < world && !peer_wait(fred + threadIdx.x, seq)) ok = 0;
    __syncthreads();
    if (!ok) {
      if (threadIdx.x == 0) sc->done = 5;
    } else {
      __threadfence_system();
      const double* Gp = G + (size_t)(seq & 1ull) * world * PEER_RED_CAP;
      for (int k = threadIdx.x; k < 4 + nc; k += blockDim.x) {
        double acc = 0.0;
        for (int q = 0; q < world; ++q) acc += __ldcg(Gp + (size_t)q * PEER_RED_CAP + k);
        red[k] = acc;
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    Pcg2Scalars s = *sc;
    double alpha = 0.0, beta = 0.0;
    if (s.done == 5) {          // a peer-memory wait timed out: stop here, the host reports it
      sdone = 5;
    } else if (!s.done && s.stage > 0) {
      const double gamma = red[0], delta = red[1], rr = red[2];
      if (s.stage == 1) s.bb = rr;          // x_0 = 0: r_0 = b
      s.rr = rr;
      if (!(s.bb > 0.0)) {
        s.done = 3;
      } else if (!(rr > s.rtol2 * s.bb)) {
        s.done = 1;
      } else {
        beta = (s.stage == 1) ? 0.0 : gamma / s.gamma_old;
        const double den = (s.stage == 1) ? delta : delta - beta * gamma / s.alpha_old;
        if (!(den > 0.0) || !(gamma > 0.0)) {
          s.done = 2;
          beta = 0.0;
        } else {
          alpha = gamma / den;
          s.gamma_old = gamma;
          s.alpha_old = alpha;
          s.iters += 1;
        }
      }
    }
    s.alpha = alpha;
    s.beta = beta;
    if (!s.done) {
      s.stage += 1;
      s.seq += 1;      // the pushes of this pass carry this value, the waits of this pass and of the next gather expect it
    }
    *sc = s;
    sab[0] = alpha;
    sab[1] = beta;
    sdone = s.done;
  }
  __syncthreads();
  if (sdone) return;
  const double alpha = sab[0], beta = sab[1];
  for (int k = threadIdx.x; k < nc; k += blockDim.x) {
    const double t = fma(beta, rcs[k], red[4 + k]);
    rcs[k] = t;
    rc[k] = fma(-alpha, t, rc[k]);
  }
}

// zc = Ainv rc : one warp per row of the dense inverse
static __global__ void pcg2_gemv_kernel(const Pcg2Scalars* sc, int nc, const double* __restrict__ Ainv,
                                        const double* __restrict__ rc, double* __restrict__ zc) {
  if (sc->done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nc) return;
  const double* a = Ainv + (size_t)row * nc;
  // four independent 256-byte requests per warp in flight (the kernel streams nc^2 doubles once: bandwidth-bound)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = lane;
  for (; k + 96 < nc; k += 128) {
    const double a0 = a[k], a1 = a[k + 32], a2 = a[k + 64], a3 = a[k + 96];
    s0 = fma(a0, rc[k], s0);
    s1 = fma(a1, rc[k + 32], s1);
    s2 = fma(a2, rc[k + 64], s2);
    s3 = fma(a3, rc[k + 96], s3);
  }
  for (; k < nc; k += 32) s0 = fma(a[k], rc[k], s0);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) zc[row] = s;
}

// Peer-memory transport: every rank holds only ITS columns [c0, c0 + w) of the inverse of the coarse matrix (Xs, nc x w
// column-major, from potrf + potrs with unit right-hand sides) and computes that slice of zc = Ainv rc -- one warp per
// column -- storing it into the zc area of EVERY rank's slab; the last CTA raises this rank's flag on every rank.  The
// dense apply per iteration shrinks by the rank count, and so does the O(nc^3) inverse (potrf only, no potri).
static __global__ void pcg3_gemv_push_kernel(const Pcg2Scalars* sc, int nc, int c0, int w, const double* __restrict__ Xs,
                                             const double* __restrict__ rc, PeerZcTable t, unsigned int* counter) {
  __shared__ int last;
  if (sc->done) return;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (col < w) {
    const double* a = Xs + (size_t)col * nc;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = lane;
    for (; k + 96 < nc; k += 128) {
      const double a0 = a[k], a1 = a[k + 32], a2 = a[k + 64], a3 = a[k + 96];
      s0 = fma(a0, rc[k], s0);
      s1 = fma(a1, rc[k + 32], s1);
      s2 = fma(a2, rc[k + 64], s2);
      s3 = fma(a3, rc[k + 96], s3);
    }
    for (; k < nc; k += 32) s0 = fma(a[k], rc[k], s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0)
      for (int q = 0; q < t.world; ++q) t.zc[q][c0 + col] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0;
    __threadfence_system();
    if ((int)threadIdx.x < t.world) *(volatile unsigned long long*)(t.fzc[threadIdx.x] + t.me) = sc->seq;
  }
}

// p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = Dinv r [+ P zc]      (owned vertices)
// and the per-CTA partial sums of r.u and r.r of the NEW r, u (part[0..grid), part[2 grid..3 grid); w.u comes from the SpMV)
template <int D, bool COARSE>
__global__ void __launch_bounds__(PCG_THREADS, 5)
pcg2_update_kernel(const Pcg2Scalars* sc, int v0, int v1, const double* __restrict__ dinv, const int* __restrict__ agg,
                   const double* __restrict__ xrel, const double* __restrict__ zc, const double* __restrict__ w,
                   double* __restrict__ u, double* __restrict__ p, double* __restrict__ s, double* __restrict__ x,
                   double* __restrict__ r, double* part, int zc_world, const volatile unsigned long long* fzc,
                   Pcg2Scalars* sc_w) {
  __shared__ double sh[32];
  if (sc->done) return;
  constexpr int NC = CoarseDim<D>::NC;
  (void)zc_world; (void)fzc; (void)sc_w;   // the slices of zc were waited for by pcg3_wait_kernel
  const double alpha = sc->alpha, beta = sc->beta;
  double ru = 0.0, rr = 0.0;
  for (int v = v0 + blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += gridDim.x * blockDim.x) {
    double rl[D], ul[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const size_t k = (size_t)v * D + i;
      const double pn = fma(beta, p[k], u[k]);
      const double sn = fma(beta, s[k], w[k]);
      p[k] = pn;
      s[k] = sn;
      x[k] = fma(alpha, pn, x[k]);
      rl[i] = fma(-alpha, sn, r[k]);
      r[k] = rl[i];
    }
    apply_dinv<D>(dinv, v, rl, ul);
    if (COARSE) {
      const int I = agg[v];
      double xr[D], cc[NC], o[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xr[i] = xrel[(size_t)v * D + i];
#pragma unroll
      for (int a = 0; a < NC; ++a) cc[a] = __ldcg(zc + I * NC + a);
      coarse_P<D>(xr, cc, o);
#pragma unroll
      for (int i = 0; i < D; ++i) ul[i] += o[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      u[(size_t)v * D + i] = ul[i];
      ru = fma(rl[i], ul[i], ru);
      rr = fma(rl[i], rl[i], rr);
    }
  }
  ru = block_reduce_sum(ru, sh);
  rr = block_reduce_sum(rr, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = ru;
    part[2 * gridDim.x + blockIdx.x] = rr;
  }
}

static __global__ void pcg2_pack_kernel(const Pcg2Scalars* sc, const double* __restrict__ v, const int32_t* __restrict__ idx,
                                        int64_t n, double* __restrict__ buf) {
  if (sc->done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}
static __global__ void pcg2_unpack_kernel(const Pcg2Scalars* sc, double* __restrict__ v, const int32_t* __restrict__ idx,
                                          int64_t n, const double* __restrict__ buf) {
  if (sc->done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = buf[i];
}

// w = (A + diag) u on the owned block rows; per-CTA partial sums of w.u (part[grid..2 grid))
template <int D>
__global__ void __launch_bounds__(PCG_THREADS, PCG_CTAS_PER_SM)
pcg2_spmv_kernel(const Pcg2Scalars* sc, PlanView p, const double* __restrict__ vals, const double* __restrict__ dadd,
                 const double* __restrict__ u, double* __restrict__ w, int v0, int v1, double* part) {
  __shared__ double sh[32];
  if (sc->done) return;
  constexpr int GW = 32 / SPMV_GROUP;
  const int lane = threadIdx.x & (SPMV_GROUP - 1);
  const int gw = (threadIdx.x & 31) / SPMV_GROUP;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double wu = 0.0;
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
      const double xv = u[(size_t)p.bcol[b0 + j] * D + k];
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
        const size_t k = (size_t)v * D + i;
        const double ui = u[k];
        double wi = acc[i];
        if (dadd) wi = fma(dadd[k], ui, wi);
        w[k] = wi;
        wu = fma(wi, ui, wu);
      }
    }
  }
  wu = block_reduce_sum(wu, sh);
  if (threadIdx.x == 0) part[gridDim.x + blockIdx.x] = wu;
}

// CTAs [0, n_agg): red[4 + I*NC + a] = sum over the OWNED vertices of aggregate I of (P_v^T w_v)_a  (fixed order);
// the last CTA: red[0..2] = the three dot products, summed over the per-CTA partials in a fixed order
// PUSH: instead of leaving the sums in `red` for an ncclAllReduce, every CTA stores its values into this rank's slot of
// EVERY rank's gather buffer (NVLink stores), and the last CTA to finish raises this rank's flag on every rank.
template <int D, bool PUSH>
__global__ void pcg2_restrict_kernel(const Pcg2Scalars* sc, int n_agg, const int* __restrict__ vord,
                                     const int* __restrict__ aptr, const double* __restrict__ xrel,
                                     const double* __restrict__ w, const double* part, int npart, double* red,
                                     PeerRedTable pt, unsigned int* counter) {
  constexpr int NC = CoarseDim<D>::NC;
  __shared__ double sh[32];
  __shared__ int last;
  if (sc->done) return;
  const unsigned long long seq = PUSH ? sc->seq : 0ull;
  const size_t slot = PUSH ? ((size_t)(seq & 1ull) * pt.world + pt.me) * PEER_RED_CAP : 0;
  auto publish = [&](int k, double v) {
    if (PUSH) {
      for (int q = 0; q < pt.world; ++q) pt.G[q][slot + k] = v;
    } else {
      red[k] = v;
    }
  };
  auto finish = [&]() {
    if (!PUSH) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
      if (threadIdx.x == 0) *counter = 0;
      __threadfence_system();
      if ((int)threadIdx.x < pt.world) *(volatile unsigned long long*)(pt.fred[threadIdx.x] + pt.me) = seq;
    }
  };
  const int I = blockIdx.x;
  if (I == n_agg) {
    for (int q = 0; q < 3; ++q) {
      const double v = reduce_partials(part + (size_t)q * npart, npart, sh);
      if (threadIdx.x == 0) publish(q, v);
    }
    if (threadIdx.x == 0) publish(3, 0.0);
    finish();
    return;
  }
  double acc[NC];
#pragma unroll
  for (int a = 0; a < NC; ++a) acc[a] = 0.0;
  for (int k = aptr[I] + threadIdx.x; k < aptr[I + 1]; k += blockDim.x) {
    const int v = vord[k];
    double y[D], x[D], o[NC];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      y[i] = w[(size_t)v * D + i];
      x[i] = xrel[(size_t)v * D + i];
    }
    coarse_Pt<D>(x, y, o);
#pragma unroll
    for (int a = 0; a < NC; ++a) acc[a] += o[a];
  }
#pragma unroll
  for (int a = 0; a < NC; ++a) {
    const double s = block_reduce_sum(acc[a], sh);
    if (threadIdx.x == 0) publish(4 + I * NC + a, s);
  }
  finish();
}

// Dinv of the owned diagonal blocks of (A + diag)
template <int D>
__global__ void pcg2_dinv_kernel(PlanView p, const double* vals, const double* dadd, int v0, int v1, double* dinv) {
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
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) dinv[(size_t)v * (D * D) + i * D + k] = c.m[k][i] * inv;
  }
}

#if defined(SKB_HAVE_NCCL_H)
struct Pcg2State {
  int64_t nd = 0;
  int nc = 0;
  dvec<double> r, u, w, p, s, dinv, red, rc, rcs, zc, Ainv, part, sbuf, rbuf;
  dvec<int32_t> sidx, ridx;
  std::vector<int64_t> soff, roff;   // per peer: offsets into sbuf / rbuf (n_peers + 1)
  dvec<Pcg2Scalars> sc;
  // graph of one chunk of iterations, valid while the arguments it baked in stay the same
  cudaGraphExec_t gexec = nullptr;
  const void *g_vals = nullptr, *g_diag = nullptr, *g_x = nullptr;
  int g_v0 = -1, g_v1 = -1, g_every = 0, g_coarse = -1, g_halo = -1;
  double last_ms[4] = {0, 0, 0, 0};  // setup, coarse inverse, iterations, total (host clock around stream syncs)
  unsigned long long seq_host = 0;   // Pcg2Scalars::seq after the last solve (carried into the next one)
  // peer-memory transport (skb_pcg2_peer_export / skb_pcg2_peer_import)
  void* slab = nullptr;              // this rank's slab (cudaMalloc, IPC-exported)
  PeerSlabLayout lay{};
  void* peer_base[PEER_MAXW] = {};   // mappings of the other ranks' slabs (own rank: slab)
  bool peer_ready = false;
  PeerRedTable red_tab{};
  PeerHaloTable halo_tab{};
  PeerZcTable zc_tab{};
  dvec<double> Xs;                   // my columns of the inverse of the coarse matrix (nc x zc_w, column-major)
  int zc_c0 = 0, zc_w = 0;
  dvec<unsigned int> counters;       // [2] last-CTA counters of the two pushing kernels
  int g_transport = -1;
};
void pcg2_destroy(Pcg2State* s) {
  if (!s) return;
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  for (int q = 0; q < PEER_MAXW; ++q)
    if (s->peer_base[q] && s->peer_base[q] != s->slab) cudaIpcCloseMemHandle(s->peer_base[q]);
  if (s->slab) cudaFree(s->slab);
  delete s;
}
#endif

#if defined(SKB_HAVE_NCCL_H)
// concatenated halo lists (one pack and one unpack kernel for all peers)
static int pcg2_prepare_halo(DistNative& d, Pcg2State& S, cudaStream_t st) {
  if (S.g_halo == (int)d.halo.size() && S.soff.size() == d.halo.size() + 1) return SKB_OK;
  S.soff.assign(d.halo.size() + 1, 0);
  S.roff.assign(d.halo.size() + 1, 0);
  for (size_t i = 0; i < d.halo.size(); ++i) {
    S.soff[i + 1] = S.soff[i] + d.halo[i].ns;
    S.roff[i + 1] = S.roff[i] + d.halo[i].nr;
  }
  S.sidx.resize(S.soff.back() > 0 ? S.soff.back() : 1);
  S.ridx.resize(S.roff.back() > 0 ? S.roff.back() : 1);
  S.sbuf.resize(S.soff.back() > 0 ? S.soff.back() : 1);
  S.rbuf.resize(S.roff.back() > 0 ? S.roff.back() : 1);
  for (size_t i = 0; i < d.halo.size(); ++i) {
    if (d.halo[i].ns) SKB_CUDA(cudaMemcpyAsync(raw(S.sidx) + S.soff[i], d.halo[i].sidx, d.halo[i].ns * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (d.halo[i].nr) SKB_CUDA(cudaMemcpyAsync(raw(S.ridx) + S.roff[i], d.halo[i].ridx, d.halo[i].nr * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  S.g_halo = (int)d.halo.size();
  if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
  return SKB_OK;
}
#endif

}  // namespace skb

extern "C" {

// ---- peer-memory transport: set-up (collective, driven by Shard.enable_peer_transport) -------------------------------
// export: allocates this rank's slab and returns its IPC handle (64 bytes) and meta = [nr_total, recv offset of the
// values arriving from rank 0, 1, ... world-1 (-1: not a neighbour)]
int skb_pcg2_peer_export(skb_plan* pl, void* handle64, int64_t* meta, int64_t meta_len) {
#if !defined(SKB_HAVE_NCCL_H)
  return fail(SKB_EINVAL, "built without nccl.h");
#else
  DistNative* dp = state_of(pl);
  if (!dp || !dp->comm) return fail(SKB_EINVAL, "skb_nccl_init / skb_nccl_set_halo first");
  DistNative& d = *dp;
  if (!handle64 || !meta || meta_len < 1 + d.world) return fail(SKB_EINVAL, "meta needs 1 + world entries");
  if (d.world > PEER_MAXW || d.halo.size() > 8) return fail(SKB_EINVAL, "peer-memory transport: at most 16 ranks and 8 neighbours");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  if (!d.pcg2) d.pcg2 = new Pcg2State();
  Pcg2State& S = *d.pcg2;
  int rc = pcg2_prepare_halo(d, S, pl->stream);
  if (rc) return rc;
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  if (!S.slab) {
    S.lay = PeerSlabLayout::make(d.world, S.roff.back());
    SKB_CUDA(cudaMalloc(&S.slab, S.lay.bytes));
    SKB_CUDA(cudaMemset(S.slab, 0, S.lay.bytes));
    SKB_CUDA(cudaDeviceSynchronize());
    S.counters.assign(4, 0u);
  }
  cudaIpcMemHandle_t h;
  SKB_CUDA(cudaIpcGetMemHandle(&h, S.slab));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  meta[0] = S.roff.back();
  for (int q = 0; q < d.world; ++q) meta[1 + q] = -1;
  for (size_t i = 0; i < d.halo.size(); ++i)
    if (d.halo[i].nr > 0) meta[1 + d.halo[i].peer] = S.roff[i];
  return SKB_OK;
  SKB_CATCH
#endif
}

// import: handles (world x 64 bytes) and metas (world x (1 + world)) of all ranks, in rank order
int skb_pcg2_peer_import(skb_plan* pl, const void* handles, const int64_t* metas) {
#if !defined(SKB_HAVE_NCCL_H)
  return fail(SKB_EINVAL, "built without nccl.h");
#else
  DistNative* dp = state_of(pl);
  if (!dp || !dp->pcg2 || !dp->pcg2->slab) return fail(SKB_EINVAL, "skb_pcg2_peer_export first");
  if (!handles || !metas) return fail(SKB_EINVAL, "null argument");
  DistNative& d = *dp;
  Pcg2State& S = *d.pcg2;
  SKB_CUDA(cudaSetDevice(pl->device));
  const int W = d.world, me = d.rank;
  for (int q = 0; q < W; ++q) {
    if (q == me) {
      S.peer_base[q] = S.slab;
      continue;
    }
    if (S.peer_base[q]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)q * 64, 64);
    void* base = nullptr;
    SKB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    S.peer_base[q] = base;
  }
  S.red_tab.world = W;
  S.red_tab.me = me;
  for (int q = 0; q < W; ++q) {
    const PeerSlabLayout L = PeerSlabLayout::make(W, metas[(size_t)q * (1 + W)]);
    S.red_tab.G[q] = reinterpret_cast<double*>((char*)S.peer_base[q] + L.off_G);
    S.red_tab.fred[q] = reinterpret_cast<unsigned long long*>((char*)S.peer_base[q] + L.off_fred);
  }
  S.zc_tab.world = W;
  S.zc_tab.me = me;
  for (int q = 0; q < W; ++q) {
    const PeerSlabLayout L = PeerSlabLayout::make(W, metas[(size_t)q * (1 + W)]);
    S.zc_tab.zc[q] = reinterpret_cast<double*>((char*)S.peer_base[q] + L.off_zc);
    S.zc_tab.fzc[q] = reinterpret_cast<unsigned long long*>((char*)S.peer_base[q] + L.off_fzc);
  }
  S.halo_tab.np = (int)d.halo.size();
  for (size_t i = 0; i <= d.halo.size(); ++i) S.halo_tab.soff[i] = S.soff[i];
  for (size_t i = 0; i < d.halo.size(); ++i) {
    const int q = d.halo[i].peer;
    const int64_t* mq = metas + (size_t)q * (1 + W);
    const PeerSlabLayout L = PeerSlabLayout::make(W, mq[0]);
    const int64_t off = mq[1 + me];          // where rank q receives MY values
    if (d.halo[i].ns > 0 && off < 0) return fail(SKB_EINVAL, "peer-memory transport: inconsistent halo lists");
    S.halo_tab.rbuf[i] = reinterpret_cast<double*>((char*)S.peer_base[q] + L.off_rbuf) + (off < 0 ? 0 : off);
    S.halo_tab.fhalo[i] = reinterpret_cast<unsigned long long*>((char*)S.peer_base[q] + L.off_fhalo) + me;
    S.halo_tab.peer[i] = q;
  }
  S.peer_ready = true;
  if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
  return SKB_OK;
#endif
}

int skb_dist_pcg2(skb_plan* pl, const skb_dist_pcg2_args* a, int32_t* iters, double* relres) {
#if !defined(SKB_HAVE_NCCL_H)
  return fail(SKB_EINVAL, "built without nccl.h: the native NCCL path is not available");
#else
  if (!nccl().ok) return fail(SKB_EINVAL, "NCCL unavailable: " + nccl().why);
  if (!pl || !a || !iters || !relres) return fail(SKB_EINVAL, "null argument");
  DistNative* dp = state_of(pl);
  if (!dp || !dp->comm) return fail(SKB_EINVAL, "skb_nccl_init first");
  DistNative& d = *dp;
  if (!a->vals || !a->rhs || !a->x) return fail(SKB_EINVAL, "null vector");
  const int v0 = a->v0, v1 = a->v1;
  if (v0 < 0 || v1 > pl->d.n || v0 > v1) return fail(SKB_EINVAL, "bad owned row range");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  if (!d.pcg2) d.pcg2 = new Pcg2State();
  Pcg2State& S = *d.pcg2;
  NcclApi& n = nccl();
  const int D = pl->d.dim;
  const int64_t nd = pl->ndof();
  const PlanView pv = pl->view();
  const bool use_graph = a->use_graph != 0;
  cudaStream_t caller = (cudaStream_t)a->stream;
  cudaStream_t st = pl->stream;   // capturable, non-default; joined with the caller's stream on entry and exit
  cudaEvent_t ev = nullptr;
  struct Guard {
    cudaEvent_t& ev;
    cudaStream_t st, caller;
    ~Guard() {
      if (ev) {
        cudaEventRecord(ev, st);
        cudaStreamWaitEvent(caller, ev, 0);
        cudaEventDestroy(ev);
      }
    }
  } guard{ev, st, caller};
  SKB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SKB_CUDA(cudaEventRecord(ev, caller));
  SKB_CUDA(cudaStreamWaitEvent(st, ev, 0));

  // ---- state ------------------------------------------------------------------------------------------------
  bool coarse = a->use_coarse && pl->coarse && pl->coarse->n_agg > 0;
  const int NC = D == 3 ? 6 : 3;
  const int n_agg = coarse ? pl->coarse->n_agg : 0;
  const int nc = NC * n_agg;
  if (S.nd != nd) {
    S.r.assign(nd, 0.0); S.u.assign(nd, 0.0); S.w.assign(nd, 0.0); S.p.assign(nd, 0.0); S.s.assign(nd, 0.0);
    S.dinv.assign((size_t)pl->d.n * D * D, 0.0);
    S.part.assign(3 * PCG2_GRID, 0.0);
    S.sc.resize(1);
    S.nd = nd;
    S.g_halo = -1;
  }
  if (S.nc != nc || S.red.size() != (size_t)(4 + nc)) {
    S.red.assign(4 + nc, 0.0);
    S.rc.assign(nc > 0 ? nc : 1, 0.0); S.rcs.assign(nc > 0 ? nc : 1, 0.0); S.zc.assign(nc > 0 ? nc : 1, 0.0);
    if (nc > 0) S.Ainv.resize((size_t)nc * nc);
    S.nc = nc;
    if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
  }
  int rc_;
  if ((rc_ = pcg2_prepare_halo(d, S, st))) return rc_;
  // transport of the per-iteration exchanges: 1 = stores into the other ranks' HBM over NVLink (CUDA IPC mappings,
  // flags instead of collective calls), 0 = NCCL (grouped send/recv + all-reduce)
  const bool peer = a->transport == 1 && S.peer_ready && 4 + nc <= PEER_RED_CAP;
  if (a->transport == 1 && !peer) return fail(SKB_EINVAL, "peer-memory transport requested but not set up (skb_pcg2_peer_export / _import)");
  Pcg2Scalars* sc = raw(S.sc);
  double *r = raw(S.r), *u = raw(S.u), *w = raw(S.w), *p = raw(S.p), *s = raw(S.s), *x = a->x, *red = raw(S.red);
  const int64_t ns = S.soff.back(), nr = S.roff.back();

  // ---- set-up of this solve -----------------------------------------------------------------------------------
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  SKB_CUDA(cudaStreamSynchronize(st));
  const double t_begin = now();
  double t_inv0 = t_begin, t_inv1 = t_begin;
  SKB_CUDA(cudaMemsetAsync(u, 0, nd * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(w, 0, nd * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(p, 0, nd * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(s, 0, nd * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(x, 0, nd * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(red, 0, (4 + nc) * sizeof(double), st));
  SKB_CUDA(cudaMemcpyAsync(r, a->rhs, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (D == 3)
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_dinv_kernel<3><<<PCG2_GRID, PCG_THREADS, 0, st>>>(pv, a->vals, a->diag, v0, v1, raw(S.dinv)));
  else
    SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_dinv_kernel<2><<<PCG2_GRID, PCG_THREADS, 0, st>>>(pv, a->vals, a->diag, v0, v1, raw(S.dinv)));
  if (coarse) {
    // coarse matrix of this system: owned fine blocks per rank, summed over the ranks, inverted by every rank
    rc_ = D == 3 ? coarse_assemble_launch<3>(pl, a->vals, a->diag, raw(S.Ainv), st) : coarse_assemble_launch<2>(pl, a->vals, a->diag, raw(S.Ainv), st);
    if (rc_) return rc_;
    if ((rc_ = all_reduce(d, raw(S.Ainv), (size_t)nc * nc, st))) return rc_;
    SKB_CUDA(cudaStreamSynchronize(st));
    t_inv0 = now();
    if (peer) {
      // every rank factors the (identical) coarse matrix and keeps only ITS columns of the inverse
      const int W = d.world;
      S.zc_c0 = (int)((int64_t)nc * d.rank / W);
      S.zc_w = (int)((int64_t)nc * (d.rank + 1) / W) - S.zc_c0;
      if (S.Xs.size() < (size_t)nc * (S.zc_w > 0 ? S.zc_w : 1)) S.Xs.resize((size_t)nc * (S.zc_w > 0 ? S.zc_w : 1));
      if (coarse_factor_slice(pl, raw(S.Ainv), nc, S.zc_c0, S.zc_w, raw(S.Xs), st) != 0) coarse = false;
    } else if (coarse_invert(pl, raw(S.Ainv), nc, st) != 0) {
      coarse = false;   // degenerate aggregate: block-Jacobi for this solve
    }
    SKB_CUDA(cudaStreamSynchronize(st));
    t_inv1 = now();
  }
  SKB_CUDA(cudaMemsetAsync(raw(S.rc), 0, S.rc.size() * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(raw(S.rcs), 0, S.rcs.size() * sizeof(double), st));
  SKB_CUDA(cudaMemsetAsync(raw(S.zc), 0, S.zc.size() * sizeof(double), st));
  Pcg2Scalars h0;
  memset(&h0, 0, sizeof(h0));
  h0.gamma_old = h0.alpha_old = 1.0;
  h0.rtol2 = a->rtol * a->rtol;
  h0.seq = S.seq_host;
  SKB_CUDA(cudaMemcpyAsync(sc, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  const CoarseSpace* cs = coarse ? pl->coarse : nullptr;
  if (coarse) {
    // rc = P^T r_0 over the owned vertices, summed over the ranks (the only extra collective of the solve)
    CoarseView cv;
    cv.n_agg = n_agg; cv.nc = nc; cv.agg = raw(cs->agg); cv.xrel = raw(cs->xrel); cv.vord = raw(cs->vord); cv.aptr = raw(cs->aptr);
    cv.Ainv = nullptr; cv.rc = raw(S.rc); cv.zc = nullptr;
    if (D == 3)
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_restrict_kernel<3><<<n_agg, 256, 0, st>>>(cv, r, nullptr));
    else
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, coarse_restrict_kernel<2><<<n_agg, 256, 0, st>>>(cv, r, nullptr));
    if ((rc_ = all_reduce(d, raw(S.rc), nc, st))) return rc_;
  }
  const int nc_it = coarse ? nc : 0;
  const PeerSlabLayout& L = S.lay;
  const double* myG = peer ? reinterpret_cast<const double*>((char*)S.slab + L.off_G) : nullptr;
  const unsigned long long* my_fred = peer ? reinterpret_cast<const unsigned long long*>((char*)S.slab + L.off_fred) : nullptr;
  const unsigned long long* my_fhalo = peer ? reinterpret_cast<const unsigned long long*>((char*)S.slab + L.off_fhalo) : nullptr;
  const double* my_rbuf = peer ? reinterpret_cast<const double*>((char*)S.slab + L.off_rbuf) : nullptr;
  const double* my_zc = peer ? reinterpret_cast<const double*>((char*)S.slab + L.off_zc) : nullptr;
  const unsigned long long* my_fzc = peer ? reinterpret_cast<const unsigned long long*>((char*)S.slab + L.off_fzc) : nullptr;
  const double* zc_use = (peer && coarse) ? my_zc : raw(S.zc);
  const int zc_world = (peer && coarse) ? d.world : 0;
  SKB_CUDA(cudaStreamSynchronize(st));
  const double t_iter0 = now();

  auto iteration = [&]() -> int {
    SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg2_scalars_kernel<<<1, 1024, 0, st>>>(sc, red, nc_it, raw(S.rc), raw(S.rcs), d.world,
                                                                           peer ? myG : nullptr, my_fred));
    if (coarse && peer)
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg3_gemv_push_kernel<<<(S.zc_w * 32 + 255) / 256 > 0 ? (S.zc_w * 32 + 255) / 256 : 1, 256, 0, st>>>(
                     sc, nc, S.zc_c0, S.zc_w, raw(S.Xs), raw(S.rc), S.zc_tab, raw(S.counters) + 2));
    else if (coarse)
      SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_gemv_kernel<<<(nc * 32 + 255) / 256, 256, 0, st>>>(sc, nc, raw(S.Ainv), raw(S.rc), raw(S.zc)));
    const int* agg = coarse ? raw(cs->agg) : nullptr;
    const double* xrel = coarse ? raw(cs->xrel) : nullptr;
    if (coarse && peer)   // the slices of zc from all ranks
      SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg3_wait_kernel<<<1, 32, 0, st>>>(sc, my_fzc, d.world, 0, S.halo_tab));
    if (D == 3) {
      if (coarse)
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_update_kernel<3, true><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, v0, v1, raw(S.dinv), agg, xrel, zc_use, w, u, p, s, x, r, raw(S.part), zc_world, my_fzc, sc));
      else
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_update_kernel<3, false><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, v0, v1, raw(S.dinv), agg, xrel, zc_use, w, u, p, s, x, r, raw(S.part), zc_world, my_fzc, sc));
    } else {
      if (coarse)
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_update_kernel<2, true><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, v0, v1, raw(S.dinv), agg, xrel, zc_use, w, u, p, s, x, r, raw(S.part), zc_world, my_fzc, sc));
      else
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_update_kernel<2, false><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, v0, v1, raw(S.dinv), agg, xrel, zc_use, w, u, p, s, x, r, raw(S.part), zc_world, my_fzc, sc));
    }
    // halo of u
    if (peer) {
      // stores into the neighbours' receive areas + flags; then wait for theirs and unpack
      if (ns > 0)
        SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg3_pack_push_kernel<<<(unsigned)((ns + 255) / 256 < 296 ? (ns + 255) / 256 : 296), 256, 0, st>>>(sc, u, raw(S.sidx), S.halo_tab, raw(S.counters)));
      if (nr > 0) {
        SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg3_wait_kernel<<<1, 32, 0, st>>>(sc, my_fhalo, S.halo_tab.np, 1, S.halo_tab));
        SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg3_unpack_kernel<<<(unsigned)((nr + 255) / 256 < 296 ? (nr + 255) / 256 : 296), 256, 0, st>>>(sc, u, raw(S.ridx), nr, my_rbuf));
      }
    } else {
      // pack, one grouped send/recv, unpack
      if (ns > 0)
        SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg2_pack_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(sc, u, raw(S.sidx), ns, raw(S.sbuf)));
      if (!d.halo.empty()) {
        SKB_NCCL(n.GroupStart());
        for (size_t i = 0; i < d.halo.size(); ++i)
          if (d.halo[i].ns > 0) SKB_NCCL(n.Send(raw(S.sbuf) + S.soff[i], (size_t)d.halo[i].ns, ncclDouble, d.halo[i].peer, d.comm, st));
        for (size_t i = 0; i < d.halo.size(); ++i)
          if (d.halo[i].nr > 0) SKB_NCCL(n.Recv(raw(S.rbuf) + S.roff[i], (size_t)d.halo[i].nr, ncclDouble, d.halo[i].peer, d.comm, st));
        SKB_NCCL(n.GroupEnd());
      }
      if (nr > 0)
        SKB_LAUNCH(pl, SKB_K_OTHER, st, pcg2_unpack_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(sc, u, raw(S.ridx), nr, raw(S.rbuf)));
    }
    if (D == 3)
      SKB_LAUNCH(pl, SKB_K_SPMV, st, pcg2_spmv_kernel<3><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, pv, a->vals, a->diag, u, w, v0, v1, raw(S.part)));
    else
      SKB_LAUNCH(pl, SKB_K_SPMV, st, pcg2_spmv_kernel<2><<<PCG2_GRID, PCG_THREADS, 0, st>>>(sc, pv, a->vals, a->diag, u, w, v0, v1, raw(S.part)));
    const int* vord = coarse ? raw(cs->vord) : nullptr;
    const int* aptr = coarse ? raw(cs->aptr) : nullptr;
    const int ngrid_r = n_agg * (coarse ? 1 : 0) + 1;
    const int nagg_r = coarse ? n_agg : 0;
    unsigned int* ctr = peer ? raw(S.counters) + 1 : nullptr;
    if (peer) {
      if (D == 3)
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_restrict_kernel<3, true><<<ngrid_r, 256, 0, st>>>(sc, nagg_r, vord, aptr, xrel, w, raw(S.part), PCG2_GRID, red, S.red_tab, ctr));
      else
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_restrict_kernel<2, true><<<ngrid_r, 256, 0, st>>>(sc, nagg_r, vord, aptr, xrel, w, raw(S.part), PCG2_GRID, red, S.red_tab, ctr));
    } else {
      if (D == 3)
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_restrict_kernel<3, false><<<ngrid_r, 256, 0, st>>>(sc, nagg_r, vord, aptr, xrel, w, raw(S.part), PCG2_GRID, red, S.red_tab, ctr));
      else
        SKB_LAUNCH(pl, SKB_K_PCG_VECTOR, st, pcg2_restrict_kernel<2, false><<<ngrid_r, 256, 0, st>>>(sc, nagg_r, vord, aptr, xrel, w, raw(S.part), PCG2_GRID, red, S.red_tab, ctr));
      SKB_NCCL(n.AllReduce(red, red, (size_t)(4 + nc_it), ncclDouble, ncclSum, d.comm, st));
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(SKB_ECUDA, std::string("pcg2 iteration: ") + cudaGetErrorString(le));
    return SKB_OK;
  };

  // ---- iterations: the first chunk runs eagerly (NCCL sets up its connections), full chunks replay one graph ----
  const int every = a->check_every > 0 ? a->check_every : 25;
  const bool graph_ok = use_graph && S.gexec && S.g_vals == a->vals && S.g_diag == a->diag && S.g_x == a->x && S.g_v0 == v0 &&
                        S.g_v1 == v1 && S.g_every == every && S.g_coarse == (coarse ? 1 : 0) && S.g_transport == (peer ? 1 : 0);
  if (!graph_ok && S.gexec) {
    cudaGraphExecDestroy(S.gexec);
    S.gexec = nullptr;
  }
  Pcg2Scalars h;
  int launched = 0;      // iteration passes issued (the bootstrap pass included)
  const int max_pass = a->max_iter + 1;
  bool first_chunk = true;
  for (;;) {
    const int nrun = (max_pass - launched) < every ? (max_pass - launched) : every;
    if (nrun <= 0) break;
    if (use_graph && nrun == every && (S.gexec || !first_chunk)) {
      if (!S.gexec) {
        const bool timing = pl->timing;
        pl->timing = false;   // no event records inside the capture
        cudaGraph_t graph = nullptr;
        SKB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        rc_ = SKB_OK;
        for (int k = 0; k < nrun && rc_ == SKB_OK; ++k) rc_ = iteration();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);   // always ends the capture, also after a failed step
        pl->timing = timing;
        if (rc_) {
          if (graph) cudaGraphDestroy(graph);
          return rc_;
        }
        if (ce != cudaSuccess) return fail(SKB_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
        const cudaError_t ci = cudaGraphInstantiate(&S.gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) {
          S.gexec = nullptr;
          return fail(SKB_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ci));
        }
        S.g_vals = a->vals; S.g_diag = a->diag; S.g_x = a->x; S.g_v0 = v0; S.g_v1 = v1; S.g_every = every;
        S.g_coarse = coarse ? 1 : 0;
        S.g_transport = peer ? 1 : 0;
      }
      SKB_CUDA(cudaGraphLaunch(S.gexec, st));
      pl->launches += nrun * 7;
    } else {
      for (int k = 0; k < nrun; ++k)
        if ((rc_ = iteration())) return rc_;
    }
    first_chunk = false;
    launched += nrun;
    SKB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
    SKB_CUDA(cudaStreamSynchronize(st));
    if (h.done) break;
  }
  S.seq_host = h.seq;
  const double t_end = now();
  S.last_ms[0] = (t_iter0 - t_begin) - (t_inv1 - t_inv0);   // set-up without the dense inverse (incl. the nc^2 all-reduce)
  S.last_ms[1] = t_inv1 - t_inv0;                           // cuSOLVER potrf + potri of the coarse matrix
  S.last_ms[2] = t_end - t_iter0;                           // the iterations
  S.last_ms[3] = t_end - t_begin;
  *iters = h.iters;
  *relres = (h.bb > 0.0) ? sqrt(h.rr / h.bb) : 0.0;
  if (h.done == 5) return fail(SKB_ECUDA, "distributed PCG: a peer-memory wait timed out (a neighbour rank did not deliver its halo values or reduction partials)");
  if (h.done == 2) return fail(SKB_ECUDA, "distributed PCG broke down (matrix not positive definite on the Krylov space, or NaN)");
  return SKB_OK;
  SKB_CATCH
#endif
}

/* host-clock breakdown of the last skb_dist_pcg2 of this plan, ms: set-up, dense coarse inverse, iterations, total */
int skb_dist_pcg2_times(skb_plan* pl, double out[4]) {
#if defined(SKB_HAVE_NCCL_H)
  DistNative* d = state_of(pl);
  if (!d || !d->pcg2 || !out) return fail(SKB_EINVAL, "no single-reduction solve has run on this plan");
  for (int i = 0; i < 4; ++i) out[i] = d->pcg2->last_ms[i];
  return SKB_OK;
#else
  return fail(SKB_EINVAL, "built without nccl.h");
#endif
}

}  // extern "C"
