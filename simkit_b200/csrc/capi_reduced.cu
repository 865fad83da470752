// C ABI, part 4: reduced (subspace) operators.
//
//   Hr = JB^T blockdiag(vol * psd(He)) JB      energies/elastic.py:749-782 and the *_hessian_u
//                                              tier called with a dense operator
//   fast_sandwich_transform_clustered          fast_sandwich_transform_clustered.py:15-158
//
// Structure of the reduced Hessian: per element chunk the CTA stages the rows
// JB_e (b x r) and Y_e = He JB_e in shared memory and accumulates JB_e^T Y_e on
// FP64 tensor-core tiles (mma.sync m8n8k4 -> SASS DMMA.8x8x4, the only FP64 MMA
// sm_100a has: tcgen05 has no .kind::f64).  CTAs own disjoint element ranges and
// write per-CTA partial matrices that a second kernel sums in CTA order, so the
// result is deterministic.
#include <algorithm>
#include <numeric>
#include <vector>

#include "capi_common.cuh"

namespace skb {

static double g_reduced_ms[3] = {0.0, 0.0, 0.0};  // skb_reduced_last_times

constexpr int RH_THREADS = 256;  // 8 warps
constexpr int RH_ELEMS = 4;      // elements per k-chunk (rows = RH_ELEMS * b, multiple of 4 for b in {4, 9})

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// F = rows . z + Jx0 for one element, He and P; writes Y_e rows and P-weighted vector.
template <int D>
__device__ __forceinline__ void reduced_element(int material, int psd_mode, const Mat<D>& F, double mu, double lam,
                                                double vol, double* H /* b*b */, Mat<D>& P, double& psi) {
  constexpr int B = D * D;
  psi = vol * energy_density<D>(material, F, mu, lam);
  P = pk1<D>(material, F, mu, lam);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) P.m[i][j] *= vol;
  if (material == MAT_LINEAR_ELASTICITY) {
    // dispatcher semantics: floor lifts the zero (skew) modes; own-module callers pass PSD_NONE
    double w_sym = 2.0 * mu, w_tr = 2.0 * mu + D * lam, w_skew = 0.0;
    const double pre = (psd_mode == PSD_BEFORE_VOL) ? 1.0 : vol;
    const double post = (psd_mode == PSD_BEFORE_VOL) ? vol : 1.0;
    w_sym *= pre; w_tr *= pre; w_skew *= pre;
    if (psd_mode != PSD_NONE) {
      w_sym = psd_clamp(w_sym, psd_mode);
      w_tr = psd_clamp(w_tr, psd_mode);
      w_skew = psd_clamp(w_skew, psd_mode);
    }
    w_sym *= post; w_tr *= post; w_skew *= post;
    const double cI = 0.5 * (w_sym + w_skew), cT = 0.5 * (w_sym - w_skew), cR = (w_tr - w_sym) / D;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j)
        for (int k = 0; k < D; ++k)
          for (int l = 0; l < D; ++l) {
            double v = 0.0;
            if (i == k && j == l) v += cI;
            if (i == l && j == k) v += cT;
            if (i == j && k == l) v += cR;
            H[(i * D + j) * B + k * D + l] = v;
          }
    return;
  }
  Mat<D> U, V;
  Vec<D> s;
  svd_rv(F, U, s, V);
  Principal<D> h = principal_hessian<D>(material, s, mu, lam);
  weight_and_project<D>(h, vol, psd_mode);
  expand_hessian<D>(h, U, V, H);
}

// Pass 1: per element F, He (to global scratch, b*b per element), vol*P, vol*psi.
// rows(e) come either from a dense JB (t*b, r) or, when JB == nullptr, from the mesh:
// F = J (B z + x0) was formed by the caller into Fin.
template <int D>
__global__ void reduced_pass1_kernel(int material, int psd_mode, int64_t t, const double* Fin, const double* mu,
                                     int mu_s, const double* lam, int lam_s, const double* vol, int vol_s,
                                     double* Hout, double* Pout, double* psiout) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= t) return;
  constexpr int B = D * D;
  Mat<D> F;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) F.m[i][j] = Fin[e * B + i * D + j];
  double H[B * B];
  Mat<D> P;
  double psi;
  reduced_element<D>(material, psd_mode, F, mu[e * mu_s], lam ? lam[e * lam_s] : 0.0, vol[e * vol_s], H, P, psi);
  for (int i = 0; i < B * B; ++i) Hout[e * B * B + i] = H[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) Pout[e * B + i * D + j] = P.m[i][j];
  psiout[e] = psi;
}

// y[row] = sum_c M[row][c] z[c] + add[row]   (one warp per row, coalesced)
__global__ void gemv_rows_kernel(int64_t rows, int cols, const double* M, const double* z, const double* add,
                                 double* y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s = 0.0;
  for (int c = lane; c < cols; c += 32) s = fma(M[row * cols + c], z[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s + (add ? add[row] : 0.0);
}

// Rows of the reduced operator for element e, written to shared memory [row][r]:
//   dense:      JB[(e*b + i*D + j), :]
//   from basis: sum_a D[j][a] * Bm[T[e][a]*D + i, :]
template <int D>
__device__ __forceinline__ void stage_rows(const double* JB, const PlanView* p, const double* Bm, int64_t e, int r,
                                           int rpad, double* srow /* b x rpad */) {
  constexpr int B = D * D;
  constexpr int K = D + 1;
  if (JB) {
    for (int idx = threadIdx.x; idx < B * r; idx += blockDim.x) {
      const int row = idx / r, c = idx - row * r;
      srow[row * rpad + c] = JB[(e * B + row) * (int64_t)r + c];
    }
  } else {
    for (int idx = threadIdx.x; idx < B * r; idx += blockDim.x) {
      const int row = idx / r, c = idx - row * r;
      const int i = row / D, j = row - i * D;
      double s = 0.0, d0 = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const double d = p->Dm[(size_t)(j * D + a) * p->t + e];
        d0 -= d;
        s = fma(d, Bm[((size_t)p->T32[e * K + a + 1] * D + i) * r + c], s);
      }
      s = fma(d0, Bm[((size_t)p->T32[e * K] * D + i) * r + c], s);
      srow[row * rpad + c] = s;
    }
  }
}

// Pass 2: partial Hr and gr per CTA over its element range.  blockIdx.x = element range,
// blockIdx.y = output column panel (tile columns [y*tcp, (y+1)*tcp)).  Output tiles are 8x8
// DMMA tiles; warp w owns panel tiles w, w+nw, ... (fully unrolled so accumulators stay in registers).
constexpr int RH_MAXT = 44;
template <int D>
__global__ void __launch_bounds__(RH_THREADS, 1)
reduced_pass2_kernel(int64_t t, int r, int tcp, const double* JB, PlanView pv, int use_plan, const double* Bm,
                     const double* He, const double* Pw, double* Hpart, double* gpart) {
  constexpr int B = D * D;
  constexpr int ROWS = RH_ELEMS * B;  // 36 or 16: multiple of 4
  extern __shared__ double sm[];
  const int rt = (r + 7) / 8;          // tiles per side
  const int rpad = rt * 8 + 1;         // +1: de-conflict column reads
  double* sJ = sm;                     // ROWS x rpad
  double* sY = sm + (size_t)ROWS * rpad;
  const PlanView* p = use_plan ? &pv : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int tc0 = blockIdx.y * tcp;
  const int tc1 = (tc0 + tcp < rt) ? tc0 + tcp : rt;
  const int ntiles = rt * (tc1 - tc0);
  double acc[RH_MAXT][2];
#pragma unroll
  for (int i = 0; i < RH_MAXT; ++i) acc[i][0] = acc[i][1] = 0.0;
  double gacc = 0.0;  // thread c (< r) accumulates gr[c]  (panel 0 only)

  const int64_t per = (t + gridDim.x - 1) / gridDim.x;
  const int64_t e0 = (int64_t)blockIdx.x * per;
  const int64_t e1 = (e0 + per < t) ? e0 + per : t;
  for (int64_t eb = e0; eb < e1; eb += RH_ELEMS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ROWS * rpad; idx += blockDim.x) {
      sJ[idx] = 0.0;
      sY[idx] = 0.0;
    }
    __syncthreads();
    for (int le = 0; le < RH_ELEMS; ++le) {
      const int64_t e = eb + le;
      if (e < e1) stage_rows<D>(use_plan ? nullptr : JB, p, Bm, e, r, rpad, sJ + (size_t)le * B * rpad);
    }
    __syncthreads();
    // Y_e = He_e * JB_e (panel columns only) and gradient accumulation
    const int c_lo = tc0 * 8, c_hi = (tc1 * 8 < r) ? tc1 * 8 : r;
    const int pc = c_hi - c_lo;
    for (int idx = threadIdx.x; idx < ROWS * pc; idx += blockDim.x) {
      const int row = idx / pc, c = c_lo + (idx - row * pc);
      const int le = row / B, rb = row - le * B;
      const int64_t e = eb + le;
      if (e < e1) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < B; ++k) s = fma(He[(e * B + rb) * B + k], sJ[(le * B + k) * rpad + c], s);
        sY[row * rpad + c] = s;
      }
    }
    if (blockIdx.y == 0 && threadIdx.x < r) {
      const int c = threadIdx.x;
      for (int row = 0; row < ROWS; ++row) {
        const int64_t e = eb + row / B;
        if (e < e1) gacc = fma(sJ[row * rpad + c], Pw[e * B + (row % B)], gacc);
      }
    }
    __syncthreads();
    // Hr += sJ^T sY on 8x8 tiles: A[m][k] = sJ[k][m], B[k][n] = sY[k][n]
#pragma unroll
    for (int slot = 0; slot < RH_MAXT; ++slot) {
      const int tile = warp + slot * nw;
      if (tile < ntiles) {
        const int tm = tile / (tc1 - tc0), tn = tc0 + (tile - tm * (tc1 - tc0));
        double c0 = acc[slot][0], c1 = acc[slot][1];
#pragma unroll
        for (int k0 = 0; k0 < ROWS; k0 += 4) {
          const double a = sJ[(k0 + (lane & 3)) * rpad + tm * 8 + (lane >> 2)];
          const double b = sY[(k0 + (lane & 3)) * rpad + tn * 8 + (lane >> 2)];
          dmma_m8n8k4(c0, c1, a, b);
        }
        acc[slot][0] = c0;
        acc[slot][1] = c1;
      }
    }
  }
#pragma unroll
  for (int slot = 0; slot < RH_MAXT; ++slot) {
    const int tile = warp + slot * nw;
    if (tile < ntiles) {
      const int tm = tile / (tc1 - tc0), tn = tc0 + (tile - tm * (tc1 - tc0));
      const int row = tm * 8 + (lane >> 2), col = tn * 8 + 2 * (lane & 3);
      if (row < r && col < r) Hpart[((size_t)blockIdx.x * r + row) * r + col] = acc[slot][0];
      if (row < r && col + 1 < r) Hpart[((size_t)blockIdx.x * r + row) * r + col + 1] = acc[slot][1];
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < r) gpart[(size_t)blockIdx.x * r + threadIdx.x] = gacc;
}

// ---------------------------------------------------------------------------------------------
// Pass 2, register-blocked and TMA-staged (used when r is even, so that every operator row is a
// 16-byte multiple).  One persistent CTA of 8 warps per SM walks its element range in chunks of CE
// elements:
//   * the raw operator rows of the NEXT chunk -- the 12 (6 in 2D) rows B[(v_a, i), :] of each element's
//     corner vertices, or its b rows of a dense JB -- are fetched by the TMA unit (1-D bulk copies
//     completing on an mbarrier) into a double-buffered landing area while the current chunk computes;
//   * sJ = rows of J B for the chunk (4 FMAs per value from the landing area), sY = He . sJ;
//   * Hr += sJ^T sY on FP64 tensor-core tiles (DMMA.8x8x4).  The symmetric output is cut into blocks of
//     RB_T x RB_T tiles; only the block pairs (bi <= bj) are computed (15 of 25 at r = 200) and a warp
//     owns up to two of them, so one k-step of a pair loads 2 x RB_T operand fragments from shared memory
//     for RB_T^2 DMMAs (10 loads per 25 instead of 50 per 25).
// The shared-memory row stride is = 4 (mod 16) doubles, which makes every fragment load conflict-free.
// Per-CTA partial matrices are summed in CTA order and the lower triangle is mirrored from the upper one
// (deterministic, exactly symmetric).
constexpr int RB_T = 5;           // tiles per block side (40 columns)
constexpr int RB_THREADS = 256;   // 8 warps
constexpr int RB_PPW = 2;         // block pairs per warp (2 x 50 accumulator doubles in registers)
constexpr int RB_MAXPAIRS = RB_PPW * RB_THREADS / 32;  // block pairs per pass

template <int D>
struct RBShape {
  static constexpr int B = D * D;
  static constexpr int CE = (D == 3) ? 2 : 4;            // elements per chunk
  static constexpr int KR = ((CE * B + 3) / 4) * 4;       // rows per chunk, padded to the DMMA k = 4
  static constexpr int RAWROWS = (D + 1) * D;             // landing rows per element (from a basis)
};

// Row stride (doubles) of the sJ / sY staging: every tile column a block pair can touch exists (zero beyond r,
// so the fragment loads need no guards), rounded to 16 and + 4 so that the stride is = 4 (mod 16).
__host__ __device__ inline int rb_stride(int r) {
  const int rt = (r + 7) / 8;
  const int cols = ((rt + RB_T - 1) / RB_T) * RB_T * 8;
  return ((cols + 15) / 16) * 16 + 4;
}

template <int D>
__global__ void __launch_bounds__(RB_THREADS, 1)
reduced_pass2_blocked_kernel(int64_t t, int r, const double* JB, PlanView pv, int use_plan, const double* Bm,
                             const double* He, const double* Pw, double* Hpart, double* gpart) {
  using S = RBShape<D>;
  constexpr int B = S::B, CE = S::CE, KR = S::KR, K = D + 1;
  constexpr int NSMALL = CE * (B * B + B + D * D);         // He, vol*P and D of a chunk: one value per thread
  static_assert(NSMALL <= RB_THREADS, "small operands are prefetched one per thread");
  extern __shared__ __align__(16) double sm[];
  const int rt = (r + 7) / 8;
  const int RP = rb_stride(r);
  const int nblk = (rt + RB_T - 1) / RB_T;
  const int npairs_all = nblk * (nblk + 1) / 2;
  const int pair0 = blockIdx.y * RB_MAXPAIRS;
  const int rawrows = use_plan ? S::RAWROWS : B;           // landing rows per element
  double* sJ = sm;                                         // 2 x KR x RP
  double* sY = sJ + (size_t)2 * KR * RP;                   // 2 x KR x RP
  double* sRaw = sY + (size_t)2 * KR * RP;                 // 2 x CE x rawrows x r
  double* sSmall = sRaw + (size_t)2 * CE * rawrows * r;    // 2 x NSMALL: [He | vol*P | D] of the chunk
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sSmall + 2 * NSMALL);  // 2 mbarriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // this warp's block pairs (bi <= bj), enumerated row-major over the upper triangle
  int pbi[RB_PPW], pbj[RB_PPW];
#pragma unroll
  for (int s = 0; s < RB_PPW; ++s) {
    int pi = pair0 + warp + s * (RB_THREADS / 32);
    pbi[s] = -1;
    pbj[s] = -1;
    if (pi < npairs_all && pi < pair0 + RB_MAXPAIRS) {
      int bi = 0;
      while (pi >= nblk - bi) {
        pi -= nblk - bi;
        ++bi;
      }
      pbi[s] = bi;
      pbj[s] = bi + pi;
    }
  }
  double acc[RB_PPW][RB_T][RB_T][2];
#pragma unroll
  for (int s = 0; s < RB_PPW; ++s)
#pragma unroll
    for (int a = 0; a < RB_T; ++a)
#pragma unroll
      for (int b = 0; b < RB_T; ++b) acc[s][a][b][0] = acc[s][a][b][1] = 0.0;
  double gacc = 0.0;

  for (int idx = threadIdx.x; idx < 4 * KR * RP; idx += blockDim.x) sm[idx] = 0.0;  // pad rows / columns stay zero
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(sBar), 1);
    mbar_init(smem_u32(sBar + 1), 1);
  }
  __syncthreads();

  const int64_t per = ((t + gridDim.x - 1) / gridDim.x + CE - 1) / CE * CE;
  const int64_t e0 = (int64_t)blockIdx.x * per;
  const int64_t e1 = (e0 + per < t) ? e0 + per : t;
  const unsigned rowbytes = (unsigned)r * 8u;

  auto issue = [&](int64_t eb, int buf) {  // thread 0: fetch the landing rows of chunk eb into buffer buf
    const unsigned mbar = smem_u32(sBar + buf);
    int ne = 0;
    for (int le = 0; le < CE; ++le) ne += (eb + le < e1) ? 1 : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(mbar, (unsigned)(ne * rawrows) * rowbytes);
    for (int le = 0; le < CE; ++le) {
      const int64_t e = eb + le;
      if (e >= e1) break;
      double* dst = sRaw + ((size_t)buf * CE + le) * rawrows * r;
      if (use_plan) {
        for (int a = 0; a < K; ++a) {
          const int v = pv.T32[e * K + a];
          // the D rows (v, 0..D-1) of B are contiguous
          tma_bulk_g2s(smem_u32(dst + (size_t)a * D * r), Bm + (size_t)v * D * r, (unsigned)D * rowbytes, mbar);
        }
      } else {
        tma_bulk_g2s(smem_u32(dst), JB + (size_t)e * B * r, (unsigned)B * rowbytes, mbar);
      }
    }
  };
  // one small operand of chunk eb per thread (registers; stored to shared memory one iteration later)
  auto fetch_small = [&](int64_t eb) -> double {
    const int idx = threadIdx.x;
    if (idx < CE * B * B) {
      const int64_t e = eb + idx / (B * B);
      return (e < e1) ? He[e * B * B + (idx % (B * B))] : 0.0;
    }
    if (idx < CE * (B * B + B)) {
      const int k = idx - CE * B * B;
      const int64_t e = eb + k / B;
      return (e < e1) ? Pw[e * B + (k % B)] : 0.0;
    }
    if (idx < NSMALL && use_plan) {
      const int k = idx - CE * (B * B + B);
      const int64_t e = eb + k / (D * D);
      return (e < e1) ? pv.Dm[(size_t)(k % (D * D)) * pv.t + e] : 0.0;  // [j*D + c] -> D[j][c+1]
    }
    return 0.0;
  };
  // sJ / sY of chunk eb into buffer b: a thread owns one column (all rows of both), so nothing it reads was
  // written by another thread of this phase; sY re-reads the thread's own sJ column from shared memory, which
  // keeps this phase at a handful of live registers next to the 100 accumulator doubles
  auto form = [&](int64_t eb, int b) {
    const int c = threadIdx.x;
    if (c >= r) return;
    const double* raw = sRaw + (size_t)b * CE * rawrows * r;
    const double* sH = sSmall + (size_t)b * NSMALL;
    const double* sP = sH + CE * B * B;
    const double* sD = sP + CE * B;
    double* oJ = sJ + (size_t)b * KR * RP + c;
    double* oY = sY + (size_t)b * KR * RP + c;
#pragma unroll 1
    for (int le = 0; le < CE; ++le) {
      const bool live = eb + le < e1;
#pragma unroll 1
      for (int i = 0; i < D; ++i) {
        if (use_plan) {
          const double* rr = raw + (size_t)le * rawrows * r + (size_t)i * r + c;  // + a*D*r for corner a
          double x[K];
#pragma unroll
          for (int a = 0; a < K; ++a) x[a] = live ? rr[(size_t)a * D * r] : 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            double v = 0.0, d0 = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
              const double d = sD[le * D * D + j * D + a];
              d0 -= d;
              v = fma(d, x[a + 1], v);
            }
            oJ[(le * B + i * D + j) * RP] = fma(d0, x[0], v);
          }
        } else {
#pragma unroll
          for (int j = 0; j < D; ++j)
            oJ[(le * B + i * D + j) * RP] = live ? raw[(size_t)le * rawrows * r + (size_t)(i * D + j) * r + c] : 0.0;
        }
      }
      double jv[B];
#pragma unroll
      for (int k = 0; k < B; ++k) jv[k] = oJ[(le * B + k) * RP];
      if (blockIdx.y == 0) {
#pragma unroll
        for (int k = 0; k < B; ++k) gacc = fma(jv[k], sP[le * B + k], gacc);
      }
#pragma unroll 1
      for (int rb = 0; rb < B; ++rb) {
        const double* hrow = sH + (le * B + rb) * B;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < B; ++k) v = fma(hrow[k], jv[k], v);
        oY[(le * B + rb) * RP] = v;
      }
    }
  };
  // Hr += sJ^T sY on the warp's block pairs.  Fragment addresses: one base per operand and pair, compile-time
  // offsets per tile, one pointer bump per k-step -- 10 loads + 25 DMMAs + 2 adds per step.
  auto contract = [&](int b) {
    const size_t lane_off = (size_t)(lane & 3) * RP + (lane >> 2);
    const double* bJ = sJ + (size_t)b * KR * RP + lane_off;
    const double* bY = sY + (size_t)b * KR * RP + lane_off;
    const int kstep = 4 * RP;
#pragma unroll
    for (int s = 0; s < RB_PPW; ++s) {
      if (pbi[s] < 0) continue;
      const double* ja = bJ + pbi[s] * (RB_T * 8);
      const double* yb = bY + pbj[s] * (RB_T * 8);
#pragma unroll
      for (int k0 = 0; k0 < KR; k0 += 4, ja += kstep, yb += kstep) {
        double fa[RB_T], fb[RB_T];
#pragma unroll
        for (int x = 0; x < RB_T; ++x) {
          fa[x] = ja[x * 8];
          fb[x] = yb[x * 8];
        }
#pragma unroll
        for (int a = 0; a < RB_T; ++a)
#pragma unroll
          for (int bb = 0; bb < RB_T; ++bb) dmma_m8n8k4(acc[s][a][bb][0], acc[s][a][bb][1], fa[a], fb[bb]);
      }
    }
  };

  // software pipeline: while chunk c is contracted, chunk c+1 is formed from its landed rows and the rows of
  // chunk c+2 are in flight; one CTA barrier per chunk
  unsigned phase[2] = {0u, 0u};
  if (e0 < e1) {
    if (threadIdx.x == 0) issue(e0, 0);
    double small = fetch_small(e0);
    mbar_wait(smem_u32(sBar), phase[0]);
    phase[0] ^= 1u;
    if (threadIdx.x < NSMALL) sSmall[threadIdx.x] = small;
    __syncthreads();
    if (threadIdx.x == 0 && e0 + CE < e1) issue(e0 + CE, 1);
    small = fetch_small(e0 + CE);
    form(e0, 0);
    int b = 0;
    for (int64_t eb = e0; eb < e1; eb += CE, b ^= 1) {
      const bool have_next = eb + CE < e1;
      if (have_next) {
        mbar_wait(smem_u32(sBar + (b ^ 1)), phase[b ^ 1]);
        phase[b ^ 1] ^= 1u;
        if (threadIdx.x < NSMALL) sSmall[(size_t)(b ^ 1) * NSMALL + threadIdx.x] = small;
      }
      __syncthreads();  // chunk eb is fully formed; everybody is done contracting chunk eb - CE and forming eb
      if (have_next) {
        if (threadIdx.x == 0 && eb + 2 * CE < e1) issue(eb + 2 * CE, b);
        small = fetch_small(eb + 2 * CE);
        form(eb + CE, b ^ 1);
      }
      contract(b);
    }
  }
#pragma unroll
  for (int s = 0; s < RB_PPW; ++s) {
    if (pbi[s] < 0) continue;
#pragma unroll
    for (int a = 0; a < RB_T; ++a)
#pragma unroll
      for (int b = 0; b < RB_T; ++b) {
        const int tm = pbi[s] * RB_T + a, tn = pbj[s] * RB_T + b;
        const int row = tm * 8 + (lane >> 2), col = tn * 8 + 2 * (lane & 3);
        if (row < r && col < r) Hpart[((size_t)blockIdx.x * r + row) * r + col] = acc[s][a][b][0];
        if (row < r && col + 1 < r) Hpart[((size_t)blockIdx.x * r + row) * r + col + 1] = acc[s][a][b][1];
      }
  }
  if (blockIdx.y == 0 && threadIdx.x < r) gpart[(size_t)blockIdx.x * r + threadIdx.x] = gacc;
}

// lower triangle := transpose of the upper triangle (the blocked kernel computes block pairs bi <= bj only)
__global__ void mirror_upper_kernel(int r, double* H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= r * r) return;
  const int row = idx / r, col = idx - row * r;
  if (row > col) H[idx] = H[(size_t)col * r + row];
}

// out[i] = sum_{c} part[c][i]   in CTA order
__global__ void sum_partials_kernel(int nparts, int64_t len, const double* part, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int c = 0; c < nparts; ++c) s += part[(size_t)c * len + i];
  out[i] = s;
}

// ---- FST -------------------------------------------------------------------
// grid (m1, n_clusters); threads over q.  Elements of a cluster are visited in
// ascending element order, so each ARBs entry has a fixed summation order.
template <int D>
__global__ void fst_precompute_kernel(int64_t t, int m1, int m2, int ncl, const double* A, const double* Bm,
                                      const int* order, const int* cptr, double* ARBs) {
  constexpr int B = D * D;
  const int p = blockIdx.x, c = blockIdx.y;
  for (int q = threadIdx.x; q < m2; q += blockDim.x) {
    double acc[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) acc[i][j] = 0.0;
    for (int s = cptr[c]; s < cptr[c + 1]; ++s) {
      const int64_t e = order[s];
      double a[B], b[B];
#pragma unroll
      for (int k = 0; k < B; ++k) {
        a[k] = A[(int64_t)p * B * t + e * B + k];
        b[k] = Bm[(e * B + k) * m2 + q];
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
          for (int k = 0; k < D; ++k) acc[i][j] = fma(a[D * i + k], b[D * j + k], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) ARBs[((((int64_t)p * m2 + q) * ncl + c) * D + i) * D + j] = acc[i][j];
  }
}

// Register-tiled form (default): a CTA owns PT rows p of one cluster and 128 columns q; the PT x D*D slice of A of
// FST_CH elements at a time is staged in shared memory (read back as broadcasts), a thread keeps PT x D x D accumulators
// for its column q and loads each element's D*D values of B once for all PT rows: PT*D^3 FMAs per D*D global loads
// instead of D^3 per 2 D*D.  Same summation order per entry as fst_precompute_kernel => bit-identical values.
// (DMMA tiles would not help: the DMMA peak of this GPU is the FMA peak)
constexpr int FST_CH = 16;
template <int D, int PT>
__global__ void __launch_bounds__(128) fst_precompute_tiled_kernel(int64_t t, int m1, int m2, int ncl, const double* __restrict__ A,
                                                                   const double* __restrict__ Bm, const int* __restrict__ order,
                                                                   const int* __restrict__ cptr, double* __restrict__ ARBs) {
  constexpr int B = D * D;
  __shared__ double sa[FST_CH][PT][B];
  const int p0 = blockIdx.x * PT, c = blockIdx.y;
  const int q = blockIdx.z * blockDim.x + threadIdx.x;
  double acc[PT][D][D];
#pragma unroll
  for (int pp = 0; pp < PT; ++pp)
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) acc[pp][i][j] = 0.0;
  const int s_end = cptr[c + 1];
  for (int s0 = cptr[c]; s0 < s_end; s0 += FST_CH) {
    const int n = min(FST_CH, s_end - s0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * PT * B; idx += blockDim.x) {
      const int se = idx / (PT * B), rem = idx - se * (PT * B);
      const int pp = rem / B, k = rem - pp * B;
      const int p = p0 + pp;
      sa[se][pp][k] = (p < m1) ? A[(int64_t)p * B * t + (int64_t)order[s0 + se] * B + k] : 0.0;
    }
    __syncthreads();
    if (q < m2) {
      for (int se = 0; se < n; ++se) {
        const int64_t e = order[s0 + se];
        double b[B];
#pragma unroll
        for (int k = 0; k < B; ++k) b[k] = Bm[(e * B + k) * m2 + q];
#pragma unroll
        for (int pp = 0; pp < PT; ++pp)
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j)
#pragma unroll
              for (int k = 0; k < D; ++k) acc[pp][i][j] = fma(sa[se][pp][D * i + k], b[D * j + k], acc[pp][i][j]);
      }
    }
  }
  if (q < m2) {
#pragma unroll
    for (int pp = 0; pp < PT; ++pp) {
      if (p0 + pp >= m1) break;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) ARBs[((((int64_t)(p0 + pp) * m2 + q) * ncl + c) * D + i) * D + j] = acc[pp][i][j];
    }
  }
}

template <int D>
static int reduced_run(skb_plan* pl, int material, int psd_mode, int64_t t, int64_t r, const double* JB_h,
                       const double* Jx0_h, const double* B_h, const double* x0_h, const double* z_h,
                       const double* mu_h, int64_t mu_n, const double* lam_h, int64_t lam_n, const double* vol_h,
                       int64_t vol_n, double* energy, double* gr, double* Hr) {
  constexpr int Bk = D * D;
  if (r <= 0 || r > RH_THREADS) return fail(SKB_EINVAL, "reduced dimension r must be in [1, 256]");
  if (material < 0 || material >= MAT_COUNT) return fail(SKB_EINVAL, "unknown material id");
  SKB_TRY
  cudaStream_t st = 0;
  // per-element work arrays: the plan's (kept between calls) when there is a plan
  dvec<double> F_own, He_own, Pw_own, psi_own;
  static int keep_env = -1;   // SKB_REDUCED_KEEP=0: allocate per call (A/B)
  if (keep_env < 0) {
    const char* ev = getenv("SKB_REDUCED_KEEP");
    keep_env = (ev && strcmp(ev, "0") == 0) ? 0 : 1;
  }
  const bool keep = pl && keep_env;
  dvec<double>& F = keep ? pl->rw_F : F_own;
  dvec<double>& He = keep ? pl->rw_He : He_own;
  dvec<double>& Pw = keep ? pl->rw_Pw : Pw_own;
  dvec<double>& psi = keep ? pl->rw_psi : psi_own;
  if (F.size() != (size_t)t * Bk) F.resize((size_t)t * Bk);
  if (He.size() != (size_t)t * Bk * Bk) He.resize((size_t)t * Bk * Bk);
  if (Pw.size() != (size_t)t * Bk) Pw.resize((size_t)t * Bk);
  if (psi.size() != (size_t)t) psi.resize(t);
  dvec<double> z(z_h, z_h + r);
  dvec<double> JB, Bm, mu, lam, vol;
  const double *mu_p, *lam_p = nullptr, *vol_p;
  const double* Bm_p = nullptr;
  int mu_s, lam_s = 0, vol_s;
  if (pl) {
    // F = J (B z + x0) through the mesh plan
    const int64_t nd = pl->ndof();
    if (B_h) Bm.assign(B_h, B_h + nd * r);  // else: the plan's resident basis (skb_plan_set_basis)
    dvec<double>& x = pl->rw_x;
    if (x.size() != (size_t)nd) x.resize(nd);
    dvec<double> x0;
    if (x0_h) x0.assign(x0_h, x0_h + nd);
    Bm_p = B_h ? raw(Bm) : raw(pl->basis);
    gemv_rows_kernel<<<(unsigned)((nd * 32 + 255) / 256), 256, 0, st>>>(nd, (int)r, Bm_p, raw(z), x0_h ? raw(x0) : nullptr, raw(x));
    // F_e from x: same gather as load_element in kernels.cuh
    const PlanView p = pl->view();
    thrust::counting_iterator<int> it0(0);
    const double* xp = raw(x);
    double* Fp = raw(F);
    thrust::for_each(thrust::cuda::par.on(st), it0, it0 + (int)t, [=] __device__(int e) {
      Mat<D> Fm;
      double Dm[D][D];
      constexpr int K = D + 1;
      const int* Te = p.T32 + (size_t)e * K;
      for (int j = 0; j < D; ++j)
        for (int c = 0; c < D; ++c) Dm[j][c] = p.Dm[(size_t)(j * D + c) * p.t + e];
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Fm.m[i][j] = 0.0;
      for (int c = 0; c < D; ++c)
        for (int i = 0; i < D; ++i) {
          double d = xp[(size_t)Te[c + 1] * D + i] - xp[(size_t)Te[0] * D + i];
          for (int j = 0; j < D; ++j) Fm.m[i][j] = fma(Dm[j][c], d, Fm.m[i][j]);
        }
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Fp[(size_t)e * D * D + i * D + j] = Fm.m[i][j];
    });
    if (!pl->have_materials) return fail(SKB_EINVAL, "materials not set (skb_set_materials)");
    mu_p = raw(pl->mu); mu_s = pl->mu_n > 1;
    lam_p = raw(pl->lam); lam_s = pl->lam_n > 1;
    vol_p = raw(pl->vol); vol_s = pl->vol_n > 1;
  } else {
    if (!mu_h || (mu_n != 1 && mu_n != t)) return fail(SKB_EINVAL, "mu must have 1 or t entries");
    if (lam_h && lam_n != 1 && lam_n != t) return fail(SKB_EINVAL, "lam must have 1 or t entries");
    if (!vol_h || (vol_n != 1 && vol_n != t)) return fail(SKB_EINVAL, "vol must have 1 or t entries");
    JB.assign(JB_h, JB_h + t * Bk * r);
    dvec<double> Jx0;
    if (Jx0_h) Jx0.assign(Jx0_h, Jx0_h + t * Bk);
    gemv_rows_kernel<<<(unsigned)((t * Bk * 32 + 255) / 256), 256, 0, st>>>(t * Bk, (int)r, raw(JB), raw(z), Jx0_h ? raw(Jx0) : nullptr, raw(F));
    mu.assign(mu_h, mu_h + mu_n);
    if (lam_h) lam.assign(lam_h, lam_h + lam_n);
    vol.assign(vol_h, vol_h + vol_n);
    mu_p = raw(mu); mu_s = mu_n > 1;
    if (lam_h) { lam_p = raw(lam); lam_s = lam_n > 1; }
    vol_p = raw(vol); vol_s = vol_n > 1;
  }
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) SKB_CUDA(cudaEventCreate(&ev[i]));
  SKB_CUDA(cudaEventRecord(ev[0], st));
  reduced_pass1_kernel<D><<<(unsigned)((t + 127) / 128), 128, 0, st>>>(material, psd_mode, t, raw(F), mu_p, mu_s, lam_p, lam_s,
                                                                       vol_p, vol_s, raw(He), raw(Pw), raw(psi));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaEventRecord(ev[1], st));
  SKB_CUDA(cudaEventRecord(ev[2], st));
  SKB_CUDA(cudaEventRecord(ev[3], st));
  if (energy) {
    // fixed-order two-stage sum
    const int nb = 1024;
    dvec<double> part(nb), out(1);
    const double* ps = raw(psi);
    double* pp = raw(part);
    const int64_t per = (t + nb - 1) / nb;
    thrust::counting_iterator<int> it0(0);
    thrust::for_each(thrust::cuda::par.on(st), it0, it0 + nb, [=] __device__(int b) {
      double s = 0.0;
      for (int64_t e = b * per; e < (b + 1) * per && e < t; ++e) s += ps[e];
      pp[b] = s;
    });
    reduce_final_kernel<<<1, 1024, 0, st>>>(raw(part), nb, raw(out));
    SKB_CUDA(cudaMemcpy(energy, raw(out), sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (gr || Hr) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms;
    const int64_t chunks = (t + RH_ELEMS - 1) / RH_ELEMS;
    if (grid > chunks) grid = (int)chunks;
    const int rt = (int)(r + 7) / 8;
    const int nwarps = RH_THREADS / 32;
    int npanels = 1, tcp = rt;
    while ((rt * tcp + nwarps - 1) / nwarps > RH_MAXT) {
      ++npanels;
      tcp = (rt + npanels - 1) / npanels;
    }
    npanels = (rt + tcp - 1) / tcp;
    const int rpad = rt * 8 + 1;
    const size_t smem = (size_t)2 * RH_ELEMS * Bk * rpad * sizeof(double);
    SKB_CUDA(cudaFuncSetAttribute(reduced_pass2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dvec<double> Hpart((size_t)grid * r * r), gpart((size_t)grid * r), Hd(r * r), gd(r);
    PlanView pv;
    memset(&pv, 0, sizeof(pv));
    if (pl) pv = pl->view();
    // register-blocked, TMA-staged kernel when every operator row is a 16-byte multiple and it fits
    using RS = RBShape<D>;
    const int rawrows = pl ? RS::RAWROWS : Bk;
    const size_t bsmem = ((size_t)4 * RS::KR * rb_stride((int)r) + (size_t)2 * RS::CE * rawrows * r +
                          2 * RS::CE * (Bk * Bk + Bk + D * D)) * sizeof(double) + 16;
    static int blocked_env = -1;
    if (blocked_env < 0) {
      const char* evv = getenv("SKB_REDUCED");
      blocked_env = (evv && strcmp(evv, "simple") == 0) ? 0 : 1;
    }
    const bool blocked = blocked_env && (r % 2 == 0) && bsmem <= 227 * 1024;
    SKB_CUDA(cudaEventRecord(ev[2], st));
    if (blocked) {
      const int nblk = (rt + RB_T - 1) / RB_T;
      const int npasses = (nblk * (nblk + 1) / 2 + RB_MAXPAIRS - 1) / RB_MAXPAIRS;
      SKB_CUDA(cudaFuncSetAttribute(reduced_pass2_blocked_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SKB_CUDA(cudaMemsetAsync(raw(Hpart), 0, Hpart.size() * sizeof(double), st));
      reduced_pass2_blocked_kernel<D><<<dim3(grid, npasses), RB_THREADS, bsmem, st>>>(
          t, (int)r, pl ? nullptr : raw(JB), pv, pl ? 1 : 0, Bm_p, raw(He), raw(Pw), raw(Hpart), raw(gpart));
    } else {
      reduced_pass2_kernel<D><<<dim3(grid, npanels), RH_THREADS, smem, st>>>(t, (int)r, tcp, pl ? nullptr : raw(JB), pv, pl ? 1 : 0,
                                                             Bm_p, raw(He), raw(Pw), raw(Hpart), raw(gpart));
    }
    SKB_CUDA(cudaGetLastError());
    SKB_CUDA(cudaEventRecord(ev[3], st));
    sum_partials_kernel<<<(unsigned)((r * r + 255) / 256), 256, 0, st>>>(grid, r * r, raw(Hpart), raw(Hd));
    if (blocked) mirror_upper_kernel<<<(unsigned)((r * r + 255) / 256), 256, 0, st>>>((int)r, raw(Hd));
    sum_partials_kernel<<<(unsigned)((r + 255) / 256), 256, 0, st>>>(grid, r, raw(gpart), raw(gd));
    SKB_CUDA(cudaDeviceSynchronize());
    if (Hr) SKB_CUDA(cudaMemcpy(Hr, raw(Hd), r * r * sizeof(double), cudaMemcpyDeviceToHost));
    if (gr) SKB_CUDA(cudaMemcpy(gr, raw(gd), r * sizeof(double), cudaMemcpyDeviceToHost));
  }
  SKB_CUDA(cudaDeviceSynchronize());
  {
    float a = 0.f, b = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[2], ev[3]);
    cudaEventElapsedTime(&c, ev[0], ev[3]);
    g_reduced_ms[0] = a;
    g_reduced_ms[1] = b;
    g_reduced_ms[2] = c;
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  }
  return SKB_OK;
  SKB_CATCH
}

}  // namespace skb

using namespace skb;

extern "C" {

int skb_reduced_gradient_hessian(int material, int psd_mode, int dim, int64_t t, int64_t r, const double* JB,
                                 const double* Jx0, const double* z, const double* mu, int64_t mu_n,
                                 const double* lam, int64_t lam_n, const double* vol, int64_t vol_n,
                                 double* energy, double* gr, double* Hr) {
  if (!JB || !z) return fail(SKB_EINVAL, "null argument");
  if (dim != 2 && dim != 3) return fail(SKB_EINVAL, "Only dim == 2 or 3 are supported");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  return dim == 3 ? reduced_run<3>(nullptr, material, psd_mode, t, r, JB, Jx0, nullptr, nullptr, z, mu, mu_n, lam, lam_n, vol, vol_n, energy, gr, Hr)
                  : reduced_run<2>(nullptr, material, psd_mode, t, r, JB, Jx0, nullptr, nullptr, z, mu, mu_n, lam, lam_n, vol, vol_n, energy, gr, Hr);
}

int skb_reduced_last_times(double out[3]) {
  if (!out) return fail(SKB_EINVAL, "null argument");
  for (int i = 0; i < 3; ++i) out[i] = g_reduced_ms[i];
  return SKB_OK;
}

int skb_plan_set_basis(skb_plan* pl, int64_t r, const double* B) {
  if (!pl) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  if (!B || r <= 0) {
    pl->basis.clear();
    pl->basis.shrink_to_fit();
    pl->basis_r = 0;
    for (dvec<double>* w : {&pl->rw_F, &pl->rw_He, &pl->rw_Pw, &pl->rw_psi, &pl->rw_x}) {   // and the work arrays
      w->clear();
      w->shrink_to_fit();
    }
    return SKB_OK;
  }
  const int64_t nd = pl->ndof();
  pl->basis.assign(B, B + nd * r);
  pl->basis_r = r;
  return SKB_OK;
  SKB_CATCH
}

int skb_reduced_hessian_from_basis(skb_plan* pl, int material, int psd_mode, int64_t r, const double* B,
                                   const double* x0, const double* z, double* energy, double* gr, double* Hr) {
  if (!pl || !z) return fail(SKB_EINVAL, "null argument");
  if (!B && (pl->basis_r != r || pl->basis_r == 0))
    return fail(SKB_EINVAL, "B is NULL and the plan holds no resident basis of this dimension (skb_plan_set_basis)");
  SKB_CUDA(cudaSetDevice(pl->device));
  // the elements that count in sums over elements: all of them, or a shard's own (which the plan lists first) when it
  // also evaluates its lower neighbour's interface elements -- every element then enters the all-reduced sums once
  const int64_t t = pl->d.t_energy;
  return pl->d.dim == 3 ? reduced_run<3>(pl, material, psd_mode, t, r, nullptr, nullptr, B, x0, z, nullptr, 0, nullptr, 0, nullptr, 0, energy, gr, Hr)
                        : reduced_run<2>(pl, material, psd_mode, t, r, nullptr, nullptr, B, x0, z, nullptr, 0, nullptr, 0, nullptr, 0, energy, gr, Hr);
}

int skb_fst_precompute(int dim, int64_t t, int64_t m1, int64_t m2, int64_t ncl, const double* A, const double* B,
                       const int32_t* l, double* ARBs) {
  if (!A || !B || !l || !ARBs) return fail(SKB_EINVAL, "null argument");
  if (dim != 2 && dim != 3) return fail(SKB_EINVAL, "Only dim == 2 or 3 are supported");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  if (m1 <= 0 || m2 <= 0 || ncl <= 0 || t <= 0 || ncl > 65535) return fail(SKB_EINVAL, "bad sizes");
  SKB_TRY
  const int b = dim * dim;
  // cluster -> ascending element list (stable counting sort on the host: one-off setup)
  std::vector<int> cptr(ncl + 1, 0), order(t);
  for (int64_t e = 0; e < t; ++e) {
    if (l[e] < 0 || l[e] >= ncl) return fail(SKB_EINVAL, "cluster label out of range");
    cptr[l[e] + 1]++;
  }
  for (int64_t c = 0; c < ncl; ++c) cptr[c + 1] += cptr[c];
  {
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t e = 0; e < t; ++e) order[fill[l[e]]++] = (int)e;
  }
  dvec<double> Ad(A, A + m1 * b * t), Bd(B, B + b * t * m2), out((size_t)m1 * m2 * ncl * b);
  dvec<int> od(order.begin(), order.end()), cd(cptr.begin(), cptr.end());
  static int fst_simple = -1;   // SKB_FST=simple: the one-row-per-CTA kernel (A/B and the bit-identity test)
  if (fst_simple < 0) {
    const char* ev = getenv("SKB_FST");
    fst_simple = (ev && strcmp(ev, "simple") == 0) ? 1 : 0;
  }
  if (fst_simple) {
    dim3 grid((unsigned)m1, (unsigned)ncl);
    const int threads = m2 >= 128 ? 128 : (int)((m2 + 31) / 32 * 32);
    if (dim == 3)
      fst_precompute_kernel<3><<<grid, threads>>>(t, (int)m1, (int)m2, (int)ncl, raw(Ad), raw(Bd), raw(od), raw(cd), raw(out));
    else
      fst_precompute_kernel<2><<<grid, threads>>>(t, (int)m1, (int)m2, (int)ncl, raw(Ad), raw(Bd), raw(od), raw(cd), raw(out));
  } else {
    constexpr int PT = 8;
    dim3 grid((unsigned)((m1 + PT - 1) / PT), (unsigned)ncl, (unsigned)((m2 + 127) / 128));
    if (dim == 3)
      fst_precompute_tiled_kernel<3, PT><<<grid, 128>>>(t, (int)m1, (int)m2, (int)ncl, raw(Ad), raw(Bd), raw(od), raw(cd), raw(out));
    else
      fst_precompute_tiled_kernel<2, PT><<<grid, 128>>>(t, (int)m1, (int)m2, (int)ncl, raw(Ad), raw(Bd), raw(od), raw(cd), raw(out));
  }
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  SKB_CUDA(cudaMemcpy(ARBs, raw(out), out.size() * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

int skb_fst_eval(int dim, int64_t m1, int64_t m2, int64_t ncl, const double* ARBs, const double* r, double* out) {
  if (!ARBs || !r || !out) return fail(SKB_EINVAL, "null argument");
  if (dim != 2 && dim != 3) return fail(SKB_EINVAL, "Only dim == 2 or 3 are supported");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  const int64_t rows = m1 * m2, cols = ncl * dim * dim;
  dvec<double> Ad(ARBs, ARBs + rows * cols), rd(r, r + cols), od(rows);
  gemv_rows_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256>>>(rows, (int)cols, raw(Ad), raw(rd), nullptr, raw(od));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  SKB_CUDA(cudaMemcpy(out, raw(od), rows * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"
