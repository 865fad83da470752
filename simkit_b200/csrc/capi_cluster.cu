// C ABI, part 10: spectral clustering / cubature of the subspace construction:
//   average_onto_simplex.py:8-37      per-simplex mean of a per-vertex basis
//   spectral_clustering.py:9-48       scipy.cluster.vq.kmeans2(B, k, seed=seed, minit="++")
//   spectral_cubature.py:60-66        nearest element of every centroid, cluster volumes
//
// kmeans2 is restated from scipy 1.16 (scipy/cluster/vq: _kpp + the Lloyd loop of kmeans2): k-means++ seeding -- first
// centre = a uniformly drawn row, every further centre drawn with probability proportional to the squared distance to
// the nearest centre so far (cumulative sum + searchsorted of one uniform number) -- then `iter` rounds of nearest-
// centre assignment (ties to the lowest index) and centre = mean of its members (an empty cluster keeps its centre).
// The random numbers come from the caller, who draws them from the very generator scipy would use, in scipy's order
// (one integer, then k - 1 uniforms), so labels agree with the reference's exactly unless a draw lands within rounding
// of a cumulative-probability boundary.  Everything per row (distances, minima, assignment, means) runs on the GPU with
// fixed-shape reductions: results are reproducible run to run.
#include "capi_common.cuh"

#include <thrust/binary_search.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

using namespace skb;

namespace {

constexpr int KM_BLOCK = 1024;   // rows per block of the seeding pass (one partial sum each)

// fixed-shape block sum (blockDim.x a multiple of 32, <= 1024); valid in thread 0
__device__ __forceinline__ double km_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = ((int)threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;
}

__global__ void average_onto_simplex_kernel(int64_t t, int p, int K, const double* __restrict__ A, const int32_t* __restrict__ T,
                                            double* __restrict__ At) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t * p) return;
  const int64_t e = i / p;
  const int f = (int)(i - e * p);
  double acc = 0.0;   // the reference adds A[T[:, c]] / K corner after corner (average_onto_simplex.py:35-36)
  for (int c = 0; c < K; ++c) acc += A[(size_t)T[e * K + c] * p + f] / (double)K;
  At[i] = acc;
}

// seeding: d2min[i] = min(d2min[i], |x_i - c|^2) for the centre row `c`; bsum[b] = sum of d2min over block b's rows
__global__ void __launch_bounds__(KM_BLOCK) kpp_update_kernel(int64_t n, int p, const double* __restrict__ data, int64_t c, int first,
                                                              double* __restrict__ d2min, double* __restrict__ bsum) {
  __shared__ double sh[32];
  const int64_t i = (int64_t)blockIdx.x * KM_BLOCK + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    double d2 = 0.0;
    for (int f = 0; f < p; ++f) {
      const double d = data[(size_t)i * p + f] - data[(size_t)c * p + f];
      d2 = fma(d, d, d2);
    }
    v = first ? d2 : fmin(d2min[i], d2);
    d2min[i] = v;
  }
  v = km_block_sum(v, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = v;
}

// assignment: nearest centre of every row, ties to the lowest index (scipy vq)
__global__ void assign_kernel(int64_t n, int p, int k, const double* __restrict__ data, const double* __restrict__ cen,
                              int32_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* x = data + (size_t)i * p;
  double best = 1e308 * 10.0;
  int bj = 0;
  for (int j = 0; j < k; ++j) {
    const double* c = cen + (size_t)j * p;
    double d2 = 0.0;
    for (int f = 0; f < p; ++f) {
      const double d = x[f] - c[f];
      d2 = fma(d, d, d2);
    }
    if (d2 < best) {
      best = d2;
      bj = j;
    }
  }
  labels[i] = bj;
}

// centre j = mean of its members (rows order[off[j] .. off[j+1]), ascending row index); empty: unchanged, flagged
__global__ void cluster_mean_kernel(int p, const double* __restrict__ data, const int32_t* __restrict__ order,
                                    const int32_t* __restrict__ off, double* __restrict__ cen, int32_t* __restrict__ n_empty) {
  __shared__ double sh[32];
  const int j = blockIdx.x;
  const int a = off[j], b = off[j + 1];
  if (a == b) {
    if (threadIdx.x == 0) atomicAdd(n_empty, 1);
    return;
  }
  for (int f = 0; f < p; ++f) {
    double v = 0.0;
    for (int m = a + threadIdx.x; m < b; m += blockDim.x) v += data[(size_t)order[m] * p + f];
    v = km_block_sum(v, sh);
    if (threadIdx.x == 0) cen[(size_t)j * p + f] = v / (double)(b - a);
  }
}

// cluster volume = sum of the weights of its members; nearest row of centre j over ALL rows (first minimum)
__global__ void cluster_weight_kernel(const double* __restrict__ w, const int32_t* __restrict__ order, const int32_t* __restrict__ off,
                                      double* __restrict__ mc) {
  __shared__ double sh[32];
  const int j = blockIdx.x;
  double v = 0.0;
  for (int m = off[j] + threadIdx.x; m < off[j + 1]; m += blockDim.x) v += w[order[m]];
  v = km_block_sum(v, sh);
  if (threadIdx.x == 0) mc[j] = v;
}

__global__ void nearest_row_kernel(int64_t n, int p, const double* __restrict__ data, const double* __restrict__ cen,
                                   int64_t* __restrict__ out) {
  __shared__ double sd[1024];
  __shared__ long long si[1024];
  const int j = blockIdx.x;
  const double* c = cen + (size_t)j * p;
  double best = 1e308 * 10.0;
  long long bi = -1;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {   // ascending i per thread: the first minimum is kept
    double d2 = 0.0;
    for (int f = 0; f < p; ++f) {
      const double d = data[(size_t)i * p + f] - c[f];
      d2 = fma(d, d, d2);
    }
    if (d2 < best) {
      best = d2;
      bi = i;
    }
  }
  sd[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const double od = sd[threadIdx.x + s];
      const long long oi = si[threadIdx.x + s];
      if (oi >= 0 && (si[threadIdx.x] < 0 || od < sd[threadIdx.x] || (od == sd[threadIdx.x] && oi < si[threadIdx.x]))) {
        sd[threadIdx.x] = od;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[j] = si[0];
}

// rows grouped by label, ascending row index inside a group
int group_by_label(int64_t n, int k, const dvec<int32_t>& labels, dvec<int32_t>& order, dvec<int32_t>& off) {
  dvec<int32_t> keys = labels;
  order.resize(n);
  thrust::sequence(order.begin(), order.end());
  thrust::stable_sort_by_key(keys.begin(), keys.end(), order.begin());
  off.resize(k + 1);
  thrust::counting_iterator<int32_t> c0(0);
  thrust::lower_bound(keys.begin(), keys.end(), c0, c0 + (k + 1), off.begin());
  return SKB_OK;
}

}  // namespace

extern "C" {

int skb_average_onto_simplex(int64_t n, int64_t p, int64_t t, int K, const double* A, const int32_t* T, double* At) {
  if (!A || !T || !At) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || p <= 0 || t <= 0 || K <= 0 || t * p >= ((int64_t)1 << 40)) return fail(SKB_EINVAL, "bad sizes");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  dvec<double> dA(A, A + n * p), dAt((size_t)t * p);
  dvec<int32_t> dT(T, T + t * K);
  const int64_t items = t * p;
  average_onto_simplex_kernel<<<(unsigned)((items + 255) / 256), 256>>>(t, (int)p, K, raw(dA), raw(dT), raw(dAt));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaMemcpy(At, raw(dAt), sizeof(double) * items, cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

int skb_kmeans2_pp(int64_t n, int64_t p, int64_t k, int iters, int64_t first, const double* uniforms, const double* data,
                   double* centroids, int32_t* labels, int32_t* n_empty) {
  if (!data || !centroids || !labels || (k > 1 && !uniforms)) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || p <= 0 || k <= 0 || iters < 1 || first < 0 || first >= n || n >= ((int64_t)1 << 31) || k >= (1 << 24))
    return fail(SKB_EINVAL, "bad sizes");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  dvec<double> dX(data, data + n * p), dC((size_t)k * p), d2((size_t)n);
  const int nb = (int)((n + KM_BLOCK - 1) / KM_BLOCK);
  dvec<double> bsum(nb);
  std::vector<double> hb(nb), hblk(KM_BLOCK);
  // ---- k-means++ seeding (scipy _kpp)
  int64_t c = first;
  SKB_CUDA(cudaMemcpy(raw(dC), raw(dX) + (size_t)c * p, sizeof(double) * p, cudaMemcpyDeviceToDevice));
  for (int64_t i = 1; i < k; ++i) {
    kpp_update_kernel<<<nb, KM_BLOCK>>>(n, (int)p, raw(dX), c, i == 1 ? 1 : 0, raw(d2), raw(bsum));
    SKB_CUDA(cudaGetLastError());
    SKB_CUDA(cudaMemcpy(hb.data(), raw(bsum), sizeof(double) * nb, cudaMemcpyDeviceToHost));
    double total = 0.0;
    for (int b = 0; b < nb; ++b) total += hb[b];
    // first row whose cumulative probability reaches the drawn number (numpy searchsorted, side='left')
    const double target = uniforms[i - 1] * total;
    double cum = 0.0;
    int b = 0;
    while (b < nb - 1 && cum + hb[b] < target) cum += hb[b++];
    const int64_t r0 = (int64_t)b * KM_BLOCK, cnt = (n - r0 < KM_BLOCK) ? n - r0 : KM_BLOCK;
    SKB_CUDA(cudaMemcpy(hblk.data(), raw(d2) + r0, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    int64_t pick = r0 + cnt - 1;
    for (int64_t m = 0; m < cnt; ++m) {
      cum += hblk[m];
      if (cum >= target) {
        pick = r0 + m;
        break;
      }
    }
    c = pick;
    SKB_CUDA(cudaMemcpy(raw(dC) + (size_t)i * p, raw(dX) + (size_t)c * p, sizeof(double) * p, cudaMemcpyDeviceToDevice));
  }
  // ---- Lloyd rounds
  dvec<int32_t> dL((size_t)n), order, off, dEmpty(1, 0);
  for (int it = 0; it < iters; ++it) {
    assign_kernel<<<(unsigned)((n + 127) / 128), 128>>>(n, (int)p, (int)k, raw(dX), raw(dC), raw(dL));
    SKB_CUDA(cudaGetLastError());
    group_by_label(n, (int)k, dL, order, off);
    cluster_mean_kernel<<<(unsigned)k, 256>>>((int)p, raw(dX), raw(order), raw(off), raw(dC), raw(dEmpty));
    SKB_CUDA(cudaGetLastError());
  }
  SKB_CUDA(cudaMemcpy(centroids, raw(dC), sizeof(double) * k * p, cudaMemcpyDeviceToHost));
  SKB_CUDA(cudaMemcpy(labels, raw(dL), sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  if (n_empty) SKB_CUDA(cudaMemcpy(n_empty, raw(dEmpty), sizeof(int32_t), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

int skb_cubature_pick(int64_t n, int64_t p, int64_t k, const double* data, const double* centroids, const int32_t* labels,
                      const double* vol, int64_t* lI, double* mc) {
  if (!data || !centroids || !labels || !vol || !lI || !mc) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || p <= 0 || k <= 0 || n >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "bad sizes");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  for (int64_t i = 0; i < n; ++i)
    if (labels[i] < 0 || labels[i] >= k) return fail(SKB_EINVAL, "label out of range");
  SKB_TRY
  dvec<double> dX(data, data + n * p), dC(centroids, centroids + k * p), dV(vol, vol + n), dM((size_t)k);
  dvec<int32_t> dL(labels, labels + n), order, off;
  dvec<int64_t> dI((size_t)k);
  group_by_label(n, (int)k, dL, order, off);
  cluster_weight_kernel<<<(unsigned)k, 256>>>(raw(dV), raw(order), raw(off), raw(dM));
  SKB_CUDA(cudaGetLastError());
  nearest_row_kernel<<<(unsigned)k, 1024>>>(n, (int)p, raw(dX), raw(dC), raw(dI));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaMemcpy(lI, raw(dI), sizeof(int64_t) * k, cudaMemcpyDeviceToHost));
  SKB_CUDA(cudaMemcpy(mc, raw(dM), sizeof(double) * k, cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"
