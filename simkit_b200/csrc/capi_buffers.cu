// C ABI, part 8: device-resident value buffers behind the lazy Hessian of the drop-in `*_hessian_x` functions
// (simkit_b200/device_csr.py).
//
// In the reference every `*_hessian_x` returns a host scipy matrix (e.g. energies/stable_neo_hookean.py:506-538), the
// caller adds its mass / penalty / contact matrices to it on the host (integrators/backward_euler.py:84,
// examples/interactive_demos/010_interactive_contact_plane_3D.py:104-108) and hands the sum to spsolve
// (solvers/newton.py:52).  Here the 8*nnz bytes of values (2.9 GB at 16 M tets) stay in HBM through that whole chain:
// these entry points are the few vector operations the chain needs on them.  A buffer is plain cudaMalloc memory.
#include "capi_common.cuh"

using namespace skb;

namespace skb {
static __global__ void buf_axpy_kernel(double* dst, double a, const double* src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = fma(a, src[i], dst[i]);
}
static __global__ void buf_scale_kernel(double* dst, double a, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] *= a;
}
// dst[pos[i]] += a * v[i]; the positions are distinct (the host sums duplicates first), so no atomics are needed
static __global__ void buf_index_add_kernel(double* dst, const int32_t* pos, const double* v, double a, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && pos[i] >= 0) dst[pos[i]] = fma(a, v[i], dst[pos[i]]);
}
// vals[position of (i, i)] += a * diag[i]: the position comes from a binary search of the block row (csr_value_position)
template <int D>
static __global__ void buf_add_diagonal_kernel(const int* bptr, const int* bcol, int64_t nd, const double* diag, double a,
                                               double* vals, int* missing) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nd) return;
  const double d = diag[i];
  if (d == 0.0) return;
  const int pos = csr_value_position<D>(bptr, bcol, (int)i, (int)i);
  if (pos < 0) {
    *missing = 1;   // benign race: every writer stores the same value
    return;
  }
  vals[pos] = fma(a, d, vals[pos]);
}
static inline unsigned buf_grid(int64_t n) {
  const int64_t g = (n + 255) / 256;
  return (unsigned)(g < 148 * 16 ? (g > 0 ? g : 1) : 148 * 16);
}
}  // namespace skb

extern "C" {

int skb_buf_alloc(int device, int64_t n, double** out) {
  if (!out || n <= 0) return fail(SKB_EINVAL, "bad argument");
  if (skb_device_count() <= device) return fail(SKB_ENOGPU, "no CUDA device " + std::to_string(device));
  SKB_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)n * sizeof(double));
  if (e != cudaSuccess) return fail(SKB_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  *out = static_cast<double*>(p);
  return SKB_OK;
}

void skb_buf_free(int device, double* buf) {
  if (!buf) return;
  cudaSetDevice(device);
  cudaFree(buf);
}

int skb_buf_copy(int device, double* dst, const double* src, int64_t n) {
  if (!dst || !src || n < 0) return fail(SKB_EINVAL, "bad argument");
  SKB_CUDA(cudaSetDevice(device));
  SKB_CUDA(cudaMemcpy(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice));
  return SKB_OK;
}

int skb_buf_axpy(int device, double* dst, double a, const double* src, int64_t n) {
  if (!dst || !src || n < 0) return fail(SKB_EINVAL, "bad argument");
  SKB_CUDA(cudaSetDevice(device));
  buf_axpy_kernel<<<buf_grid(n), 256>>>(dst, a, src, n);
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  return SKB_OK;
}

int skb_buf_scale(int device, double* dst, double a, int64_t n) {
  if (!dst || n < 0) return fail(SKB_EINVAL, "bad argument");
  SKB_CUDA(cudaSetDevice(device));
  buf_scale_kernel<<<buf_grid(n), 256>>>(dst, a, n);
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  return SKB_OK;
}

// dst[pos[i]] += a * vals[i] for host arrays pos (int32, distinct, -1 = skip) and vals of `count` entries
int skb_buf_index_add(int device, double* dst, const int32_t* pos, const double* vals, int64_t count, double a) {
  if (!dst || count < 0 || (count > 0 && (!pos || !vals))) return fail(SKB_EINVAL, "bad argument");
  if (count == 0) return SKB_OK;
  SKB_CUDA(cudaSetDevice(device));
  SKB_TRY
  dvec<int32_t> p(pos, pos + count);
  dvec<double> v(vals, vals + count);
  buf_index_add_kernel<<<(unsigned)((count + 255) / 256), 256>>>(dst, raw(p), raw(v), a, count);
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaDeviceSynchronize());
  return SKB_OK;
  SKB_CATCH
}

// vals_dev (the plan's CSR value layout) += a * diag(d) for a HOST vector d of n*dim entries: the lumped mass, Dirichlet
// penalty or any other diagonal term a caller adds to the Hessian (integrators/backward_euler.py:84).  Returns
// SKB_EINVAL if a non-zero diagonal entry has no slot (a vertex no element references).
int skb_buf_add_diagonal(skb_plan* pl, double* vals_dev, const double* diag, double a) {
  if (!pl || !vals_dev || !diag) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  const int64_t nd = pl->ndof();
  pl->w_diag.resize(nd);
  dvec<int> missing(1, 0);
  SKB_CUDA(cudaMemcpyAsync(raw(pl->w_diag), diag, nd * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  const unsigned grid = (unsigned)((nd + 255) / 256);
  if (pl->d.dim == 3)
    buf_add_diagonal_kernel<3><<<grid, 256, 0, pl->stream>>>(raw(pl->d.bptr), raw(pl->d.bcol), nd, raw(pl->w_diag), a, vals_dev, raw(missing));
  else
    buf_add_diagonal_kernel<2><<<grid, 256, 0, pl->stream>>>(raw(pl->d.bptr), raw(pl->d.bcol), nd, raw(pl->w_diag), a, vals_dev, raw(missing));
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  if ((int)missing[0] != 0) return fail(SKB_EINVAL, "a non-zero diagonal entry lies outside the mesh's CSR pattern");
  return SKB_OK;
  SKB_CATCH
}

int skb_buf_download(int device, const double* src, int64_t n, double* host) {
  if (!src || !host || n < 0) return fail(SKB_EINVAL, "bad argument");
  SKB_CUDA(cudaSetDevice(device));
  SKB_CUDA(cudaMemcpy(host, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
}

// page-locked host memory for the one download a lazy Hessian may need (pageable destinations run at a quarter of the
// PCIe rate); simkit_b200/device_csr.py pools the blocks, so the cost of locking pages is paid once per size
int skb_host_alloc(int64_t nbytes, void** out) {
  if (!out || nbytes <= 0) return fail(SKB_EINVAL, "bad argument");
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, (size_t)nbytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SKB_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  }
  *out = p;
  return SKB_OK;
}

void skb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// The global tiers with the CSR values LEFT ON THE DEVICE in the caller's buffer (skb_buf_alloc, nnz doubles): host
// pointers in for the state and the materials as in skb_gradient_hessian, the gradient (optional) comes back to the host.
int skb_gradient_hessian_resident(skb_plan* pl, int material, int psd_mode, const double* x, const double* Fbar,
                                  const double* mu, int64_t mu_n, const double* lam, int64_t lam_n, const double* vol,
                                  int64_t vol_n, double* g, double* vals_dev) {
  if (!pl || !x || !vals_dev) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  int rc = upload_materials(pl, mu, mu_n, lam, lam_n, vol, vol_n, false, pl->stream);
  if (rc) return rc;
  pl->x.resize(pl->ndof());
  SKB_CUDA(cudaMemcpyAsync(raw(pl->x), x, pl->ndof() * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  if (Fbar) {
    const size_t nf = (size_t)pl->d.t * pl->d.dim * pl->d.dim;
    pl->fbar.resize(nf);
    SKB_CUDA(cudaMemcpyAsync(raw(pl->fbar), Fbar, nf * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  }
  if (g) pl->g.resize(pl->ndof());
  EvalArgs a;
  rc = make_args(pl, material, psd_mode, raw(pl->x), Fbar ? raw(pl->fbar) : nullptr, g ? raw(pl->g) : nullptr, vals_dev, a);
  if (rc) return rc;
  rc = launch_assemble(pl, a, pl->stream);
  if (rc) return rc;
  if (g) SKB_CUDA(cudaMemcpyAsync(g, raw(pl->g), pl->ndof() * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"
