// C ABI, part 9: dense tall-skinny building blocks of the subspace construction:
// the QR behind `orthonormalize` (orthonormalize.py:9-49) and the Gram products behind `project_into_subspace`
// (project_into_subspace.py:9-59).  One-off set-up operations on an (n*dim) x r basis: plain library calls
// (cuSOLVER geqrf + orgqr, cuBLAS gemm / geam), the LAPACK algorithms numpy itself runs, so the factors carry LAPACK's
// sign convention and match the reference's output column for column.
#include "capi_common.cuh"

#include <cublas_v2.h>
#include <cusolverDn.h>

using namespace skb;

namespace {
#define SKB_CUBLAS(call)                                                                   \
  do {                                                                                     \
    cublasStatus_t _s = (call);                                                            \
    if (_s != CUBLAS_STATUS_SUCCESS) return fail(SKB_ECUDA, "cuBLAS call failed: " #call); \
  } while (0)
#define SKB_CUSOLVER2(call)                                                                      \
  do {                                                                                           \
    cusolverStatus_t _s = (call);                                                                \
    if (_s != CUSOLVER_STATUS_SUCCESS) return fail(SKB_ECUDA, "cuSOLVER call failed: " #call);   \
  } while (0)

struct Handles {
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
  ~Handles() {
    if (solver) cusolverDnDestroy(solver);
    if (blas) cublasDestroy(blas);
  }
};
}  // namespace

extern "C" {

// Thin Householder QR of a row-major (n x r) matrix, n >= r: Q (n x r, row-major), R (r x r, row-major, upper
// triangular) exactly as numpy.linalg.qr(A) returns them (LAPACK dgeqrf + dorgqr).
int skb_qr_thin(int64_t n, int64_t r, const double* A, double* Q, double* R) {
  if (!A || !Q || !R) return fail(SKB_EINVAL, "null argument");
  if (r <= 0 || n < r || n >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "need n >= r >= 1");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  Handles h;
  SKB_CUBLAS(cublasCreate(&h.blas));
  SKB_CUSOLVER2(cusolverDnCreate(&h.solver));
  const int ni = (int)n, ri = (int)r;
  dvec<double> Arm(A, A + n * r), Acm((size_t)n * r), tau(r);
  const double one = 1.0, zero = 0.0;
  // row-major (n x r) is column-major (r x n): transpose into column-major (n x r)
  SKB_CUBLAS(cublasDgeam(h.blas, CUBLAS_OP_T, CUBLAS_OP_N, ni, ri, &one, raw(Arm), ri, &zero, raw(Acm), ni, raw(Acm), ni));
  int lw1 = 0, lw2 = 0;
  SKB_CUSOLVER2(cusolverDnDgeqrf_bufferSize(h.solver, ni, ri, raw(Acm), ni, &lw1));
  SKB_CUSOLVER2(cusolverDnDorgqr_bufferSize(h.solver, ni, ri, ri, raw(Acm), ni, raw(tau), &lw2));
  dvec<double> work(lw1 > lw2 ? lw1 : lw2);
  dvec<int> info(1, 0);
  SKB_CUSOLVER2(cusolverDnDgeqrf(h.solver, ni, ri, raw(Acm), ni, raw(tau), raw(work), (int)work.size(), raw(info)));
  // R = upper triangle of the factored matrix (column-major n x r) -> row-major r x r on the host
  thrust::host_vector<double> Ah = Acm;
  for (int i = 0; i < ri; ++i)
    for (int j = 0; j < ri; ++j) R[(size_t)i * ri + j] = (j >= i) ? Ah[(size_t)j * ni + i] : 0.0;
  SKB_CUSOLVER2(cusolverDnDorgqr(h.solver, ni, ri, ri, raw(Acm), ni, raw(tau), raw(work), (int)work.size(), raw(info)));
  if ((int)info[0] != 0) return fail(SKB_ECUDA, "QR factorisation failed");
  // back to row-major
  SKB_CUBLAS(cublasDgeam(h.blas, CUBLAS_OP_T, CUBLAS_OP_N, ri, ni, &one, raw(Acm), ni, &zero, raw(Arm), ri, raw(Arm), ri));
  SKB_CUDA(cudaMemcpy(Q, raw(Arm), (size_t)n * r * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

// G = A^T diag(w) B  (r x s, row-major) for row-major A (n x r), B (n x s) and an optional weight vector w (n):
// B^T M B and B^T M y of project_into_subspace.py:49-53 for a diagonal M.
int skb_weighted_gram(int64_t n, int64_t r, int64_t s, const double* A, const double* w, const double* B, double* G) {
  if (!A || !B || !G) return fail(SKB_EINVAL, "null argument");
  if (n <= 0 || r <= 0 || s <= 0 || n >= ((int64_t)1 << 31)) return fail(SKB_EINVAL, "bad size");
  if (skb_device_count() <= 0) return fail(SKB_ENOGPU, "no CUDA device");
  SKB_TRY
  Handles h;
  SKB_CUBLAS(cublasCreate(&h.blas));
  const int ni = (int)n, ri = (int)r, si = (int)s;
  dvec<double> Ad(A, A + n * r), Bd(B, B + n * s), Gd((size_t)r * s);
  if (w) {
    // rows of B scaled by w: B is column-major (s x n), so this is a right multiplication by diag(w)
    dvec<double> wd(w, w + n), Bs((size_t)n * s);
    SKB_CUBLAS(cublasDdgmm(h.blas, CUBLAS_SIDE_RIGHT, si, ni, raw(Bd), si, raw(wd), 1, raw(Bs), si));
    Bd.swap(Bs);
  }
  const double one = 1.0, zero = 0.0;
  // row-major G (r x s) = column-major (s x r) = Bcm (s x n) * Acm^T (n x r)
  SKB_CUBLAS(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_T, si, ri, ni, &one, raw(Bd), si, raw(Ad), ri, &zero, raw(Gd), si));
  SKB_CUDA(cudaMemcpy(G, raw(Gd), (size_t)r * s * sizeof(double), cudaMemcpyDeviceToHost));
  return SKB_OK;
  SKB_CATCH
}

}  // extern "C"
