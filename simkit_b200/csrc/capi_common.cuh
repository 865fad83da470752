// Shared pieces of the C-ABI translation units: error plumbing and the plan object.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <exception>
#include <string>
#include <vector>

#include "../../include/simkit_b200.h"
#include "kernels.cuh"

namespace skb {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define SKB_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return skb::fail(SKB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define SKB_TRY try {
#define SKB_CATCH                                                      \
  }                                                                    \
  catch (const std::bad_alloc& ex) {                                   \
    return skb::fail(SKB_ENOMEM, std::string("out of memory: ") + ex.what()); \
  }                                                                    \
  catch (const std::exception& ex) {                                   \
    return skb::fail(SKB_ECUDA, ex.what());                            \
  }

template <class T>
using dvec = thrust::device_vector<T>;

template <class T>
inline T* raw(dvec<T>& v) {
  return thrust::raw_pointer_cast(v.data());
}
template <class T>
inline const T* raw(const dvec<T>& v) {
  return thrust::raw_pointer_cast(v.data());
}

}  // namespace skb

namespace skb {
struct CoarseSpace;                      // coarse.cuh (two-level PCG preconditioner), owned by the plan
void coarse_destroy(CoarseSpace* c);     // capi_solver.cu
}  // namespace skb

// The opaque plan handle of the C ABI.
struct skb_plan {
  int device = 0;
  skb::PlanData<skb::DeviceBackend> d;
  cudaStream_t stream = nullptr;  // stream of the host-pointer entry points
  // device staging of the host-pointer entry points / resident materials
  skb::dvec<double> x, fbar, mu, lam, vol, g, vals;
  skb::dvec<skb::PlanView> pview_dev;  // device copy of view() for kernels that take it by pointer
  skb::dvec<double> stage_elem;  // staging of a per-element host array before it is permuted into the plan's order
  int64_t mu_n = 0, lam_n = 0, vol_n = 0;
  bool have_materials = false;
  // scratch of the reductions
  skb::dvec<double> pblocks, pverts, esums, scalar;
  // PCG / Newton work vectors (allocated on first use)
  skb::dvec<double> w_r, w_z, w_p, w_q, w_dx, w_xt, w_x, w_xtrial, w_dinv, w_diag, w_mass, w_fext, w_xtilde, w_red;
  skb::CoarseSpace* coarse = nullptr;  // skb_pcg_set_coarse
  void* dist = nullptr;                // skb::DistNative (nccl_api.cuh): communicator, halo lists, solver state; skb_nccl_init
  // plane contact springs of the device-resident Newton step (skb_newton_set_contact_plane)
  bool contact_on = false;
  double contact_k = 0.0, contact_p[3] = {0, 0, 0}, contact_n[3] = {0, 0, 0};
  skb::dvec<double> contact_w;
  bool sphere_on = false;      // skb_newton_set_contact_sphere
  double sphere_k = 0.0, sphere_p[3] = {0, 0, 0}, sphere_r = 0.0;
  skb::dvec<double> sphere_w;
  // general sparse quadratic term of the device-resident Newton step (skb_newton_set_quadratic)
  bool quad_on = false;
  skb::dvec<int> quad_ptr, quad_col, quad_pos;
  skb::dvec<double> quad_val, quad_b;
  // resident subspace basis of the reduced tier (skb_plan_set_basis)
  skb::dvec<double> basis;
  int64_t basis_r = 0;
  // work arrays of the reduced tier (F, per-element Hessians He = 81 doubles per tet, weighted stress, energies, x),
  // kept between calls: allocating and freeing 2.7 GB per call at 4 M tets cost ~10 ms of a 120 ms call.  Released
  // with the basis (skb_plan_set_basis(plan, 0, NULL)).
  skb::dvec<double> rw_F, rw_He, rw_Pw, rw_psi, rw_x;
  int launches = 0;
  // optional per-kernel CUDA-event timing (skb_kernel_timing / skb_kernel_times)
  bool timing = false;
  struct TimedLaunch {
    int kind;
    cudaEvent_t a, b;
  };
  std::vector<TimedLaunch> timed;

  skb::PlanView view() const { return d.view(); }
  int64_t ndof() const { return (int64_t)d.n * d.dim; }
  int64_t nnz() const { return (int64_t)d.nnzb * d.dim * d.dim; }
};

// Wraps one kernel launch: counts it and, when timing is on, brackets it with CUDA events on its stream.
#define SKB_LAUNCH(pl, kkind, st, ...)                         \
  do {                                                         \
    skb_plan::TimedLaunch _tl;                                 \
    const bool _t = (pl) && (pl)->timing;                      \
    if (_t) {                                                  \
      _tl.kind = (kkind);                                      \
      cudaEventCreate(&_tl.a);                                 \
      cudaEventCreate(&_tl.b);                                 \
      cudaEventRecord(_tl.a, (st));                            \
    }                                                          \
    __VA_ARGS__;                                               \
    if (_t) {                                                  \
      cudaEventRecord(_tl.b, (st));                            \
      (pl)->timed.push_back(_tl);                              \
    }                                                          \
    if (pl) (pl)->launches++;                                  \
  } while (0)

namespace skb {
// launches shared between translation units
int launch_assemble(skb_plan* pl, const EvalArgs& a, cudaStream_t st);
int launch_energy(skb_plan* pl, const EvalArgs& a, double* out_dev, cudaStream_t st);
int make_args(skb_plan* pl, int material, int psd_mode, const double* x, const double* fbar, double* g,
              double* vals, EvalArgs& a);
int coarse_build(skb_plan* pl, int n_agg, const int* agg_h, const double* xrel_h, int v0, int v1);  // capi_solver.cu
int upload_materials(skb_plan* pl, const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                     const double* vol, int64_t vol_n, bool from_device, cudaStream_t st);
}  // namespace skb
