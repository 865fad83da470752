// C ABI, part 1: plan lifetime, pattern export and the global energy / gradient /
// Hessian entry points (include/simkit_b200.h).
#include "capi_common.cuh"

#include <stdlib.h>

#include <algorithm>
#include <memory>

namespace skb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int make_args(skb_plan* pl, int material, int psd_mode, const double* x, const double* fbar, double* g,
              double* vals, EvalArgs& a) {
  if (material < 0 || material >= MAT_COUNT) return fail(SKB_EINVAL, "unknown material id");
  if (psd_mode < 0 || psd_mode > PSD_ABS_AFTER_VOL) return fail(SKB_EINVAL, "unknown psd mode");
  if (!pl->have_materials) return fail(SKB_EINVAL, "materials not set (skb_set_materials)");
  a.material = material;
  a.psd_mode = psd_mode;
  a.x = x;
  a.Fbar = fbar;
  a.eorder = pl->d.eorder.empty() ? nullptr : raw(pl->d.eorder);
  a.mu = raw(pl->mu);
  a.lam = raw(pl->lam);
  a.vol = raw(pl->vol);
  a.mu_stride = pl->mu_n > 1 ? 1 : 0;
  a.lam_stride = pl->lam_n > 1 ? 1 : 0;
  a.vol_stride = pl->vol_n > 1 ? 1 : 0;
  a.want_grad = g != nullptr;
  a.want_hess = vals != nullptr;
  a.g = g;
  a.vals = vals;
  const size_t rec_stride = pl->d.dim == 3 ? RecStride<3>::value : RecStride<2>::value;
  if (a.want_hess && pl->pblocks.size() < (size_t)pl->d.blocks.n_ts * rec_stride)
    pl->pblocks.resize((size_t)pl->d.blocks.n_ts * rec_stride);
  if (a.want_grad && pl->pverts.size() < (size_t)pl->d.verts.n_ts * pl->d.dim)
    pl->pverts.resize((size_t)pl->d.verts.n_ts * pl->d.dim);
  a.pblocks = raw(pl->pblocks);
  a.pverts = raw(pl->pverts);
  return SKB_OK;
}

// dst[i] = src[order[i]]: per-element inputs arrive in the caller's element order, the plan keeps them in its own
__global__ void permute_in_kernel(const double* src, const int* order, int64_t n, double* dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[order[i]];
}
// dst[order[i]] = src[i]: the way back, for per-element exports
__global__ void permute_out_kernel(const double* src, const int* order, int64_t n, double* dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[order[i]] = src[i];
}

// one per-element (or scalar) input: copied as is when scalar or when the plan keeps the caller's element order,
// otherwise staged and gathered into the plan's internal element order
static int upload_per_element(skb_plan* pl, dvec<double>& dst, const double* src, int64_t cnt, bool from_device,
                              cudaStream_t st) {
  const cudaMemcpyKind kind = from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  dst.resize(cnt);
  if (cnt <= 1 || pl->d.eorder.empty()) {
    SKB_CUDA(cudaMemcpyAsync(raw(dst), src, cnt * sizeof(double), kind, st));
    return SKB_OK;
  }
  const double* dsrc = src;
  if (!from_device) {
    pl->stage_elem.resize(cnt);
    SKB_CUDA(cudaMemcpyAsync(raw(pl->stage_elem), src, cnt * sizeof(double), kind, st));
    dsrc = raw(pl->stage_elem);
  }
  permute_in_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(dsrc, raw(pl->d.eorder), cnt, raw(dst));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int upload_materials(skb_plan* pl, const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                     const double* vol, int64_t vol_n, bool from_device, cudaStream_t st) {
  const int64_t t = pl->d.t;
  if (!mu || (mu_n != 1 && mu_n != t)) return fail(SKB_EINVAL, "mu must have 1 or t entries");
  if (lam && lam_n != 1 && lam_n != t) return fail(SKB_EINVAL, "lam must have 1 or t entries");
  if (vol && vol_n != 1 && vol_n != t) return fail(SKB_EINVAL, "vol must have 1 or t entries");
  int rc = upload_per_element(pl, pl->mu, mu, mu_n, from_device, st);
  if (rc) return rc;
  pl->mu_n = mu_n;
  if (lam) {
    rc = upload_per_element(pl, pl->lam, lam, lam_n, from_device, st);
    if (rc) return rc;
    pl->lam_n = lam_n;
  } else {
    pl->lam.assign(1, 0.0);
    pl->lam_n = 1;
  }
  if (vol) {
    rc = upload_per_element(pl, pl->vol, vol, vol_n, from_device, st);
    if (rc) return rc;
    pl->vol_n = vol_n;
  } else {
    if (!pl->d.has_vol0) return fail(SKB_EINVAL, "vol is required for a plan built from an operator");
    pl->vol = pl->d.vol0;  // already in the internal order
    pl->vol_n = t;
  }
  pl->have_materials = true;
  return SKB_OK;
}

// shape of the pipelined assembly CTA: G groups of E threads (one tile of E elements each) sharing NBUF
// staging buffers.  Overridable at build time for A/B experiments (scripts/ab.sh).
#ifndef SKB_PIPE_E
#define SKB_PIPE_E 128
#endif
#ifndef SKB_PIPE_G
#define SKB_PIPE_G 3
#endif
#ifndef SKB_PIPE_NBUF
#define SKB_PIPE_NBUF 2
#endif
// default assembly kernel when SKB_ASSEMBLE is not set: 1 = pipelined, 2 = warp-specialised (falls back to the
// pipelined kernel, then to one CTA per tile, when the tile size or the schedule does not fit)
#ifndef SKB_ASSEMBLE_DEFAULT
#define SKB_ASSEMBLE_DEFAULT 2
#endif

template <int D, int G, int NBUF, int MAT>
static int launch_pipelined(skb_plan* pl, const PlanView& p, const EvalArgs& a, size_t psmem, int grid, cudaStream_t st) {
  constexpr int E = SKB_PIPE_E;
  static bool attr_set = false;
  if (!attr_set) {
    SKB_CUDA(cudaFuncSetAttribute(assemble_pipelined_kernel<D, G, NBUF, MAT, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    attr_set = true;
  }
  SKB_LAUNCH(pl, SKB_K_ASSEMBLE, st, assemble_pipelined_kernel<D, G, NBUF, MAT, E><<<grid, G * E, psmem, st>>>(p, a));
  return SKB_OK;
}

// warp-specialised persistent kernel (kernels.cuh assemble_ws_kernel): 2 compute + 2 reducer warpgroups per CTA
template <int D, int MAT>
static int launch_ws(skb_plan* pl, const PlanView& p, const EvalArgs& a, size_t wsmem, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SKB_CUDA(cudaFuncSetAttribute(assemble_ws_kernel<D, MAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  SKB_LAUNCH(pl, SKB_K_ASSEMBLE, st, assemble_ws_kernel<D, MAT><<<grid, 512, wsmem, st>>>(p, a));
  return SKB_OK;
}

template <int D>
static int launch_assemble_t(skb_plan* pl, const EvalArgs& a, cudaStream_t st) {
  const PlanView p = pl->view();
  const int E = p.tile_elems;
  const size_t smem = assemble_smem_bytes<D>(p);
  if (smem > 227 * 1024) return fail(SKB_EINVAL, "tile_elems too large for the 227 KB of shared memory");
  // pipelined persistent kernel (default 3 groups, 2 staging buffers) when the tile has SKB_PIPE_E elements and its
  // schedule fits
  constexpr int G = SKB_PIPE_G, NBUF = SKB_PIPE_NBUF;
  const size_t psmem = PipeSmem<D>::total(p, G, NBUF);
  static int use_pipe_env = -1;
  if (use_pipe_env < 0) {
    const char* ev = getenv("SKB_ASSEMBLE");
    use_pipe_env = (ev && strcmp(ev, "tile") == 0) ? 0 : (ev && strcmp(ev, "ws") == 0) ? 2 : (ev && strcmp(ev, "pipe") == 0) ? 1 : SKB_ASSEMBLE_DEFAULT;
  }
  const size_t wsmem = WsSmem<D>::total(p);
  const bool ws = use_pipe_env == 2 && E == 128 && wsmem <= 227 * 1024 && p.n_tiles >= 2 * WsSmem<D>::P;
  const bool pipe = !ws && use_pipe_env && E == SKB_PIPE_E && psmem <= 227 * 1024 && p.n_tiles >= 2 * G;
  static bool told = false;
  if (!told && getenv("SKB_VERBOSE")) {
    told = true;
    fprintf(stderr, "simkit_b200: assembly %s, tile %d elements, %d groups, %d buffers, %zu B shared (pipelined) / %zu B (tile) / %zu B (warp-specialised)\n",
            ws ? "warp-specialised" : pipe ? "pipelined" : "one CTA per tile", E, G, NBUF, psmem, smem, wsmem);
  }
  if (ws) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    constexpr int P = WsSmem<D>::P;
    int grid = sms;
    if (grid * P > p.n_tiles) grid = (p.n_tiles + P - 1) / P;
    int rc = SKB_OK;
    if (D == 3) {
      switch (a.material) {
        case MAT_STABLE_NEO_HOOKEAN: rc = launch_ws<3, MAT_STABLE_NEO_HOOKEAN>(pl, p, a, wsmem, grid, st); break;
        case MAT_NEO_HOOKEAN: rc = launch_ws<3, MAT_NEO_HOOKEAN>(pl, p, a, wsmem, grid, st); break;
        case MAT_ARAP: rc = launch_ws<3, MAT_ARAP>(pl, p, a, wsmem, grid, st); break;
        case MAT_STVK: rc = launch_ws<3, MAT_STVK>(pl, p, a, wsmem, grid, st); break;
        case MAT_FCR: rc = launch_ws<3, MAT_FCR>(pl, p, a, wsmem, grid, st); break;
        case MAT_MACKLIN_MUELLER_NEO_HOOKEAN: rc = launch_ws<3, MAT_MACKLIN_MUELLER_NEO_HOOKEAN>(pl, p, a, wsmem, grid, st); break;
        default: rc = launch_ws<3, MAT_LINEAR_ELASTICITY>(pl, p, a, wsmem, grid, st); break;
      }
    } else {
      rc = launch_ws<2, -1>(pl, p, a, wsmem, grid, st);
    }
    if (rc) return rc;
  } else if (pipe) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    int grid = sms;
    if (grid * G > p.n_tiles) grid = (p.n_tiles + G - 1) / G;
    int rc = SKB_OK;
    if (D == 3) {
      // tets: one kernel per material (compile-time constitutive model)
      switch (a.material) {
        case MAT_STABLE_NEO_HOOKEAN: rc = launch_pipelined<3, G, NBUF, MAT_STABLE_NEO_HOOKEAN>(pl, p, a, psmem, grid, st); break;
        case MAT_NEO_HOOKEAN: rc = launch_pipelined<3, G, NBUF, MAT_NEO_HOOKEAN>(pl, p, a, psmem, grid, st); break;
        case MAT_ARAP: rc = launch_pipelined<3, G, NBUF, MAT_ARAP>(pl, p, a, psmem, grid, st); break;
        case MAT_STVK: rc = launch_pipelined<3, G, NBUF, MAT_STVK>(pl, p, a, psmem, grid, st); break;
        case MAT_FCR: rc = launch_pipelined<3, G, NBUF, MAT_FCR>(pl, p, a, psmem, grid, st); break;
        case MAT_MACKLIN_MUELLER_NEO_HOOKEAN:
          rc = launch_pipelined<3, G, NBUF, MAT_MACKLIN_MUELLER_NEO_HOOKEAN>(pl, p, a, psmem, grid, st);
          break;
        default: rc = launch_pipelined<3, G, NBUF, MAT_LINEAR_ELASTICITY>(pl, p, a, psmem, grid, st); break;
      }
    } else {
      rc = launch_pipelined<2, G, NBUF, -1>(pl, p, a, psmem, grid, st);
    }
    if (rc) return rc;
  } else {
    static bool attr_set[2] = {false, false};
    if (!attr_set[D - 2]) {
      SKB_CUDA(cudaFuncSetAttribute(assemble_tile_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[D - 2] = true;
    }
    SKB_LAUNCH(pl, SKB_K_ASSEMBLE, st, assemble_tile_kernel<D><<<p.n_tiles, E, smem, st>>>(p, a));
  }
  if (a.want_hess) {
    const int fin_grid = (p.nu * D * D + SKB_FIN_THREADS * SKB_FIN_PER_THREAD - 1) / (SKB_FIN_THREADS * SKB_FIN_PER_THREAD);
    SKB_LAUNCH(pl, SKB_K_FINALIZE_BLOCKS, st, finalize_blocks_kernel<D><<<fin_grid, SKB_FIN_THREADS, 0, st>>>(p, a.pblocks, a.vals));
  }
  if (a.want_grad) {
    SKB_LAUNCH(pl, SKB_K_FINALIZE_VERTS, st,
               finalize_verts_kernel<D><<<(p.n + 255) / 256, 256, 0, st>>>(p, a.pverts, a.g));
  }
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int launch_assemble(skb_plan* pl, const EvalArgs& a, cudaStream_t st) {
  return pl->d.dim == 3 ? launch_assemble_t<3>(pl, a, st) : launch_assemble_t<2>(pl, a, st);
}

int launch_energy(skb_plan* pl, const EvalArgs& a, double* out_dev, cudaStream_t st) {
  const PlanView p = pl->view();
  const int count = pl->d.t_energy;   // a shard that re-evaluates its neighbour's interface elements counts only its own
  const int nb = (count + 255) / 256;
  if (pl->esums.size() < (size_t)nb) pl->esums.resize(nb);
  if (p.dim == 3)
    SKB_LAUNCH(pl, SKB_K_ENERGY, st, energy_kernel<3><<<nb, 256, 0, st>>>(p, a, count, raw(pl->esums)));
  else
    SKB_LAUNCH(pl, SKB_K_ENERGY, st, energy_kernel<2><<<nb, 256, 0, st>>>(p, a, count, raw(pl->esums)));
  SKB_LAUNCH(pl, SKB_K_OTHER, st, reduce_final_kernel<<<1, 1024, 0, st>>>(raw(pl->esums), nb, out_dev));
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

// dst[i] = src[idx[i]]  /  dst[idx[i]] += src[i]  (idx entries are distinct within one call, so the
// scatter is race-free and the result does not depend on thread order)
__global__ void gather_kernel(const double* src, const int32_t* idx, int64_t n, double* dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(double* dst, const int32_t* idx, int64_t n, const double* src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}
__global__ void scatter_add_kernel(double* dst, const int32_t* idx, int64_t n, const double* src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] += src[i];
}

}  // namespace skb

using namespace skb;

extern "C" {

const char* skb_last_error(void) { return g_err.c_str(); }

const char* skb_version(void) { return "simkit_b200 0.1.0 sm_100a"; }

int skb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int plan_create_common(const double* X, const double* Dop, const void* T, int index_bytes, int64_t n,
                              int64_t t, int dim, int device, int tile_elems, skb_plan** out, int64_t t_total = -1,
                              int64_t t_energy = 0) {
  if (t_total < t) t_total = t;
  if (t_energy <= 0 || t_energy > t) t_energy = t;
  if ((!X && !Dop) || !T || !out) return fail(SKB_EINVAL, "null argument");
  if (dim != 2 && dim != 3) return fail(SKB_EINVAL, "Only dim == 2 or 3 are supported");
  if (index_bytes != 4 && index_bytes != 8) return fail(SKB_EINVAL, "index_bytes must be 4 or 8");
  if (n <= 0 || t <= 0) return fail(SKB_EINVAL, "empty mesh");
  const int K = dim + 1;
  if (t_total * K * K >= (int64_t)1 << 31 || n * dim >= (int64_t)1 << 31) return fail(SKB_EINVAL, "mesh too large for int32 indexing");
  if (tile_elems == 0) tile_elems = SKB_PIPE_E;
  if (tile_elems < 32 || tile_elems > 256 || tile_elems % 32) return fail(SKB_EINVAL, "tile_elems must be a multiple of 32 in [32, 256]");
  if (skb_device_count() <= device) return fail(SKB_ENOGPU, "no CUDA device " + std::to_string(device));
  SKB_CUDA(cudaSetDevice(device));
  SKB_TRY
  // validate + narrow indices on the host (one pass; the plan build itself runs on the device)
  thrust::host_vector<int> Th((size_t)t_total * K);
  for (int64_t i = 0; i < t_total * K; ++i) {
    int64_t v = index_bytes == 8 ? ((const int64_t*)T)[i] : (int64_t)((const int32_t*)T)[i];
    if (v < 0 || v >= n) return fail(SKB_EINVAL, "element index out of range");
    Th[i] = (int)v;
  }
  std::unique_ptr<skb_plan> pl(new skb_plan());
  pl->device = device;
  SKB_CUDA(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
  dvec<int> Td = Th;
  dvec<double> Xd;
  if (X) {
    Xd.assign(X, X + n * dim);
    // spatial element order of the plan's own (plan.cuh str_element_order); SKB_ELEMENT_ORDER=input keeps the caller's
    const char* eo = getenv("SKB_ELEMENT_ORDER");
    if (!(eo && strcmp(eo, "input") == 0) && t > tile_elems)
      apply_element_order<DeviceBackend>(pl->d, Td, Xd, (int)n, (int)t, dim, tile_elems, (int)t_energy);
  }
  pl->d.t_energy = (int)t_energy;
  if (!build_plan<DeviceBackend>(pl->d, Td, (int)n, (int)t, dim, tile_elems, (int)t_total))
    return fail(SKB_EINVAL, "degenerate element: a vertex is repeated within one element");
  if (X) {
    set_geometry_from_X<DeviceBackend>(pl->d, Xd);
  } else {
    dvec<double> Dd(Dop, Dop + t * dim * K);
    set_geometry_from_D<DeviceBackend>(pl->d, Dd);
  }
  SKB_CUDA(cudaDeviceSynchronize());
  pl->scalar.resize(8);
  *out = pl.release();
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_create(const double* X, const void* T, int index_bytes, int64_t n, int64_t t, int dim,
                    int device, int tile_elems, skb_plan** out) {
  if (!X) return fail(SKB_EINVAL, "null X");
  return plan_create_common(X, nullptr, T, index_bytes, n, t, dim, device, tile_elems, out);
}

int skb_plan_create_sharded(const double* X, const void* T, int index_bytes, int64_t n, int64_t t_active,
                            int64_t t_total, int64_t t_energy, int dim, int device, int tile_elems, skb_plan** out) {
  if (!X) return fail(SKB_EINVAL, "null X");
  if (t_total < t_active) return fail(SKB_EINVAL, "t_total must be >= t_active");
  if (t_energy < 0 || t_energy > t_active) return fail(SKB_EINVAL, "t_energy must be in [0, t_active]");
  return plan_create_common(X, nullptr, T, index_bytes, n, t_active, dim, device, tile_elems, out, t_total, t_energy);
}

int skb_plan_create_from_operator(const void* T, const double* D, int index_bytes, int64_t n, int64_t t,
                                  int dim, int device, int tile_elems, skb_plan** out) {
  if (!D) return fail(SKB_EINVAL, "null D");
  return plan_create_common(nullptr, D, T, index_bytes, n, t, dim, device, tile_elems, out);
}

void skb_plan_destroy(skb_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->device);
  skb_nccl_finalize(plan);
  if (plan->stream) cudaStreamDestroy(plan->stream);
  if (plan->coarse) skb::coarse_destroy(plan->coarse);
  delete plan;
}

int skb_plan_info(const skb_plan* pl, int64_t info[8]) {
  if (!pl || !info) return fail(SKB_EINVAL, "null argument");
  info[0] = pl->d.n;
  info[1] = pl->d.t;
  info[2] = pl->d.dim;
  info[3] = pl->d.nnzb;
  info[4] = pl->nnz();
  info[5] = pl->d.n_tiles;
  info[6] = pl->d.blocks.n_ts;
  info[7] = pl->d.verts.n_ts;
  return SKB_OK;
}

int skb_plan_block_pattern(const skb_plan* pl, int32_t* bptr, int32_t* bcol) {
  if (!pl || !bptr || !bcol) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_CUDA(cudaMemcpy(bptr, raw(pl->d.bptr), (pl->d.n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  SKB_CUDA(cudaMemcpy(bcol, raw(pl->d.bcol), (size_t)pl->d.nnzb * sizeof(int), cudaMemcpyDeviceToHost));
  return SKB_OK;
}

int skb_plan_csr_pattern(const skb_plan* pl, int32_t* indptr, int32_t* indices) {
  if (!pl || !indptr || !indices) return fail(SKB_EINVAL, "null argument");
  SKB_TRY
  const int n = pl->d.n, D = pl->d.dim;
  thrust::host_vector<int> bptr = pl->d.bptr, bcol = pl->d.bcol;
  int64_t pos = 0;
  for (int v = 0; v < n; ++v) {
    const int b0 = bptr[v], nb = bptr[v + 1] - b0;
    for (int i = 0; i < D; ++i) {
      indptr[(int64_t)v * D + i] = (int32_t)pos;
      for (int j = 0; j < nb; ++j)
        for (int k = 0; k < D; ++k) indices[pos++] = bcol[b0 + j] * D + k;
    }
  }
  indptr[(int64_t)n * D] = (int32_t)pos;
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_slot_map(const skb_plan* pl, int32_t* slot) {
  if (!pl || !slot) return fail(SKB_EINVAL, "null argument");
  SKB_TRY
  const int D = pl->d.dim, K = pl->d.K, t = pl->d.t;
  thrust::host_vector<int> bptr = pl->d.bptr, bcol = pl->d.bcol, Ts = pl->d.T32;
  thrust::host_vector<uint8_t> perm = pl->d.perm;
  thrust::host_vector<int> eo = pl->d.eorder;
  for (int ei = 0; ei < t; ++ei) {
    const int e = eo.empty() ? ei : eo[ei];  // caller's element index
    int To[4];
    for (int s = 0; s < K; ++s) To[perm[ei * K + s]] = Ts[ei * K + s];  // caller's corner order
    for (int a = 0; a < K; ++a) {
      const int v = To[a];
      const int b0 = bptr[v], nb = bptr[v + 1] - b0;
      for (int b = 0; b < K; ++b) {
        const int* lo = &bcol[b0];
        const int s = b0 + (int)(std::lower_bound(lo, lo + nb, To[b]) - lo);
        for (int i = 0; i < D; ++i)
          for (int k = 0; k < D; ++k)
            slot[((((int64_t)e * K + a) * D + i) * K + b) * D + k] =
                (int32_t)((int64_t)b0 * D * D + (int64_t)i * nb * D + (int64_t)(s - b0) * D + k);
      }
    }
  }
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_element_D(const skb_plan* pl, double* Dout) {
  if (!pl || !Dout) return fail(SKB_EINVAL, "null argument");
  SKB_TRY
  const int D = pl->d.dim, K = pl->d.K, t = pl->d.t;
  thrust::host_vector<double> Dm = pl->d.Dm;
  thrust::host_vector<uint8_t> perm = pl->d.perm;
  thrust::host_vector<int> eo = pl->d.eorder;
  for (int ei = 0; ei < t; ++ei) {
    const int e = eo.empty() ? ei : eo[ei];  // caller's element index
    for (int j = 0; j < D; ++j) {
      double s0 = 0.0;
      for (int s = 1; s < K; ++s) {
        double v = Dm[(size_t)(j * D + (s - 1)) * t + ei];
        Dout[((size_t)e * D + j) * K + perm[ei * K + s]] = v;
        s0 -= v;
      }
      Dout[((size_t)e * D + j) * K + perm[ei * K]] = s0;
    }
  }
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_volume(const skb_plan* pl, double* vol) {
  if (!pl || !vol) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  thrust::host_vector<double> v0 = pl->d.vol0;
  thrust::host_vector<int> eo = pl->d.eorder;
  for (int i = 0; i < pl->d.t; ++i) vol[eo.empty() ? i : eo[i]] = v0[i];
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_element_order(const skb_plan* pl, int32_t* order) {
  if (!pl || !order) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  thrust::host_vector<int> eo = pl->d.eorder;
  for (int i = 0; i < pl->d.t; ++i) order[i] = eo.empty() ? i : eo[i];
  return SKB_OK;
  SKB_CATCH
}

int skb_plan_vertex_masses(const skb_plan* pl, const double* rho, int64_t rho_n, double* m) {
  if (!pl || !rho || !m) return fail(SKB_EINVAL, "null argument");
  if (rho_n != 1 && rho_n != pl->d.t) return fail(SKB_EINVAL, "rho must have 1 or t entries");
  SKB_TRY
  // one-off setup quantity (massmatrix.py:41-49); summed in element order on the host
  const int K = pl->d.K, t = pl->d.t, n = pl->d.n;
  thrust::host_vector<double> vol = pl->d.vol0;
  thrust::host_vector<int> T = pl->d.T32;
  thrust::host_vector<int> eo = pl->d.eorder;
  for (int v = 0; v < n; ++v) m[v] = 0.0;
  for (int e = 0; e < t; ++e) {
    const double me = vol[e] * rho[rho_n > 1 ? (eo.empty() ? e : eo[e]) : 0];
    for (int a = 0; a < K; ++a) m[T[e * K + a]] += me;
  }
  for (int v = 0; v < n; ++v) m[v] /= K;
  return SKB_OK;
  SKB_CATCH
}

int skb_set_materials(skb_plan* pl, const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                      const double* vol, int64_t vol_n) {
  if (!pl) return fail(SKB_EINVAL, "null plan");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  int rc = upload_materials(pl, mu, mu_n, lam, lam_n, vol, vol_n, false, pl->stream);
  if (rc) return rc;
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

int skb_set_materials_dev(skb_plan* pl, const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                          const double* vol, int64_t vol_n, void* stream) {
  if (!pl) return fail(SKB_EINVAL, "null plan");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  return upload_materials(pl, mu, mu_n, lam, lam_n, vol, vol_n, true, (cudaStream_t)stream);
  SKB_CATCH
}

int skb_last_launch_count(const skb_plan* pl) { return pl ? pl->launches : 0; }

int skb_kernel_timing(skb_plan* pl, int enable) {
  if (!pl) return fail(SKB_EINVAL, "null plan");
  pl->timing = enable != 0;
  return SKB_OK;
}

int skb_kernel_times(skb_plan* pl, double* ms, int64_t* launches) {
  if (!pl || !ms || !launches) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  for (int k = 0; k < SKB_K_COUNT; ++k) {
    ms[k] = 0.0;
    launches[k] = 0;
  }
  for (auto& tl : pl->timed) {
    SKB_CUDA(cudaEventSynchronize(tl.b));
    float e = 0.f;
    SKB_CUDA(cudaEventElapsedTime(&e, tl.a, tl.b));
    const int k = (tl.kind >= 0 && tl.kind < SKB_K_COUNT) ? tl.kind : SKB_K_OTHER;
    ms[k] += e;
    launches[k]++;
    cudaEventDestroy(tl.a);
    cudaEventDestroy(tl.b);
  }
  pl->timed.clear();
  return SKB_OK;
}

// ---- host-pointer global tiers -------------------------------------------
static int stage_inputs(skb_plan* pl, const double* x, const double* Fbar, const double* mu, int64_t mu_n,
                        const double* lam, int64_t lam_n, const double* vol, int64_t vol_n) {
  if (!x) return fail(SKB_EINVAL, "null x");
  int rc = upload_materials(pl, mu, mu_n, lam, lam_n, vol, vol_n, false, pl->stream);
  if (rc) return rc;
  pl->x.resize(pl->ndof());
  SKB_CUDA(cudaMemcpyAsync(raw(pl->x), x, pl->ndof() * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  if (Fbar) {
    const size_t nf = (size_t)pl->d.t * pl->d.dim * pl->d.dim;
    pl->fbar.resize(nf);
    SKB_CUDA(cudaMemcpyAsync(raw(pl->fbar), Fbar, nf * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
  }
  return SKB_OK;
}

int skb_energy(skb_plan* pl, int material, const double* x, const double* Fbar, const double* mu,
               int64_t mu_n, const double* lam, int64_t lam_n, const double* vol, int64_t vol_n,
               double* energy) {
  if (!pl || !energy) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  int rc = stage_inputs(pl, x, Fbar, mu, mu_n, lam, lam_n, vol, vol_n);
  if (rc) return rc;
  EvalArgs a;
  rc = make_args(pl, material, PSD_NONE, raw(pl->x), Fbar ? raw(pl->fbar) : nullptr, nullptr, nullptr, a);
  if (rc) return rc;
  rc = launch_energy(pl, a, raw(pl->scalar), pl->stream);
  if (rc) return rc;
  SKB_CUDA(cudaMemcpyAsync(energy, raw(pl->scalar), sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

int skb_gradient_hessian(skb_plan* pl, int material, int psd_mode, const double* x, const double* Fbar,
                         const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                         const double* vol, int64_t vol_n, double* g, double* vals) {
  if (!pl) return fail(SKB_EINVAL, "null plan");
  if (!g && !vals) return fail(SKB_EINVAL, "nothing to compute");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  pl->launches = 0;
  int rc = stage_inputs(pl, x, Fbar, mu, mu_n, lam, lam_n, vol, vol_n);
  if (rc) return rc;
  if (g) pl->g.resize(pl->ndof());
  if (vals) pl->vals.resize(pl->nnz());
  EvalArgs a;
  rc = make_args(pl, material, psd_mode, raw(pl->x), Fbar ? raw(pl->fbar) : nullptr, g ? raw(pl->g) : nullptr,
                 vals ? raw(pl->vals) : nullptr, a);
  if (rc) return rc;
  rc = launch_assemble(pl, a, pl->stream);
  if (rc) return rc;
  if (g) SKB_CUDA(cudaMemcpyAsync(g, raw(pl->g), pl->ndof() * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  if (vals) SKB_CUDA(cudaMemcpyAsync(vals, raw(pl->vals), pl->nnz() * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  SKB_CUDA(cudaStreamSynchronize(pl->stream));
  return SKB_OK;
  SKB_CATCH
}

int skb_gradient(skb_plan* pl, int material, const double* x, const double* Fbar, const double* mu,
                 int64_t mu_n, const double* lam, int64_t lam_n, const double* vol, int64_t vol_n, double* g) {
  if (!g) return fail(SKB_EINVAL, "null g");
  return skb_gradient_hessian(pl, material, SKB_PSD_NONE, x, Fbar, mu, mu_n, lam, lam_n, vol, vol_n, g, nullptr);
}

int skb_hessian(skb_plan* pl, int material, int psd_mode, const double* x, const double* Fbar,
                const double* mu, int64_t mu_n, const double* lam, int64_t lam_n, const double* vol,
                int64_t vol_n, double* vals) {
  if (!vals) return fail(SKB_EINVAL, "null vals");
  return skb_gradient_hessian(pl, material, psd_mode, x, Fbar, mu, mu_n, lam, lam_n, vol, vol_n, nullptr, vals);
}

// ---- device-pointer variants ---------------------------------------------
int skb_energy_dev(skb_plan* pl, int material, const double* x, const double* Fbar, double* energy_out,
                   void* stream) {
  if (!pl || !x || !energy_out) return fail(SKB_EINVAL, "null argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  EvalArgs a;
  int rc = make_args(pl, material, PSD_NONE, x, Fbar, nullptr, nullptr, a);
  if (rc) return rc;
  return launch_energy(pl, a, energy_out, (cudaStream_t)stream);
  SKB_CATCH
}

int skb_gradient_hessian_dev(skb_plan* pl, int material, int psd_mode, const double* x, const double* Fbar,
                             double* g, double* vals, void* stream) {
  if (!pl || !x) return fail(SKB_EINVAL, "null argument");
  if (!g && !vals) return fail(SKB_EINVAL, "nothing to compute");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  EvalArgs a;
  int rc = make_args(pl, material, psd_mode, x, Fbar, g, vals, a);
  if (rc) return rc;
  return launch_assemble(pl, a, (cudaStream_t)stream);
  SKB_CATCH
}


// ---- interface exchange helpers (device pointers) --------------------------
int skb_gather_dev(const double* src, const int32_t* idx, int64_t n, double* dst, void* stream) {
  if (n == 0) return SKB_OK;
  if (!src || !idx || !dst || n < 0) return fail(SKB_EINVAL, "bad argument");
  gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, dst);
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_scatter_dev(double* dst, const int32_t* idx, int64_t n, const double* src, void* stream) {
  if (n == 0) return SKB_OK;
  if (!src || !idx || !dst || n < 0) return fail(SKB_EINVAL, "bad argument");
  scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, idx, n, src);
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

int skb_scatter_add_dev(double* dst, const int32_t* idx, int64_t n, const double* src, void* stream) {
  if (n == 0) return SKB_OK;
  if (!src || !idx || !dst || n < 0) return fail(SKB_EINVAL, "bad argument");
  scatter_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, idx, n, src);
  SKB_CUDA(cudaGetLastError());
  return SKB_OK;
}

}  // extern "C"
