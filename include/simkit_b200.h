/*
 * simkit_b200 -- C ABI of the B200-native per-element FEM elasticity hot path.
 *
 * This is the drop-in boundary.  The reference (otmanon/simkit,
 * pure Python) has no FFI; these entry points are what a binding for the path
 * would call.  Each one cites the reference interface it replaces
 * (paths relative to /root/reference/simkit).  INTEGRATION.md shows the ctypes
 * stubs a simkit maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; all floating point is IEEE FP64, indices int32
 *    unless stated; matrices are C-contiguous (row-major).
 *  - "host" entry points (skb_*) take HOST pointers, stage through the plan's
 *    device buffers and are synchronous on return.  "_dev" entry points take
 *    DEVICE pointers plus a cudaStream_t (as void*) and are asynchronous.
 *  - every function returns 0 on success, a negative SKB_E* code otherwise;
 *    skb_last_error() gives the message (thread-local).
 *  - dim = 2 (triangles, 3 corners) or 3 (tets, 4 corners).
 *  - mu / lam / vol arguments come with a count: 1 = scalar broadcast, t = per element.
 *    lam is ignored for SKB_MAT_ARAP (the reference's ARAP functions take no lam).
 */
#ifndef SIMKIT_B200_H
#define SIMKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_OK 0
#define SKB_EINVAL (-1)  /* bad argument (maps to ValueError)            */
#define SKB_ECUDA (-2)   /* CUDA runtime failure (maps to RuntimeError)  */
#define SKB_ENOGPU (-3)  /* no usable CUDA device                        */
#define SKB_ENOMEM (-4)

/* material ids -- energies/{stable_neo_hookean,neo_hookean,arap,stvk,linear_elasticity,fcr,
 * macklin_mueller_neo_hookean}.py */
#define SKB_MAT_STABLE_NEO_HOOKEAN 0
#define SKB_MAT_NEO_HOOKEAN 1
#define SKB_MAT_ARAP 2
#define SKB_MAT_STVK 3
#define SKB_MAT_LINEAR_ELASTICITY 4
#define SKB_MAT_FCR 5                         /* energies/fcr.py:41-302 */
#define SKB_MAT_MACKLIN_MUELLER_NEO_HOOKEAN 6 /* energies/macklin_mueller_neo_hookean.py:66-370 */

/* psd modes -- where the 1e-6 eigenvalue floor of psd_project.py:37 sits relative to vol */
#define SKB_PSD_NONE 0
#define SKB_PSD_AFTER_VOL 1  /* *_hessian_x/_u : psd_project(vol*He), e.g. stable_neo_hookean.py:533-535 */
#define SKB_PSD_BEFORE_VOL 2 /* elastic dispatcher / _z tier: vol*psd_project(He), energies/elastic.py:663-664 */

typedef struct skb_plan skb_plan;

const char* skb_last_error(void);
int skb_device_count(void);
/* library build info: "simkit_b200 <version> sm_100a" */
const char* skb_version(void);

/* ------------------------------------------------------------------ plan ---
 * Per-mesh precompute, built once on the device: element gradient operators
 * D (replaces deformation_jacobian.py:43-70), quadrature weights (volume.py:14-41),
 * the canonical CSR pattern of J^T H J (vertex adjacency (x) dim x dim, sorted),
 * the element-to-slot map and the deterministic reduction schedules.
 * T is int64 (numpy default) or int32, chosen by index_bytes (8 or 4).
 * tile_elems = elements per assembly tile (0 = default).
 */
int skb_plan_create(const double* X, const void* T, int index_bytes, int64_t n, int64_t t,
                    int dim, int device, int tile_elems, skb_plan** out);
/* Same plan from the operator's own data when the caller only holds J (the *_x / *_u tiers take
 * a prebuilt J, e.g. stable_neo_hookean.py:449-474): T recovered from J's sparsity and
 * D[e][j][a] = J[e*dim*dim + j, T[e][a]*dim] (t*dim*(dim+1) doubles).  No rest weights are known:
 * vol must be passed to every evaluation. */
int skb_plan_create_from_operator(const void* T, const double* D, int index_bytes, int64_t n,
                                  int64_t t, int dim, int device, int tile_elems, skb_plan** out);
/* Plan of ONE RANK of an element-sharded mesh (DESIGN.md "Multi-GPU"; no reference counterpart -- the
 * reference is single-process).  X, T use the rank's local vertex numbering.  The first t_active
 * elements are this rank's own; the remaining t_total - t_active "pattern-only" elements are the
 * neighbours' elements that touch vertices this rank owns: they are never evaluated, they only reserve
 * CSR slots so that the neighbours' interface block-rows can be added in place.
 * t_energy (0 = t_active): the first t_energy elements count in the energy and in the plan's internal element
 * order as one group.  A rank that RE-EVALUATES its lower neighbour's interface elements instead of receiving
 * their contributions (Shard(interface="recompute"): t_active = t_total = own + neighbour elements) passes its
 * own element count, so that every element is counted once in the all-reduced energy. */
int skb_plan_create_sharded(const double* X, const void* T, int index_bytes, int64_t n, int64_t t_active,
                            int64_t t_total, int64_t t_energy, int dim, int device, int tile_elems, skb_plan** out);
void skb_plan_destroy(skb_plan* plan);

/* sizes: n, t, dim, nnzb (block non-zeros), nnz (= nnzb*dim*dim), n_tiles, n_block_partials, n_vertex_partials */
int skb_plan_info(const skb_plan* plan, int64_t info[8]);

/* canonical scalar CSR pattern: indptr[n*dim+1], indices[nnz]  (int32, sorted) */
int skb_plan_csr_pattern(const skb_plan* plan, int32_t* indptr, int32_t* indices);
/* block view: bptr[n+1], bcol[nnzb] */
int skb_plan_block_pattern(const skb_plan* plan, int32_t* bptr, int32_t* bcol);
/* scalar element-to-slot map slot[e][a][i][b][k] (int32, t*(dim+1)*dim*(dim+1)*dim), SURVEY §7 */
int skb_plan_slot_map(const skb_plan* plan, int32_t* slot);
/* D[e][j][a] (t*dim*(dim+1)) with F_ij = sum_a D[e][j][a] x[T[e][a]][i]  (deformation_jacobian.py:58-61) */
int skb_plan_element_D(const skb_plan* plan, double* D);
/* vol[e]: signed tet volume / unsigned triangle area  (tetrahedron_volumes.py:26-27, triangle_areas.py:60-79) */
int skb_plan_volume(const skb_plan* plan, double* vol);
/* order[i] (int32, t) = the caller's index of the element the plan lists i-th.  The plan keeps its active elements in a
 * spatial (sort-tile-recursive) order of its own so that the tiles of the assembly kernel are compact (north_star
 * kernel 1, "spatially sorted"); every per-element array of this ABI is in the CALLER's order.  No reference
 * counterpart (the reference evaluates elements in the order of T, e.g. energies/stable_neo_hookean.py:531). */
int skb_plan_element_order(const skb_plan* plan, int32_t* order);
/* lumped vertex masses m[v] = sum_e rho_e vol_e / (dim+1)   (massmatrix.py:41-49) */
int skb_plan_vertex_masses(const skb_plan* plan, const double* rho, int64_t rho_n, double* m);

/* ------------------------------------------------------- global tiers ------
 * x: (n*dim) positions or displacements; Fbar: optional (t*dim*dim) per-element
 * offset J@x_bar of the _u tier (NULL for the _x tier).
 *   energy   -> *_energy_x/_u    e.g. stable_neo_hookean.py:449-474, 544-576
 *   gradient -> *_gradient_x/_u  e.g. stable_neo_hookean.py:477-503, 579-611   g: (n*dim)
 *   hessian  -> *_hessian_x/_u   e.g. stable_neo_hookean.py:506-538, 614-648   vals: (nnz) in
 *               the canonical CSR order of skb_plan_csr_pattern
 */
int skb_energy(skb_plan* plan, int material, const double* x, const double* Fbar,
               const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
               const double* vol, int64_t vol_n, double* energy);
int skb_gradient(skb_plan* plan, int material, const double* x, const double* Fbar,
                 const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                 const double* vol, int64_t vol_n, double* g);
int skb_hessian(skb_plan* plan, int material, int psd_mode, const double* x, const double* Fbar,
                const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                const double* vol, int64_t vol_n, double* vals);
/* fused gradient + Hessian (one pass over the elements); g and/or vals may be NULL */
int skb_gradient_hessian(skb_plan* plan, int material, int psd_mode, const double* x,
                         const double* Fbar, const double* mu, int64_t mu_n, const double* lam,
                         int64_t lam_n, const double* vol, int64_t vol_n, double* g, double* vals);

/* device-resident variants: every pointer is a device pointer; mu/lam/vol are
 * per-element arrays or 1-element arrays (count given); asynchronous on `stream`.
 * energy_out is a device double. */
int skb_set_materials_dev(skb_plan* plan, const double* mu, int64_t mu_n, const double* lam,
                          int64_t lam_n, const double* vol, int64_t vol_n, void* stream);
int skb_energy_dev(skb_plan* plan, int material, const double* x, const double* Fbar,
                   double* energy_out, void* stream);
int skb_gradient_hessian_dev(skb_plan* plan, int material, int psd_mode, const double* x,
                             const double* Fbar, double* g, double* vals, void* stream);
/* number of kernels launched on this plan since the last host-pointer call reset it (bench accounting) */
int skb_last_launch_count(const skb_plan* plan);
/* Per-kernel device timing for bench.py's roofline: when enabled every kernel launched on the plan is
 * bracketed by CUDA events on its own stream.  skb_kernel_times waits for them and returns the summed
 * milliseconds and launch counts per kernel kind (arrays of SKB_K_COUNT), then clears the record. */
#define SKB_K_ASSEMBLE 0        /* fused gather -> F -> P, PSD Hessian -> J^T H J -> tile reduction */
#define SKB_K_FINALIZE_BLOCKS 1 /* level-2 reduction into the CSR values */
#define SKB_K_FINALIZE_VERTS 2  /* level-2 reduction into the gradient   */
#define SKB_K_ENERGY 3
#define SKB_K_SPMV 4            /* PCG: q = (A + diag) p with the p.q partials */
#define SKB_K_PCG_VECTOR 5      /* PCG: fused vector updates / dots           */
#define SKB_K_OTHER 6
#define SKB_K_COUNT 8
int skb_kernel_timing(skb_plan* plan, int enable);
int skb_kernel_times(skb_plan* plan, double* ms, int64_t* launches);
/* interface exchange helpers on device data: dst[i] = src[idx[i]] and dst[idx[i]] += src[i]
 * (idx distinct within a call => race-free, order-independent).  Used to pack / apply the
 * interface-vertex gradient rows and Hessian block-rows that travel over NCCL. */
int skb_gather_dev(const double* src, const int32_t* idx, int64_t n, double* dst, void* stream);
int skb_scatter_add_dev(double* dst, const int32_t* idx, int64_t n, const double* src, void* stream);
/* dst[idx[i]] = src[i]: refreshes the non-owned (halo / ghost) copies of a vector after a neighbour exchange */
int skb_scatter_dev(double* dst, const int32_t* idx, int64_t n, const double* src, void* stream);
/* ---- distributed PCG / Newton building blocks (one rank of a sharded mesh; DESIGN.md "Multi-GPU").
 * All pointers are device pointers in the rank's local numbering; [v0, v1) are the vertex rows the rank owns.
 * scalars: >= 8 doubles on the device: [0] r.z  [1] r.r  [2] p.q  [3] new r.z  [4] new r.r  -- rank-local sums over the
 * owned dofs that the host all-reduces (NCCL, in place) between the calls; work: >= 3*1024 doubles of scratch.
 * Replaces, together with the halo exchange of p, spsolve at solvers/newton.py:52 for a mesh spread over GPUs. */
int skb_dist_pcg_init_dev(skb_plan* plan, const double* vals, const double* diag_add, int v0, int v1, const double* rhs,
                          double* dinv, double* x, double* r, double* z, double* p, double* scalars, double* work,
                          void* stream);
int skb_dist_spmv_dot_dev(skb_plan* plan, const double* vals, const double* diag_add, int v0, int v1, const double* p,
                          double* q, double* scalars, double* work, void* stream);
int skb_dist_pcg_update_dev(skb_plan* plan, int v0, int v1, const double* dinv, const double* p, const double* q, double* x,
                            double* r, double* z, double* scalars, double* work, void* stream);
int skb_dist_pcg_direction_dev(skb_plan* plan, int v0, int v1, const double* z, double* p, double* scalars, void* stream);
/* implicit-step vector terms on the owned dofs (same formulas as skb_newton uses on one GPU) */
int skb_dist_newton_rhs_dev(skb_plan* plan, int v0, int v1, const double* x, const double* f_ext, const double* mass,
                            const double* x_tilde, double kin_scale, const double* pin_k, const double* pin_t, double* g,
                            double* rhs, double* diag, void* stream);
/* contact springs on the owned vertices of a shard (energies/contact_springs_plane.py:245-388,
 * contact_springs_sphere.py): kind 0 plane (p, nrm), 1 sphere (p, r); gradient into g and k m_v n n^T into the diagonal
 * blocks of vals when given; this rank's share of the energy into energy_out (device) when given */
int skb_dist_contact_dev(skb_plan* plan, int v0, int v1, const double* x, int kind, double k, const double* p,
                         const double* nrm, double r, const double* w, double* g, double* vals, double* energy_out,
                         double* work, void* stream);
int skb_dist_newton_terms_dev(skb_plan* plan, int v0, int v1, const double* x, const double* dx, double s,
                              const double* f_ext, const double* mass, const double* x_tilde, double kin_scale,
                              const double* pin_k, const double* pin_t, const double* g, double* xtrial, double* out,
                              double* work, void* stream);
/* Two-level preconditioner on a sharded mesh (device pointers, stream-ordered; see skb_pcg_set_coarse).  The
 * aggregates are global; a rank adds the fine blocks of its owned rows [v0, v1) to the coarse matrix Ac
 * ((6 n_agg)^2, (3 n_agg)^2 in 2D) and its owned vertices to the restricted residual rc; the caller all-reduces
 * Ac once per solve (then inverts it) and rc once per iteration. */
int skb_dist_coarse_set(skb_plan* plan, int64_t n_agg, const int32_t* agg, const double* xrel, int v0, int v1);
int skb_dist_coarse_assemble_dev(skb_plan* plan, const double* vals, const double* diag_add, double* Ac, void* stream);
int skb_dist_coarse_invert_dev(skb_plan* plan, double* Ac, void* stream);
int skb_dist_coarse_restrict_dev(skb_plan* plan, const double* r, double* rc, void* stream);
int skb_dist_coarse_correct_dev(skb_plan* plan, int v0, int v1, const double* Ainv, double* rc, double* zc,
                                const double* r, double* z, double* p, double* scalars, int slot, double* work,
                                void* stream);
/* ---- opt-in: the distributed PCG driven from C++ with NCCL called directly (csrc/capi_nccl.cu) ----
 * The sequence of skb_dist_* steps and NCCL calls that simkit_b200/sharding.py (Shard.pcg) issues from Python, as
 * one C++ loop on the caller's stream: no reference counterpart (the reference solves on one CPU,
 * solvers/newton.py:52).  NCCL is taken with dlopen from the libnccl.so.2 already in the process; the communicator is
 * created from an id made on rank 0 (skb_nccl_unique_id, 128 bytes) that the caller broadcasts.  Pointer arrays are
 * passed as int64 addresses (device pointers of the index lists and pack buffers of every neighbour). */
int skb_nccl_unique_id(void* out, int64_t nbytes);
int skb_nccl_init(skb_plan* plan, const void* id_bytes, int64_t nbytes, int rank, int world);
int skb_nccl_set_halo(skb_plan* plan, int n_peers, const int32_t* peers, const int64_t* send_n, const int64_t* send_idx,
                      const int64_t* send_buf, const int64_t* recv_n, const int64_t* recv_idx, const int64_t* recv_buf);
int skb_nccl_finalize(skb_plan* plan);
typedef struct skb_dist_pcg_args {
  const double* vals;   /* CSR values of the owned rows (complete after the interface exchange)   */
  const double* diag;   /* diagonal added to the matrix, or NULL                                  */
  const double* rhs;
  double* x;            /* solution (local numbering; owned entries are written)                  */
  double *dinv, *r, *z, *p, *q;   /* work vectors of Shard._work()                                */
  double* s;            /* >= 8 device scalars                                                    */
  double* work;         /* >= 3 * 2048 doubles                                                    */
  double *Ac, *rc, *zc; /* coarse buffers of Shard.set_coarse_space, or NULL (block-Jacobi only)   */
  void* stream;
  double rtol;
  int32_t v0, v1;       /* owned vertex rows                                                       */
  int32_t max_iter;
  int32_t check_every;  /* iterations between host reads of r.r (default 10)                       */
  int32_t use_graph;    /* 1: full chunks of check_every iterations replay one CUDA graph (captured */
  int32_t reserved;     /*    on the plan's own stream after the first, eager chunk)               */
} skb_dist_pcg_args;
int skb_dist_pcg_native(skb_plan* plan, const skb_dist_pcg_args* args, int32_t* iters, double* relres);
/* The same solve with ONE collective per iteration (csrc/capi_pcg2.cu): single-reduction (Chronopoulos-Gear) PCG whose
 * all-reduce carries gamma = r.u, delta = w.u, r.r AND the restricted vector of the two-level preconditioner; 7 kernels,
 * one grouped ncclSend/ncclRecv (halo of u) and one ncclAllReduce of 4 + 6 n_agg doubles per iteration, chunks of
 * `check_every` iterations replayed as one CUDA graph.  Work vectors belong to the plan.  Replaces scipy's spsolve of
 * solvers/newton.py:52 on a sharded mesh; needs skb_nccl_init + skb_nccl_set_halo (and skb_dist_coarse_set for the
 * two-level preconditioner).  All vectors in the rank's local numbering; x is zeroed, its owned entries are written. */
typedef struct skb_dist_pcg2_args {
  const double* vals;   /* CSR values of the owned rows (complete after the interface exchange)   */
  const double* diag;   /* diagonal added to the matrix, or NULL                                  */
  const double* rhs;
  double* x;
  void* stream;
  double rtol;
  int32_t v0, v1;       /* owned vertex rows                                                       */
  int32_t max_iter;
  int32_t check_every;  /* iterations per chunk between host reads of the convergence flag (default 25) */
  int32_t use_graph;    /* 1: full chunks replay one CUDA graph (captured after the first, eager chunk)  */
  int32_t use_coarse;   /* 1: two-level preconditioner when the plan has a coarse space                 */
  int32_t transport;    /* 0: NCCL (grouped send/recv + all-reduce); 1: stores into the peers' HBM over NVLink  */
  int32_t reserved;     /*    through CUDA IPC mappings, flags instead of collective calls (skb_pcg2_peer_*)    */
} skb_dist_pcg2_args;
int skb_dist_pcg2(skb_plan* plan, const skb_dist_pcg2_args* args, int32_t* iters, double* relres);
/* Peer-memory transport of skb_dist_pcg2 (transport = 1): every rank allocates one slab (reduction partials of all
 * ranks, halo receive area, flags), exports its CUDA IPC handle and maps the slabs of the others; per iteration the
 * halo values of u and the 4 + 6 n_agg reduction partials are then written straight into the peers' memory by the
 * kernels that produce them (NVLink stores) and every rank sums the partials itself in rank order.  Collective set-up:
 * export on every rank, all-gather handle64 (64 bytes) and meta (1 + world int64) in rank order, import on every rank. */
int skb_pcg2_peer_export(skb_plan* plan, void* handle64, int64_t* meta, int64_t meta_len);
int skb_pcg2_peer_import(skb_plan* plan, const void* handles, const int64_t* metas);
/* host-clock breakdown of the plan's last skb_dist_pcg2, ms: set-up (incl. the all-reduce of the coarse matrix), dense
 * inverse of the coarse matrix (cuSOLVER), iterations, total */
int skb_dist_pcg2_times(skb_plan* plan, double out[4]);
/* measured FP64 FMA throughput of the device (TFLOP/s, FMA = 2 flops): the compute roofline denominator */
int skb_fp64_peak(int device, double* tflops);
/* measured FP64 tensor-core throughput (DMMA.8x8x4 = mma.sync.m8n8k4.f64, 512 flops per warp instruction):
 * the roofline denominator of the reduced-Hessian contraction */
int skb_dmma_peak(int device, double* tflops);

/* ------------------------------------------------------- element tiers -----
 * Batched per-element functions on arbitrary F (host pointers):
 *   *_energy_element_F / *_gradient_element_F / *_hessian_element_F
 *   (e.g. stable_neo_hookean.py:65-129, 132-218, 221-443).  Hessian blocks are
 *   (t, b, b), b = dim*dim, row-major F layout, unweighted and unprojected.
 */
int skb_element_energy(int material, int dim, int64_t t, const double* F, const double* mu,
                       int64_t mu_n, const double* lam, int64_t lam_n, double* psi);
int skb_element_gradient(int material, int dim, int64_t t, const double* F, const double* mu,
                         int64_t mu_n, const double* lam, int64_t lam_n, double* P);
int skb_element_hessian(int material, int dim, int64_t t, const double* F, const double* mu,
                        int64_t mu_n, const double* lam, int64_t lam_n, double* H);
/* psd_project.py:12-47 on (t, b, b) symmetric blocks; method 0 = 'proj' (floor 1e-6), 1 = 'abs' */
int skb_psd_project(int64_t t, int b, const double* H, int method, double* out);
/* svd_rv.py:8-53 / polar_svd.py:59-89: U, S (diagonal matrices), V and R = U V^T, SS = V S V^T.
 * Any output pointer may be NULL. */
int skb_svd_rv(int dim, int64_t t, const double* F, double* U, double* S, double* V);
int skb_polar(int dim, int64_t t, const double* F, double* R, double* SS);
/* rotation_gradient.py:12-75: dR/dF (t, b, b) */
int skb_rotation_gradient(int dim, int64_t t, const double* F, double* K);
/* dS/dF of the polar stretch S = R^T F, (t, d*d, d*d) with rows (m,n) of F and columns (i,j) of S
 * (stretch_gradient.py:28-54: dR/dF . F + R (x) I). */
int skb_stretch_gradient(int dim, int64_t t, const double* F, double* dSdF);

/* --------------------------------------------------------- linear solve ----
 * Block-Jacobi preconditioned CG on a matrix in the plan's canonical pattern
 * (replaces scipy.sparse.linalg.spsolve at solvers/newton.py:52).
 *   vals (nnz), diag_add (n*dim, optional: added to the diagonal, e.g. M/h^2), rhs, x (n*dim).
 * Returns iterations in *iters and the final relative residual in *relres.
 */
int skb_pcg(skb_plan* plan, const double* vals, const double* diag_add, const double* rhs,
            double rtol, int max_iter, double* x, int* iters, double* relres);
int skb_pcg_dev(skb_plan* plan, const double* vals, const double* diag_add, const double* rhs,
                double rtol, int max_iter, double* x, int* iters, double* relres, void* stream);
/* ---- device-resident CSR values (the lazy Hessian of simkit_b200/device_csr.py; csrc/capi_buffers.cu) ----------
 * The reference's `*_hessian_x` return a host scipy matrix (energies/stable_neo_hookean.py:506-538) that the caller
 * sums with its mass / penalty / contact matrices (integrators/backward_euler.py:84) and passes to spsolve
 * (solvers/newton.py:52).  These entry points keep the 8*nnz bytes of values in HBM through that chain. */
int skb_buf_alloc(int device, int64_t n, double** out);            /* n doubles of device memory          */
void skb_buf_free(int device, double* buf);
int skb_buf_copy(int device, double* dst, const double* src, int64_t n);
int skb_buf_axpy(int device, double* dst, double a, const double* src, int64_t n);     /* dst += a * src    */
int skb_buf_scale(int device, double* dst, double a, int64_t n);                       /* dst *= a          */
/* dst[pos[i]] += a * vals[i]; pos (int32, distinct, -1 = skip) and vals are HOST arrays of `count` entries */
int skb_buf_index_add(int device, double* dst, const int32_t* pos, const double* vals, int64_t count, double a);
/* vals_dev (the plan's CSR value layout) += a * diag(d), d a host vector of n*dim entries (lumped mass / penalty) */
int skb_buf_add_diagonal(skb_plan* plan, double* vals_dev, const double* diag, double a);
int skb_buf_download(int device, const double* src, int64_t n, double* host);
/* page-locked host memory for that download (pooled by the Python side) */
int skb_host_alloc(int64_t nbytes, void** out);
void skb_host_free(void* p);
/* skb_gradient_hessian with the CSR values left on the device in vals_dev (nnz doubles); g (host, optional) */
int skb_gradient_hessian_resident(skb_plan* plan, int material, int psd_mode, const double* x, const double* Fbar,
                                  const double* mu, int64_t mu_n, const double* lam, int64_t lam_n, const double* vol,
                                  int64_t vol_n, double* g, double* vals_dev);
/* skb_pcg on CSR values that are already on the device (a lazy Hessian of simkit_b200/device_csr.py): replaces the
 * spsolve of solvers/newton.py:52 without the 8*nnz-byte round trip; rhs and x are host pointers. */
int skb_pcg_vals_dev(skb_plan* plan, const double* vals_dev, const double* diag_add, const double* rhs, double rtol,
                int max_iter, double* x, int* iters, double* relres);
/* pos[i] = index of scalar entry (rows[i], cols[i]) in the plan's canonical CSR values, -1 outside the pattern: lets a
 * sparse term the caller adds to the Hessian (mass / penalty / contact matrices, e.g. integrators/backward_euler.py:84,
 * examples/interactive_demos/010_interactive_contact_plane_3D.py:104-108) be added on the device. */
int skb_plan_value_positions(skb_plan* plan, int64_t count, const int32_t* rows, const int32_t* cols, int32_t* pos);
/* Same solver for ANY sparse SPD matrix in scalar CSR form (what newton_solver receives from user
 * callables, solvers/newton.py:51-52); block = size of the Jacobi blocks (1, 2 or 3; n % block == 0). */
int skb_csr_pcg(int64_t n, const int32_t* indptr, const int32_t* indices, const double* vals, int block,
                const double* rhs, double rtol, int max_iter, double* x, int* iters, double* relres);
/* Dense LU solve with partial pivoting (replaces scipy.linalg.solve at solvers/newton.py:54, the
 * reduced-space Newton system); A (n, n) row-major.  SKB_EINVAL if singular. */
int skb_dense_solve(int64_t n, const double* A, const double* b, double* x);
/* y = (A + diag(diag_add)) x   on device data */
/* Plane contact springs (energies/contact_springs_plane.py:245-388): E = k/2 sum_{v: n.(x_v - p) < 0} m_v (n.(x_v - p))^2.
 * X: (nv*dim); weights: (nv) m_v or NULL (1).  Outputs (each may be NULL): energy, grad (nv*dim), blocks (nv*dim*dim,
 * the Hessian's diagonal blocks k m_v n n^T, zero for vertices above the plane), under (nv, 1 = contacting). */
int skb_contact_springs_plane(int dim, int64_t nv, const double* X, double k, const double* p, const double* n,
                              const double* weights, double* energy, double* grad, double* blocks, int32_t* under);
/* Sphere contact springs (energies/contact_springs_sphere.py:242-360): vertices with |x_v - p| < r, normal
 * n_v = (x_v - p)/|x_v - p| held fixed in the derivatives, E = k/2 sum m_v (|x_v - p| - r)^2; same outputs. */
int skb_contact_springs_sphere(int dim, int64_t nv, const double* X, double k, const double* p, double r,
                               const double* weights, double* energy, double* grad, double* blocks, int32_t* under);
int skb_newton_set_contact_sphere(skb_plan* plan, double k, const double* p, double r, const double* weights);
/* Adds the same term to every following skb_newton on this plan (energy in the line search, gradient, Hessian
 * blocks straight into the CSR values on the device).  k <= 0 or p == NULL removes it. */
int skb_newton_set_contact_plane(skb_plan* plan, double k, const double* p, const double* n, const double* weights);
/* energies/quadratic.py:15-70: energy = 1/2 x^T Q x + b^T x (quadratic_energy :15-34), grad = Q x + b
 * (quadratic_gradient :37-54) for an n x n CSR matrix Q (int32 indptr / indices, host pointers).  b may be NULL (0);
 * energy or grad may be NULL. */
int skb_quadratic(int64_t n, const int32_t* indptr, const int32_t* indices, const double* vals, const double* b,
                  const double* x, double* energy, double* grad);
/* Adds the same term to every following skb_newton on this plan: energy in the line search, gradient, and the
 * Hessian Q (quadratic_hessian :57-70) added into the CSR values on the device.  Q: (n*dim)^2 CSR without duplicate
 * entries, symmetric, every entry inside the mesh's CSR pattern (vertex adjacency (x) dim x dim; SKB_EINVAL
 * otherwise) -- e.g. dirichlet_penalty.py:65-142, a mass or Laplacian regulariser.  indptr == NULL removes it. */
int skb_newton_set_quadratic(skb_plan* plan, const int32_t* indptr, const int32_t* indices, const double* vals,
                             const double* b);
/* Two-level preconditioner of the plan's PCG (skb_pcg*, skb_newton): block-Jacobi plus a coarse correction on the
 * rigid-body modes of vertex aggregates.  agg: (n) aggregate id of every vertex in [0, n_agg); xrel: (n*dim) vertex
 * position minus the centre of its aggregate.  n_agg = 0 removes it.  The reference solves the Newton system
 * directly (solvers/newton.py:52); the preconditioner only changes the CG iteration count, not the solution. */
int skb_pcg_set_coarse(skb_plan* plan, int64_t n_agg, const int32_t* agg, const double* xrel);
int skb_spmv_dev(skb_plan* plan, const double* vals, const double* diag_add, const double* x,
                 double* y, void* stream);

/* ------------------------------------------------------- Newton step -------
 * Device-resident implicit step for the framework's own energies: replaces
 * integrators/backward_euler.py:27-91 / bdf2.py:31-101 + solvers/newton.py:7-75 +
 * backtracking_line_search.py:8-66 when all three callables are this library's.
 *   total energy  V(x) = elastic(x) - f_ext . x + 1/2 sum_i pin_k[i] (x[i] - pin_target[i])^2
 *                        + kin_scale/2 (x - x_tilde)^T M (x - x_tilde),   kin_scale = c/h^2
 *   (the pin term is the diagonal dirichlet_penalty.py:65-142 / quadratic.py:15-70 energy)
 * Host pointers in, host pointers out; everything in between stays on the device.
 */
typedef struct skb_newton_opts {
  int material;
  int psd_mode;
  int max_iter;        /* Newton iterations (newton.py default 1)           */
  int do_line_search;  /* backtracking_line_search.py defaults when 1       */
  double tolerance;    /* stop when |alpha*dx| < tolerance (newton.py:69)   */
  double ls_alpha;     /* 0.01 */
  double ls_beta;      /* 0.5  */
  int ls_max_iter;     /* 100  */
  double ls_threshold; /* 1e-12 */
  double pcg_rtol;     /* relative residual for the linear solve            */
  int pcg_max_iter;
} skb_newton_opts;

typedef struct skb_newton_info {
  int iters;           /* index of the last Newton iteration (newton.py:67) */
  int pcg_iters_total;
  double last_alpha;
  double last_step_norm;
  double last_pcg_relres;
  double alphas[64];
} skb_newton_info;

/* x0: start (n*dim); x_tilde: inertial target or NULL (then no kinetic term);
 * mass: lumped per-dof masses (n*dim) or NULL; kin_scale = c/h^2; f_ext: (n*dim) or NULL;
 * pin_k / pin_target: per-dof penalty stiffness and target (n*dim) or NULL;
 * x_out: (n*dim). Materials must have been set with skb_set_materials / *_dev. */
int skb_set_materials(skb_plan* plan, const double* mu, int64_t mu_n, const double* lam,
                      int64_t lam_n, const double* vol, int64_t vol_n);
int skb_newton(skb_plan* plan, const skb_newton_opts* opts, const double* x0,
               const double* x_tilde, const double* mass, double kin_scale, const double* f_ext,
               const double* pin_k, const double* pin_target, double* x_out, skb_newton_info* info);

/* ------------------------------------------------------- reduced tier ------
 * Reduced Hessian / gradient with a dense operator:
 *   F = JB z + Jx0,  Hr = JB^T blockdiag(vol * psd(He)) JB   (energies/elastic.py:749-782,
 *   or the *_hessian_u tier called with a dense J).  JB: (t*b, r) row-major, z: (r), Jx0: (t*b).
 * Hr: (r, r), gr: (r).  Either output may be NULL.
 */
int skb_reduced_gradient_hessian(int material, int psd_mode, int dim, int64_t t, int64_t r,
                                 const double* JB, const double* Jx0, const double* z,
                                 const double* mu, int64_t mu_n, const double* lam, int64_t lam_n,
                                 const double* vol, int64_t vol_n, double* energy, double* gr,
                                 double* Hr);
/* Same, forming JB rows on the fly from the mesh plan and a dense basis B (n*dim, r):
 * Hr = B^T J^T H J B without materialising JB.  x0: (n*dim) offset. */
int skb_reduced_hessian_from_basis(skb_plan* plan, int material, int psd_mode, int64_t r,
                                   const double* B, const double* x0, const double* z,
                                   double* energy, double* gr, double* Hr);
/* Keeps the basis B (n*dim, r), row-major, resident on the plan's device; skb_reduced_hessian_from_basis may then
 * be called with B = NULL (the reference rebuilds JB = G J B once per simulation in ElasticEnergyZPrecomp,
 * energies/elastic.py:214-222; this is the device-side counterpart).  B = NULL or r <= 0 releases it. */
int skb_plan_set_basis(skb_plan* plan, int64_t r, const double* B);
/* Measurement hook (bench.py, no reference counterpart): CUDA-event times in ms of the last reduced call on
 * this thread's device: out[0] element pass (F -> He, P, psi), out[1] the B^T H B contraction kernel
 * (FP64 DMMA tiles), out[2] whole device section including the partial sums. */
int skb_reduced_last_times(double out[3]);

/* fast_sandwich_transform_clustered.py:15-158.
 * A: (m1, b*t) dense row-major, B: (b*t, m2) dense row-major, l: (t) cluster labels,
 * ARBs: (m1, m2, c, dim, dim).  eval: out (m1, m2) = sum ARBs * r, r: (c, dim, dim). */
int skb_fst_precompute(int dim, int64_t t, int64_t m1, int64_t m2, int64_t n_clusters,
                       const double* A, const double* B, const int32_t* l, double* ARBs);
int skb_fst_eval(int dim, int64_t m1, int64_t m2, int64_t n_clusters, const double* ARBs,
                 const double* r, double* out);

/* ---- subspace construction (csrc/capi_subspace.cu) ---------------------------------------------
 * Thin Householder QR of a row-major (n x r) matrix as numpy.linalg.qr returns it (orthonormalize.py:42), and
 * G = A^T diag(w) B (row-major r x s; w may be NULL) for B^T M B / B^T M y of project_into_subspace.py:49-53.
 * Host pointers; cuSOLVER geqrf + orgqr and cuBLAS gemm (plain library calls on a one-off set-up step). */
int skb_qr_thin(int64_t n, int64_t r, const double* A, double* Q, double* R);
int skb_weighted_gram(int64_t n, int64_t r, int64_t s, const double* A, const double* w, const double* B, double* G);

/* ---- spectral clustering / cubature (csrc/capi_cluster.cu) --------------------------------------
 * skb_average_onto_simplex: At (t x p) = mean over the K corners T (t x K, int32) of the per-vertex rows A (n x p)
 *   (average_onto_simplex.py:8-37).
 * skb_kmeans2_pp: scipy.cluster.vq.kmeans2(data, k, iter, minit="++") as spectral_clustering.py:46 calls it: `first` is
 *   the uniformly drawn first centre row and `uniforms` the k - 1 numbers in [0, 1) of the k-means++ draws, taken by the
 *   caller from scipy's own generator in scipy's order; centroids (k x p), labels (n, int32: from the LAST assignment,
 *   i.e. before the last centre update, as scipy returns them); n_empty counts empty clusters met (scipy warns).
 * skb_cubature_pick: lI[j] = first row of data nearest to centroid j, mc[j] = sum of vol over the rows labelled j
 *   (spectral_cubature.py:63-66: pairwise_distance + argmin, bincount).
 * Host pointers, row-major. */
int skb_average_onto_simplex(int64_t n, int64_t p, int64_t t, int K, const double* A, const int32_t* T, double* At);
int skb_kmeans2_pp(int64_t n, int64_t p, int64_t k, int iters, int64_t first, const double* uniforms, const double* data,
                   double* centroids, int32_t* labels, int32_t* n_empty);
int skb_cubature_pick(int64_t n, int64_t p, int64_t k, const double* data, const double* centroids, const int32_t* labels,
                      const double* vol, int64_t* lI, double* mc);

#ifdef __cplusplus
}
#endif
#endif /* SIMKIT_B200_H */
