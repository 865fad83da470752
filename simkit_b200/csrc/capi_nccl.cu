// C ABI, part 6 (opt-in): the distributed PCG of capi_dist.cu driven from C++ instead of Python.
//
// simkit_b200/sharding.py issues, per CG iteration, four NCCL calls through torch.distributed and about ten
// ctypes calls; at 8 GPUs that host work (~0.4 ms) is four times the GPU work of the iteration (DESIGN.md
// section 8, item 6).  Here the same sequence -- halo exchange of p, SpMV + p.q, all-reduce, fused update,
// [coarse correction: restriction, all-reduce, dense apply], all-reduce, direction -- is issued from one C++
// loop on the caller's stream with NCCL called directly: no Python between the steps, one host
// synchronisation every `check_every` iterations, and a sequence that a CUDA graph can capture.
//
// NCCL is not linked: its entry points are taken with dlopen(RTLD_NOLOAD)/dlsym from the libnccl.so.2 that is
// already in the process (the one torch loaded), so the default library has no NCCL dependency and two NCCL
// versions can never meet in one process.  The communicator is this library's own (ncclCommInitRank with an id that the
// Python side broadcasts through torch.distributed).
#include "nccl_api.cuh"
#include "solver.cuh"
#include "coarse.cuh"

using namespace skb;

static_assert(sizeof(skb_dist_pcg_args) == 152, "skb_dist_pcg_args layout (mirrored by ctypes in simkit_b200/_lib.py)");

#if defined(SKB_HAVE_NCCL_H)
namespace skb {

NcclApi& nccl() {

  static NcclApi a;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (a.ok) return a;   // a failed attempt is retried: torch may have been imported since
  [&] {
    // ONLY the copy that is already in the process (the one torch loaded; matched by SONAME).  Loading another
    // libnccl.so.2 here would make a later `import torch` bind to it and fail on the symbols its own build expects.
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!a.handle) {
      a.why = "libnccl.so.2 is not loaded in this process (import torch before enabling the native NCCL path)";
      return;
    }
    bool all = true;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(a.handle, name);
      if (!p) {
        all = false;
        a.why = std::string("libnccl.so.2 lacks ") + name;
      }
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = all;
  }();
  return a;
}

// refreshes the non-owned copies of v from their owners (Shard.halo_exchange): pack, one grouped send/recv, unpack
int halo_exchange(DistNative& d, double* v, cudaStream_t st) {
  NcclApi& n = nccl();
  for (const Halo& h : d.halo)
    if (h.ns > 0) {
      const int rc = skb_gather_dev(v, h.sidx, h.ns, h.sbuf, st);
      if (rc) return rc;
    }
  SKB_NCCL(n.GroupStart());
  for (const Halo& h : d.halo)
    if (h.ns > 0) SKB_NCCL(n.Send(h.sbuf, (size_t)h.ns, ncclDouble, h.peer, d.comm, st));
  for (const Halo& h : d.halo)
    if (h.nr > 0) SKB_NCCL(n.Recv(h.rbuf, (size_t)h.nr, ncclDouble, h.peer, d.comm, st));
  SKB_NCCL(n.GroupEnd());
  for (const Halo& h : d.halo)
    if (h.nr > 0) {
      const int rc = skb_scatter_dev(v, h.ridx, h.nr, h.rbuf, st);
      if (rc) return rc;
    }
  return SKB_OK;
}

int all_reduce(DistNative& d, double* buf, size_t count, cudaStream_t st) {
  SKB_NCCL(nccl().AllReduce(buf, buf, count, ncclDouble, ncclSum, d.comm, st));
  return SKB_OK;
}

}  // namespace skb
#endif  // SKB_HAVE_NCCL_H

extern "C" {

#if !defined(SKB_HAVE_NCCL_H)
#define SKB_NEED_NCCL return fail(SKB_EINVAL, "built without nccl.h: the native NCCL path is not available");
#else
#define SKB_NEED_NCCL                                         \
  if (!nccl().ok) return fail(SKB_EINVAL, "NCCL unavailable: " + nccl().why);
#endif

int skb_nccl_unique_id(void* out, int64_t nbytes) {
  SKB_NEED_NCCL
#if defined(SKB_HAVE_NCCL_H)
  if (!out || nbytes < (int64_t)sizeof(ncclUniqueId)) return fail(SKB_EINVAL, "the id buffer needs 128 bytes");
  ncclUniqueId id;
  SKB_NCCL(nccl().GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return SKB_OK;
#endif
}

int skb_nccl_init(skb_plan* pl, const void* id_bytes, int64_t nbytes, int rank, int world) {
  SKB_NEED_NCCL
#if defined(SKB_HAVE_NCCL_H)
  if (!pl || !id_bytes || nbytes < (int64_t)sizeof(ncclUniqueId) || rank < 0 || rank >= world)
    return fail(SKB_EINVAL, "bad argument");
  SKB_CUDA(cudaSetDevice(pl->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  SKB_NCCL(nccl().CommInitRank(&comm, world, id, rank));
  if (!pl->dist) pl->dist = new DistNative();
  DistNative& d = *state_of(pl);
  if (d.comm) nccl().CommDestroy(d.comm);
  d.comm = comm;
  d.rank = rank;
  d.world = world;
  return SKB_OK;
#endif
}

int skb_nccl_set_halo(skb_plan* pl, int n_peers, const int32_t* peers, const int64_t* send_n, const int64_t* send_idx,
                      const int64_t* send_buf, const int64_t* recv_n, const int64_t* recv_idx, const int64_t* recv_buf) {
  SKB_NEED_NCCL
#if defined(SKB_HAVE_NCCL_H)
  DistNative* d = pl ? state_of(pl) : nullptr;
  if (!d || !d->comm) return fail(SKB_EINVAL, "skb_nccl_init first");
  if (n_peers < 0 || (n_peers > 0 && (!peers || !send_n || !send_idx || !send_buf || !recv_n || !recv_idx || !recv_buf)))
    return fail(SKB_EINVAL, "null argument");
  d->halo.clear();
  for (int i = 0; i < n_peers; ++i) {
    Halo h;
    h.peer = peers[i];
    h.ns = send_n[i];
    h.nr = recv_n[i];
    h.sidx = reinterpret_cast<const int32_t*>((uintptr_t)send_idx[i]);
    h.ridx = reinterpret_cast<const int32_t*>((uintptr_t)recv_idx[i]);
    h.sbuf = reinterpret_cast<double*>((uintptr_t)send_buf[i]);
    h.rbuf = reinterpret_cast<double*>((uintptr_t)recv_buf[i]);
    if (h.peer < 0 || h.peer >= d->world || h.peer == d->rank || h.ns < 0 || h.nr < 0 || (h.ns > 0 && (!h.sidx || !h.sbuf)) ||
        (h.nr > 0 && (!h.ridx || !h.rbuf)))
      return fail(SKB_EINVAL, "bad halo list");
    d->halo.push_back(h);
  }
  return SKB_OK;
#endif
}

int skb_nccl_finalize(skb_plan* pl) {
#if defined(SKB_HAVE_NCCL_H)
  // called by skb_plan_destroy as well: the communicator, the halo lists (raw pointers into the caller's tensors) and
  // the solver state can never outlive the plan or be inherited by another plan allocated at the same address
  DistNative* d = state_of(pl);
  if (d) {
    if (d->pcg2) pcg2_destroy(d->pcg2);
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    delete d;
    pl->dist = nullptr;
  }
#endif
  return SKB_OK;
}

int skb_dist_pcg_native(skb_plan* pl, const skb_dist_pcg_args* a, int32_t* iters, double* relres) {
  SKB_NEED_NCCL
#if defined(SKB_HAVE_NCCL_H)
  if (!pl || !a || !iters || !relres) return fail(SKB_EINVAL, "null argument");
  DistNative* dp = state_of(pl);
  if (!dp || !dp->comm) return fail(SKB_EINVAL, "skb_nccl_init first");
  DistNative& d = *dp;
  if (!a->vals || !a->rhs || !a->x || !a->dinv || !a->r || !a->z || !a->p || !a->q || !a->s || !a->work)
    return fail(SKB_EINVAL, "null work vector");
  SKB_CUDA(cudaSetDevice(pl->device));
  SKB_TRY
  // graph mode: everything runs on the plan's own (capturable, non-default) stream, ordered after the caller's
  // stream on entry and before it on exit
  const bool use_graph = a->use_graph != 0;
  cudaStream_t caller = (cudaStream_t)a->stream;
  cudaStream_t st = use_graph ? pl->stream : caller;
  void* sp = (void*)st;
  cudaEvent_t ev = nullptr;
  cudaGraphExec_t gexec = nullptr;
  // every exit path (also the early `return rc` ones) re-joins the caller's stream and releases the event and the graph
  struct Guard {
    cudaEvent_t& ev;
    cudaGraphExec_t& gexec;
    cudaStream_t st, caller;
    bool use_graph;
    ~Guard() {
      if (gexec) cudaGraphExecDestroy(gexec);
      if (use_graph && ev) {
        cudaEventRecord(ev, st);
        cudaStreamWaitEvent(caller, ev, 0);
      }
      if (ev) cudaEventDestroy(ev);
    }
  } guard{ev, gexec, st, caller, use_graph};
  if (use_graph) {
    SKB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SKB_CUDA(cudaEventRecord(ev, caller));
    SKB_CUDA(cudaStreamWaitEvent(st, ev, 0));
  }
  const int v0 = a->v0, v1 = a->v1;
  double* s = a->s;
  int rc;
  bool coarse = a->Ac && a->rc && a->zc && pl->coarse;
  const size_t nc = coarse ? (size_t)(pl->d.dim == 3 ? 6 : 3) * pl->coarse->n_agg : 0;
  auto coarse_correct = [&](double* pvec, int slot) -> int {
    // z += P Ainv P^T r with the restriction all-reduced; s[slot] = r.z over the owned dofs (Shard._coarse_correct)
    int r2 = skb_dist_coarse_restrict_dev(pl, a->r, a->rc, sp);
    if (r2) return r2;
    r2 = all_reduce(d, a->rc, nc, st);
    if (r2) return r2;
    return skb_dist_coarse_correct_dev(pl, v0, v1, a->Ac, a->rc, a->zc, a->r, a->z, pvec, s, slot, a->work, sp);
  };
  // one CG iteration: halo exchange of p, SpMV + p.q, all-reduce, fused update, [coarse correction], all-reduce, direction
  auto iteration = [&]() -> int {
    int r2;
    if ((r2 = halo_exchange(d, a->p, st))) return r2;
    if ((r2 = skb_dist_spmv_dot_dev(pl, a->vals, a->diag, v0, v1, a->p, a->q, s, a->work, sp))) return r2;
    if ((r2 = all_reduce(d, s + 2, 1, st))) return r2;
    if ((r2 = skb_dist_pcg_update_dev(pl, v0, v1, a->dinv, a->p, a->q, a->x, a->r, a->z, s, a->work, sp))) return r2;
    if (coarse && (r2 = coarse_correct(nullptr, 3))) return r2;
    if ((r2 = all_reduce(d, s + 3, 2, st))) return r2;
    return skb_dist_pcg_direction_dev(pl, v0, v1, a->z, a->p, s, sp);
  };
  if (coarse) {
    // coarse matrix of this system: owned fine blocks per rank, summed over the ranks, inverted by every rank
    if ((rc = skb_dist_coarse_assemble_dev(pl, a->vals, a->diag, a->Ac, sp))) return rc;
    if ((rc = all_reduce(d, a->Ac, nc * nc, st))) return rc;
    if (skb_dist_coarse_invert_dev(pl, a->Ac, sp) != 0) coarse = false;  // degenerate aggregate: block-Jacobi for this solve
  }
  if ((rc = skb_dist_pcg_init_dev(pl, a->vals, a->diag, v0, v1, a->rhs, a->dinv, a->x, a->r, a->z, a->p, s, a->work, sp))) return rc;
  if (coarse && (rc = coarse_correct(a->p, 0))) return rc;
  if ((rc = all_reduce(d, s, 2, st))) return rc;
  double h2[2];
  SKB_CUDA(cudaMemcpyAsync(h2, s, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  SKB_CUDA(cudaStreamSynchronize(st));
  const double bb = h2[1];
  *iters = 0;
  *relres = 0.0;
  const int every = a->check_every > 0 ? a->check_every : 10;
  int it = 0;
  double rr = bb;
  while (bb > 0.0 && it < a->max_iter) {
    const int nrun = (a->max_iter - it) < every ? (a->max_iter - it) : every;
    if (use_graph && it > 0 && nrun == every) {
      // the first chunk ran eagerly (NCCL has set up its peer connections); full chunks from here on replay one graph
      if (!gexec) {
        const bool timing = pl->timing;
        pl->timing = false;   // no event records inside the capture
        cudaGraph_t graph = nullptr;
        SKB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        rc = SKB_OK;
        for (int k = 0; k < nrun && rc == SKB_OK; ++k) rc = iteration();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);   // always ends the capture, also after a failed step
        pl->timing = timing;
        if (rc) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        if (ce != cudaSuccess) return fail(SKB_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
        const cudaError_t ci = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) return fail(SKB_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ci));
      }
      SKB_CUDA(cudaGraphLaunch(gexec, st));
    } else {
      for (int k = 0; k < nrun; ++k)
        if ((rc = iteration())) return rc;
    }
    it += nrun;
    SKB_CUDA(cudaMemcpyAsync(h2, s, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SKB_CUDA(cudaStreamSynchronize(st));
    rr = h2[1];
    if (!(rr > a->rtol * a->rtol * bb)) break;
  }
  if (!(bb > 0.0)) return SKB_OK;
  *iters = it;
  *relres = sqrt(rr / bb);
  return SKB_OK;
  SKB_CATCH
#endif
}

}  // extern "C"
