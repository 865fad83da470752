// Per-mesh plan: everything that depends only on (X, T), built once.
//
// HBM layout (all arrays live on the device; SoA so that a warp's loads coalesce):
//   T32   [t][K]        int32   element corners, K = dim+1               (16 B/tet)
//   Dm    [dim*dim][t]  f64     D[j][a], a = 1..dim (corner 0 = -sum)    (72 B/tet)
//   vol0  [t]           f64     rest quadrature weights
//   bptr  [n+1], bcol[nnzb], brow[nnzb]  canonical block pattern (vertex adjacency, sorted)
//   bslot [t][K][K]     int32   element-to-block-slot map
// plus two deterministic reduction schedules (blocks and vertices), see ReduceSched.
//
// The builder is written against thrust with a backend tag so that the exact
// same code runs with thrust::device in the product and thrust::host inside the
// CPU test harness (tests/host_harness.cu) of the GPU-less build container.
#pragma once
#include <stdint.h>
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/for_each.h>
#include <thrust/gather.h>
#include <thrust/scatter.h>
#include <thrust/host_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/transform.h>
#include <thrust/unique.h>

#include "smallmat.cuh"

namespace skb {

struct HostBackend {
  template <class T>
  using vec = thrust::host_vector<T>;
  static auto policy() { return thrust::host; }
};
struct DeviceBackend {
  template <class T>
  using vec = thrust::device_vector<T>;
  static auto policy() { return thrust::device; }
};

// Deterministic two-level reduction schedule ("who sums what, in which order").
//
// Contributions c (one per (element, corner[, corner]) pair) are grouped by
// (slot, tile) where tile = element / tile_elems.  Each group is a *tile-slot*;
// level 1 (inside the assembly kernel, from shared memory) sums a tile-slot's
// contributions in ascending c and writes ONE partial record; level 2 (the
// finalize kernel) sums the partial records of a slot in ascending tile order.
// Partial records are numbered slot-major, so level 2 reads a contiguous run.
// No atomics anywhere: the result is bitwise reproducible.
struct ReduceSchedView {
  int n_slots;
  int n_ts;            // number of tile-slots == number of partial records
  int n_contrib;
  const int* tl_ptr;   // [n_tiles+1]  tile -> its tile-slot entries (tile-major list)
  const int* tl_q;     // [n_ts]       partial record index of each entry
  const int* tl_cptr;  // [n_ts+1]     entry -> contribution range in tc_src
  const uint16_t* tc_src;  // [n_contrib] packed (local element, corner[s])
  const int* sp_ptr;   // [n_slots+1]  slot -> contiguous range of partial records
};

template <class B>
struct ReduceSched {
  int n_slots = 0, n_ts = 0, n_contrib = 0;
  typename B::template vec<int> tl_ptr, tl_q, tl_cptr, sp_ptr;
  typename B::template vec<uint16_t> tc_src;
  ReduceSchedView view() const {
    ReduceSchedView v;
    v.n_slots = n_slots;
    v.n_ts = n_ts;
    v.n_contrib = n_contrib;
    v.tl_ptr = thrust::raw_pointer_cast(tl_ptr.data());
    v.tl_q = thrust::raw_pointer_cast(tl_q.data());
    v.tl_cptr = thrust::raw_pointer_cast(tl_cptr.data());
    v.tc_src = thrust::raw_pointer_cast(tc_src.data());
    v.sp_ptr = thrust::raw_pointer_cast(sp_ptr.data());
    return v;
  }
};

struct PlanView {
  int dim, K;
  int n, t;
  int tile_elems, n_tiles;
  int nnzb;
  const int* T32;
  const double* Dm;
  const double* vol0;
  const int* bptr;
  const int* bcol;
  const int* brow;
  const int* bslot;
  ReduceSchedView blocks;
  ReduceSchedView verts;
};

template <class B>
struct PlanData {
  int dim = 0, K = 0, n = 0, t = 0, tile_elems = 0, n_tiles = 0, nnzb = 0;
  bool has_vol0 = false;
  typename B::template vec<int> T32, bptr, bcol, brow, bslot;
  typename B::template vec<double> Dm, vol0;
  ReduceSched<B> blocks, verts;
  PlanView view() const {
    PlanView v;
    v.dim = dim; v.K = K; v.n = n; v.t = t;
    v.tile_elems = tile_elems; v.n_tiles = n_tiles; v.nnzb = nnzb;
    v.T32 = thrust::raw_pointer_cast(T32.data());
    v.Dm = thrust::raw_pointer_cast(Dm.data());
    v.vol0 = thrust::raw_pointer_cast(vol0.data());
    v.bptr = thrust::raw_pointer_cast(bptr.data());
    v.bcol = thrust::raw_pointer_cast(bcol.data());
    v.brow = thrust::raw_pointer_cast(brow.data());
    v.bslot = thrust::raw_pointer_cast(bslot.data());
    v.blocks = blocks.view();
    v.verts = verts.view();
    return v;
  }
};

// ------------------------------------------------------------------ functors
struct HeadFlag64 {
  const uint64_t* k;
  SKB_HD int operator()(int i) const { return (i == 0 || k[i] != k[i - 1]) ? 1 : 0; }
};

// key of a block contribution c = e*K*K + a*K + b : (row vertex, col vertex)
struct BlockKey {
  const int* T;
  int K;
  SKB_HD uint64_t operator()(int c) const {
    int e = c / (K * K);
    int ab = c - e * K * K;
    int a = ab / K, b = ab - a * K;
    return ((uint64_t)(uint32_t)T[e * K + a] << 32) | (uint32_t)T[e * K + b];
  }
};
// key of a vertex contribution c = e*K + a : vertex
struct VertKey {
  const int* T;
  SKB_HD uint64_t operator()(int c) const { return (uint64_t)(uint32_t)T[c]; }
};

// (slot, tile) key of the i-th contribution in slot-sorted order
struct SlotTileKey {
  const int* slot_sorted;  // slot index per sorted position
  const int* c_sorted;     // contribution id per sorted position
  int per_elem;            // contributions per element
  int tile_elems;
  SKB_HD uint64_t operator()(int i) const {
    int e = c_sorted[i] / per_elem;
    return ((uint64_t)(uint32_t)slot_sorted[i] << 32) | (uint32_t)(e / tile_elems);
  }
};
struct TileQKey {
  const int* q_sorted;
  const int* c_sorted;
  int per_elem;
  int tile_elems;
  SKB_HD uint64_t operator()(int i) const {
    int e = c_sorted[i] / per_elem;
    return ((uint64_t)(uint32_t)(e / tile_elems) << 32) | (uint32_t)q_sorted[i];
  }
};
struct PackSrc {
  const int* c;
  int per_elem;
  int tile_elems;
  SKB_HD uint16_t operator()(int i) const {
    int e = c[i] / per_elem;
    int rem = c[i] - e * per_elem;
    return (uint16_t)((e % tile_elems) * per_elem + rem);
  }
};
struct MinusOne {
  SKB_HD int operator()(int x) const { return x - 1; }
};
struct Hi32 {
  SKB_HD int operator()(uint64_t k) const { return (int)(k >> 32); }
};
struct Lo32 {
  SKB_HD int operator()(uint64_t k) const { return (int)(k & 0xffffffffu); }
};
struct IsHead {
  const int* flag;
  SKB_HD bool operator()(int i) const { return flag[i] != 0; }
};

// Builds a ReduceSched from contributions already sorted by slot (stable, so
// ascending contribution id inside a slot).  slot_sorted / c_sorted have
// n_contrib entries.
template <class B>
void build_sched(ReduceSched<B>& s, int n_slots, int n_tiles, int per_elem, int tile_elems,
                 const typename B::template vec<int>& slot_sorted,
                 const typename B::template vec<int>& c_sorted) {
  auto pol = B::policy();
  const int nc = (int)c_sorted.size();
  s.n_slots = n_slots;
  s.n_contrib = nc;
  using IV = typename B::template vec<int>;
  using KV = typename B::template vec<uint64_t>;
  thrust::counting_iterator<int> it0(0);

  // (slot, tile) keys are already non-decreasing: c ascending inside a slot => tile ascending
  KV key(nc);
  thrust::transform(pol, it0, it0 + nc, key.begin(),
                    SlotTileKey{thrust::raw_pointer_cast(slot_sorted.data()),
                                thrust::raw_pointer_cast(c_sorted.data()), per_elem, tile_elems});
  IV flag(nc), q_sorted(nc);
  thrust::transform(pol, it0, it0 + nc, flag.begin(), HeadFlag64{thrust::raw_pointer_cast(key.data())});
  thrust::inclusive_scan(pol, flag.begin(), flag.end(), q_sorted.begin());
  thrust::transform(pol, q_sorted.begin(), q_sorted.end(), q_sorted.begin(), MinusOne());
  s.n_ts = nc ? (int)q_sorted.back() + 1 : 0;

  // sp_ptr[slot] = first partial record of the slot: lower_bound over the slot of each record
  {
    IV head_pos(s.n_ts);
    thrust::copy_if(pol, it0, it0 + nc, head_pos.begin(), IsHead{thrust::raw_pointer_cast(flag.data())});
    IV rec_slot(s.n_ts);
    thrust::gather(pol, head_pos.begin(), head_pos.end(), slot_sorted.begin(), rec_slot.begin());
    s.sp_ptr.resize(n_slots + 1);
    thrust::lower_bound(pol, rec_slot.begin(), rec_slot.end(), it0, it0 + n_slots + 1, s.sp_ptr.begin());
  }

  // tile-major lists: stable sort by (tile, q)
  KV key2(nc);
  thrust::transform(pol, it0, it0 + nc, key2.begin(),
                    TileQKey{thrust::raw_pointer_cast(q_sorted.data()),
                             thrust::raw_pointer_cast(c_sorted.data()), per_elem, tile_elems});
  IV c2 = c_sorted;
  thrust::stable_sort_by_key(pol, key2.begin(), key2.end(), c2.begin());
  thrust::transform(pol, it0, it0 + nc, flag.begin(), HeadFlag64{thrust::raw_pointer_cast(key2.data())});
  IV head_pos(s.n_ts);
  thrust::copy_if(pol, it0, it0 + nc, head_pos.begin(), IsHead{thrust::raw_pointer_cast(flag.data())});
  s.tl_cptr.resize(s.n_ts + 1);
  thrust::copy(pol, head_pos.begin(), head_pos.end(), s.tl_cptr.begin());
  s.tl_cptr[s.n_ts] = nc;
  KV head_key(s.n_ts);
  thrust::gather(pol, head_pos.begin(), head_pos.end(), key2.begin(), head_key.begin());
  s.tl_q.resize(s.n_ts);
  thrust::transform(pol, head_key.begin(), head_key.end(), s.tl_q.begin(), Lo32());
  IV head_tile(s.n_ts);
  thrust::transform(pol, head_key.begin(), head_key.end(), head_tile.begin(), Hi32());
  s.tl_ptr.resize(n_tiles + 1);
  thrust::lower_bound(pol, head_tile.begin(), head_tile.end(), it0, it0 + n_tiles + 1, s.tl_ptr.begin());
  s.tc_src.resize(nc);
  thrust::transform(pol, it0, it0 + nc, s.tc_src.begin(),
                    PackSrc{thrust::raw_pointer_cast(c2.data()), per_elem, tile_elems});
}

// ---------------------------------------------------------------- geometry --
// D = (H (X_e^T H)^-1)^T, stored without corner 0 (its column is minus the sum
// of the others), and the rest quadrature weight.
template <int D>
struct GeomFunctor {
  const double* X;
  const int* T;
  double* Dm;
  double* vol;
  int t;
  SKB_HD void operator()(int e) const {
    constexpr int K = D + 1;
    Mat<D> Ed;  // edge matrix: column k-1 = x_k - x_0
#pragma unroll
    for (int k = 1; k < K; ++k)
#pragma unroll
      for (int i = 0; i < D; ++i) Ed.m[i][k - 1] = X[(size_t)T[e * K + k] * D + i] - X[(size_t)T[e * K] * D + i];
    double dt = det(Ed);
    Mat<D> c = cofactor(Ed);  // inverse = cof^T / det
    // XHi = inv(Ed); D[j][a] (a>=1) = XHi[a-1][j] = cof[j][a-1] / det
    double inv = 1.0 / dt;
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int a = 0; a < D; ++a) Dm[(size_t)(j * D + a) * t + e] = c.m[j][a] * inv;
    if (D == 3) {
      // tetrahedron_volumes.py:26-27: det of rows (x_k - x_0) / 6  (= det Ed)
      vol[e] = dt / 6.0;
    } else {
      vol[e] = 0.5 * fabs(dt);  // triangle_areas.py: unsigned
    }
  }
};

// D given directly as AoS [e][j][a] (a = 0..dim): used when the plan is rebuilt from J
struct CopyDFunctor {
  const double* Din;
  double* Dm;
  int t, D;
  SKB_HD void operator()(int e) const {
    const int K = D + 1;
    for (int j = 0; j < D; ++j)
      for (int a = 0; a < D; ++a) Dm[(size_t)(j * D + a) * t + e] = Din[((size_t)e * D + j) * K + a + 1];
  }
};

template <class B>
void set_geometry_from_X(PlanData<B>& p, const typename B::template vec<double>& X) {
  auto pol = B::policy();
  const int t = p.t, dim = p.dim;
  thrust::counting_iterator<int> it0(0);
  p.Dm.resize((size_t)dim * dim * t);
  p.vol0.resize(t);
  if (dim == 3) {
    thrust::for_each(pol, it0, it0 + t,
                     GeomFunctor<3>{thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(p.T32.data()),
                                    thrust::raw_pointer_cast(p.Dm.data()), thrust::raw_pointer_cast(p.vol0.data()), t});
  } else {
    thrust::for_each(pol, it0, it0 + t,
                     GeomFunctor<2>{thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(p.T32.data()),
                                    thrust::raw_pointer_cast(p.Dm.data()), thrust::raw_pointer_cast(p.vol0.data()), t});
  }
  p.has_vol0 = true;
}

template <class B>
void set_geometry_from_D(PlanData<B>& p, const typename B::template vec<double>& Daos) {
  auto pol = B::policy();
  const int t = p.t, dim = p.dim;
  thrust::counting_iterator<int> it0(0);
  p.Dm.resize((size_t)dim * dim * t);
  p.vol0.assign(t, 0.0);
  thrust::for_each(pol, it0, it0 + t,
                   CopyDFunctor{thrust::raw_pointer_cast(Daos.data()), thrust::raw_pointer_cast(p.Dm.data()), t, dim});
  p.has_vol0 = false;
}

// topology: pattern, slot map, reduction schedules
template <class B>
void build_plan(PlanData<B>& p, const typename B::template vec<int>& T, int n, int t, int dim, int tile_elems) {
  auto pol = B::policy();
  using IV = typename B::template vec<int>;
  using KV = typename B::template vec<uint64_t>;
  const int K = dim + 1;
  p.dim = dim; p.K = K; p.n = n; p.t = t;
  p.tile_elems = tile_elems;
  p.n_tiles = (t + tile_elems - 1) / tile_elems;
  p.T32 = T;
  thrust::counting_iterator<int> it0(0);

  // ---- block pattern + block schedule
  {
    const int nc = t * K * K;
    KV key(nc);
    IV c_sorted(nc);
    thrust::transform(pol, it0, it0 + nc, key.begin(), BlockKey{thrust::raw_pointer_cast(p.T32.data()), K});
    thrust::sequence(pol, c_sorted.begin(), c_sorted.end());
    thrust::stable_sort_by_key(pol, key.begin(), key.end(), c_sorted.begin());
    IV flag(nc), slot_sorted(nc);
    thrust::transform(pol, it0, it0 + nc, flag.begin(), HeadFlag64{thrust::raw_pointer_cast(key.data())});
    thrust::inclusive_scan(pol, flag.begin(), flag.end(), slot_sorted.begin());
    thrust::transform(pol, slot_sorted.begin(), slot_sorted.end(), slot_sorted.begin(), MinusOne());
    p.nnzb = (int)slot_sorted.back() + 1;
    KV ukey(p.nnzb);
    thrust::unique_copy(pol, key.begin(), key.end(), ukey.begin());
    p.brow.resize(p.nnzb);
    p.bcol.resize(p.nnzb);
    thrust::transform(pol, ukey.begin(), ukey.end(), p.brow.begin(), Hi32());
    thrust::transform(pol, ukey.begin(), ukey.end(), p.bcol.begin(), Lo32());
    p.bptr.resize(n + 1);
    thrust::lower_bound(pol, p.brow.begin(), p.brow.end(), it0, it0 + n + 1, p.bptr.begin());
    p.bslot.resize(nc);
    thrust::scatter(pol, slot_sorted.begin(), slot_sorted.end(), c_sorted.begin(), p.bslot.begin());
    build_sched<B>(p.blocks, p.nnzb, p.n_tiles, K * K, tile_elems, slot_sorted, c_sorted);
  }
  // ---- vertex schedule (gradient scatter)
  {
    const int nc = t * K;
    KV key(nc);
    IV c_sorted(nc);
    thrust::transform(pol, it0, it0 + nc, key.begin(), VertKey{thrust::raw_pointer_cast(p.T32.data())});
    thrust::sequence(pol, c_sorted.begin(), c_sorted.end());
    thrust::stable_sort_by_key(pol, key.begin(), key.end(), c_sorted.begin());
    IV slot_sorted(nc);
    thrust::transform(pol, key.begin(), key.end(), slot_sorted.begin(), Lo32());
    build_sched<B>(p.verts, n, p.n_tiles, K, tile_elems, slot_sorted, c_sorted);
  }
}

}  // namespace skb
