// Per-mesh plan: everything that depends only on (X, T), built once.
//
// HBM layout (all arrays live on the device; SoA so that a warp's loads coalesce):
//   T32   [t][K]        int32   element corners, K = dim+1, sorted ascending per element  (16 B/tet)
//   Dm    [dim*dim][t]  f64     D[j][a], a = 1..dim (corner 0 = -sum)    (72 B/tet)
//   vol0  [t]           f64     rest quadrature weights
//   bptr  [n+1], bcol[nnzb], brow[nnzb]  canonical block pattern (vertex adjacency, sorted)
//   upos  [nu]          4xint32 where upper block (v<=w) and its transpose sit in the CSR values
// plus two deterministic reduction schedules (upper blocks and vertices), see ReduceSched.
// The stiffness is symmetric, so only the blocks with row vertex <= column vertex are reduced;
// the finalize kernel writes each sum to (v,w) and its transpose to (w,v).  Corners are sorted per
// element so that a local corner pair a <= b is always such an upper block (no runtime transposes).
//
// The builder is written against thrust with a backend tag so that the exact
// same code runs with thrust::device in the product and thrust::host inside the
// CPU test harness (tests/host_harness.cu) of the GPU-less build container.
#pragma once
#include <stdint.h>
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <thrust/extrema.h>
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
#include <thrust/transform_reduce.h>
#include <thrust/functional.h>
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
// Contributions c (one per (element, corner) for the gradient, one per (element, corner pair a<=b)
// for the stiffness) are grouped by (slot, tile) where tile = element / tile_elems.  Each group is
// a *tile-slot*; level 1 (inside the assembly kernel, from shared memory) sums a tile-slot's
// contributions in ascending c and writes ONE partial record; level 2 (the finalize kernel) sums
// the partial records of a slot in ascending tile order.  Partial records are numbered slot-major,
// so level 2 reads a contiguous run.  No atomics anywhere: the result is bitwise reproducible.
//
// Layout, chosen so that a tile's whole schedule is two contiguous, 8/16-byte aligned ranges the
// kernel prefetches into shared memory with cp.async while phase 1 computes:
//   tl_ptr  [n_tiles+1]        tile -> its entries
//   tl_ent  [n_ts] (uint2)     .x = partial record index q, .y = cbeg | cend << 16  (offsets into
//                              the tile's tc_src range, contributions of the entry are contiguous)
//   tc_src  [n_tiles * tile_elems * per_elem] (uint16)   packed (which << 8 | local element),
//                              tile-major at a FIXED stride per tile (tile_elems <= 256)
// Inside a tile the entries are ordered by DESCENDING contribution count, so the lanes of a warp
// in phase 2 run loops of nearly equal length.
//   sp_ptr  [n_slots+1]        slot -> contiguous range of partial records
struct SchedEntry {
  unsigned q;
  unsigned range;  // cbeg | cend << 16
};

struct ReduceSchedView {
  int n_slots;
  int n_ts;  // number of tile-slots == number of partial records
  int n_contrib;
  int per_elem;
  int max_entries;  // max over tiles of the entry count (shared-memory sizing)
  const int* tl_ptr;
  const SchedEntry* tl_ent;
  const uint16_t* tc_src;
  const int* sp_ptr;
};

template <class B>
struct ReduceSched {
  int n_slots = 0, n_ts = 0, n_contrib = 0, per_elem = 0, max_entries = 0;
  typename B::template vec<int> tl_ptr, sp_ptr;
  typename B::template vec<SchedEntry> tl_ent;
  typename B::template vec<uint16_t> tc_src;
  ReduceSchedView view() const {
    ReduceSchedView v;
    v.n_slots = n_slots;
    v.n_ts = n_ts;
    v.n_contrib = n_contrib;
    v.per_elem = per_elem;
    v.max_entries = max_entries;
    v.tl_ptr = thrust::raw_pointer_cast(tl_ptr.data());
    v.tl_ent = thrust::raw_pointer_cast(tl_ent.data());
    v.tc_src = thrust::raw_pointer_cast(tc_src.data());
    v.sp_ptr = thrust::raw_pointer_cast(sp_ptr.data());
    return v;
  }
};

// where an upper block (v <= w) and its transpose live in the canonical scalar CSR values
struct UpperPos {
  int base, stride;    // block (v, w): row i of the block starts at base + i * stride
  int tbase, tstride;  // block (w, v) (== base/stride when v == w)
};

struct PlanView {
  int dim, K;
  int n, t;
  int tile_elems, n_tiles;
  int nnzb;  // block non-zeros of the full (both triangles) pattern
  int nu;    // upper block slots (v <= w)
  const int* T32;   // [t][K] corners sorted ascending by vertex id (internal order)
  const double* Dm;
  const double* vol0;
  const int* bptr;
  const int* bcol;
  const int* brow;
  const UpperPos* upos;  // [nu]
  const int* eorder;     // [t] internal element -> caller's element, or nullptr (identity); see str_element_order
  ReduceSchedView blocks;
  ReduceSchedView verts;
};

template <class B>
struct PlanData {
  int dim = 0, K = 0, n = 0, t = 0, tile_elems = 0, n_tiles = 0, nnzb = 0, nu = 0;
  int t_energy = 0;  // elements [0, t_energy) count in energies and other sums over elements (== t unless a shard also
                     // evaluates its lower neighbour's interface elements, which then follow its own)
  bool has_vol0 = false;
  typename B::template vec<int> T32, bptr, bcol, brow;
  typename B::template vec<int> eorder;    // [t] internal element i is the caller's element eorder[i]; empty = identity
  typename B::template vec<uint8_t> perm;  // [t][K]: internal corner s is the caller's corner perm[s]
  typename B::template vec<UpperPos> upos;
  typename B::template vec<double> Dm, vol0;
  ReduceSched<B> blocks, verts;
  PlanView view() const {
    PlanView v;
    v.dim = dim; v.K = K; v.n = n; v.t = t;
    v.tile_elems = tile_elems; v.n_tiles = n_tiles; v.nnzb = nnzb; v.nu = nu;
    v.T32 = thrust::raw_pointer_cast(T32.data());
    v.Dm = thrust::raw_pointer_cast(Dm.data());
    v.vol0 = thrust::raw_pointer_cast(vol0.data());
    v.bptr = thrust::raw_pointer_cast(bptr.data());
    v.bcol = thrust::raw_pointer_cast(bcol.data());
    v.brow = thrust::raw_pointer_cast(brow.data());
    v.upos = thrust::raw_pointer_cast(upos.data());
    v.eorder = eorder.empty() ? nullptr : thrust::raw_pointer_cast(eorder.data());
    v.blocks = blocks.view();
    v.verts = verts.view();
    return v;
  }
};

// pair index p in [0, K(K+1)/2) -> local corners (a <= b), row-major over the upper triangle
SKB_HD void pair_corners(int K, int p, int& a, int& b) {
  a = 0;
  int rowlen = K;
  while (p >= rowlen) {
    p -= rowlen;
    ++a;
    --rowlen;
  }
  b = a + p;
}

// ------------------------------------------------------------------ functors
struct HeadFlag64 {
  const uint64_t* k;
  SKB_HD int operator()(int i) const { return (i == 0 || k[i] != k[i - 1]) ? 1 : 0; }
};

// sorts the corners of element e ascending by vertex id (insertion sort, K <= 4); flags repeats
struct SortCorners {
  const int* Tin;
  int* Tout;
  uint8_t* perm;
  int* bad;
  int K;
  SKB_HD void operator()(int e) const {
    int v[4];
    uint8_t pm[4];
    for (int a = 0; a < K; ++a) {
      v[a] = Tin[e * K + a];
      pm[a] = (uint8_t)a;
    }
    for (int a = 1; a < K; ++a) {
      int x = v[a];
      uint8_t px = pm[a];
      int b = a - 1;
      while (b >= 0 && v[b] > x) {
        v[b + 1] = v[b];
        pm[b + 1] = pm[b];
        --b;
      }
      v[b + 1] = x;
      pm[b + 1] = px;
    }
    bool rep = false;
    for (int a = 0; a < K; ++a) {
      Tout[e * K + a] = v[a];
      perm[e * K + a] = pm[a];
      if (a > 0 && v[a] == v[a - 1]) rep = true;
    }
    if (rep) bad[0] = 1;  // benign race: every writer stores the same value
  }
};

// key of an upper block contribution c = e*NP + p : (row vertex, col vertex), row <= col
struct UpperKey {
  const int* T;  // sorted corners
  int K, NP;
  SKB_HD uint64_t operator()(int c) const {
    int e = c / NP;
    int a, b;
    pair_corners(K, c - e * NP, a, b);
    return ((uint64_t)(uint32_t)T[e * K + a] << 32) | (uint32_t)T[e * K + b];
  }
};
// key of a vertex contribution c = e*K + a : vertex
struct VertKey {
  const int* T;
  SKB_HD uint64_t operator()(int c) const { return (uint64_t)(uint32_t)T[c]; }
};
struct SwapKey {
  SKB_HD uint64_t operator()(uint64_t k) const { return (k << 32) | (k >> 32); }
};
struct IsOffDiag {
  SKB_HD bool operator()(uint64_t k) const { return (uint32_t)(k >> 32) != (uint32_t)(k & 0xffffffffu); }
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
// writes the packed source of the i-th contribution (tile-major order) at its fixed-stride position
#if defined(SKB_EXP_SRCBASE)
template <int D> SKB_HD int pair_base(int pp);       // kernels.cuh (staging layout)
template <int D> SKB_HD bool pair_is_diag(int pp);
#endif
struct ScatterSrc {
  const int* c;  // contribution ids, tile-major
  uint16_t* tc_src;
  int per_elem;
  int tile_elems;
  int tile_stride;  // tile_elems * per_elem
  SKB_HD void operator()(int i) const {
    const int e = c[i] / per_elem;
    const int rem = c[i] - e * per_elem;
    const int tile = e / tile_elems;
    // every tile before `tile` is full, so position i is tile * tile_stride + local offset
#if defined(SKB_EXP_SRCBASE)
    // A/B experiment: block schedules carry the staging offset of the pair (7 bits) and its diagonal flag (bit 15)
    // instead of the pair index, so phase 2 neither recomputes pair_base nor tests pair_is_diag per contribution
    int hi = rem;
    if (per_elem == 10) hi = pair_base<3>(rem) | (pair_is_diag<3>(rem) ? 0x80 : 0);
    else if (per_elem == 6) hi = pair_base<2>(rem) | (pair_is_diag<2>(rem) ? 0x80 : 0);
    tc_src[i] = (uint16_t)((hi << 8) | (e - tile * tile_elems));
#else
    tc_src[i] = (uint16_t)((rem << 8) | (e - tile * tile_elems));
#endif
    (void)tile_stride;
  }
};
// sort key of entry j inside its tile: descending contribution count
struct EntryOrderKey {
  const uint64_t* head_key;
  const int* head_pos;
  int n_ts, nc;
  SKB_HD uint64_t operator()(int j) const {
    const int cnt = ((j + 1 < n_ts) ? head_pos[j + 1] : nc) - head_pos[j];
    return (head_key[j] & 0xffffffff00000000ull) | (uint32_t)(0x7fffffff - cnt);
  }
};
struct MakeEntry {
  const int* order;          // sorted position -> entry j
  const uint64_t* head_key;  // (tile, q)
  const int* head_pos;       // position of the entry's first contribution (tile-major order)
  const int* dense_ptr;      // tile -> first entry (dense numbering)
  const int* pad_ptr;        // tile -> first entry (padded numbering)
  int n_ts, nc, tile_stride;
  SchedEntry* out;
  SKB_HD void operator()(int pos) const {
    const int j = order[pos];
    const int tile = (int)(head_key[j] >> 32);
    const int beg = head_pos[j] - tile * tile_stride;
    // every tile before the last is full, so a next head in the next tile sits exactly at the tile's end
    const int end = ((j + 1 < n_ts) ? head_pos[j + 1] : nc) - tile * tile_stride;
    SchedEntry en;
    en.q = (unsigned)(head_key[j] & 0xffffffffu);
    en.range = (unsigned)beg | ((unsigned)end << 16);
    out[pad_ptr[tile] + (pos - dense_ptr[tile])] = en;
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
struct PaddedCount {
  const int* p;
  SKB_HD int operator()(int i) const { return (p[i + 1] - p[i] + 1) & ~1; }
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
  s.per_elem = per_elem;
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
  KV head_key(s.n_ts);
  thrust::gather(pol, head_pos.begin(), head_pos.end(), key2.begin(), head_key.begin());
  const int tile_stride = tile_elems * per_elem;
  // dense tile -> entry offsets, then pad every tile's entry list to an even count so that each
  // list starts 16-byte aligned and is a multiple of 16 bytes long (TMA bulk copy granularity)
  IV head_tile(s.n_ts);
  thrust::transform(pol, head_key.begin(), head_key.end(), head_tile.begin(), Hi32());
  IV dense_ptr(n_tiles + 1);
  thrust::lower_bound(pol, head_tile.begin(), head_tile.end(), it0, it0 + n_tiles + 1, dense_ptr.begin());
  s.tl_ptr.assign(n_tiles + 1, 0);
  {
    IV cnt(n_tiles);
    thrust::transform(pol, it0, it0 + n_tiles, cnt.begin(), PaddedCount{thrust::raw_pointer_cast(dense_ptr.data())});
    s.max_entries = n_tiles ? (int)*thrust::max_element(pol, cnt.begin(), cnt.end()) : 0;
    thrust::inclusive_scan(pol, cnt.begin(), cnt.end(), s.tl_ptr.begin() + 1);
  }
  SchedEntry pad;
  pad.q = 0xffffffffu;
  pad.range = 0;
  s.tl_ent.assign((size_t)(int)s.tl_ptr[n_tiles], pad);
  IV order(s.n_ts);
  {
    KV okey(s.n_ts);
    thrust::transform(pol, it0, it0 + s.n_ts, okey.begin(),
                      EntryOrderKey{thrust::raw_pointer_cast(head_key.data()), thrust::raw_pointer_cast(head_pos.data()),
                                    s.n_ts, nc});
    thrust::sequence(pol, order.begin(), order.end());
    thrust::stable_sort_by_key(pol, okey.begin(), okey.end(), order.begin());
  }
  thrust::for_each(pol, it0, it0 + s.n_ts,
                   MakeEntry{thrust::raw_pointer_cast(order.data()), thrust::raw_pointer_cast(head_key.data()),
                             thrust::raw_pointer_cast(head_pos.data()),
                             thrust::raw_pointer_cast(dense_ptr.data()), thrust::raw_pointer_cast(s.tl_ptr.data()),
                             s.n_ts, nc, tile_stride, thrust::raw_pointer_cast(s.tl_ent.data())});
  // packed sources at a fixed stride per tile (padded to whole tiles so 16-byte prefetches never
  // run past the allocation)
  s.tc_src.assign((size_t)n_tiles * tile_stride, (uint16_t)0);
  thrust::for_each(pol, it0, it0 + nc,
                   ScatterSrc{thrust::raw_pointer_cast(c2.data()), thrust::raw_pointer_cast(s.tc_src.data()),
                              per_elem, tile_elems, tile_stride});
}

// ------------------------------------------------------- spatial element order --
// The assembly kernel reduces a tile of `tile_elems` CONSECUTIVE elements on chip, so the number of partial
// records that leave the SM (and the locality of the x gathers) depends on how compact a tile is in space.  The
// caller's element order is arbitrary (north_star: "spatially sorted"), so the plan lists the active elements in
// a sort-tile-recursive order of their own: key = lower corner of the element's bounding box (the 6 tets of a
// Kuhn cell, or the fan of elements hanging off one vertex, share it and stay together), sorted along x and cut
// into slabs, each slab sorted along y and cut into pencils, each pencil sorted along z; slab and pencil sizes
// are whole multiples of the tile, chosen so that a tile is roughly a cube.  On the lexicographic 139^3 Kuhn grid
// this gives 2.33 partial records per tet instead of 3.12, and the same on a randomly shuffled element list
// (host replay, tests/test_hostsim.py).  The order is internal: every per-element array crossing the C ABI is in
// the caller's order and is permuted at the boundary (eorder).
template <int D>
struct CornerKeyFunctor {
  const double* X;
  const int* T;  // caller's corners (any order)
  double lo[3], scale[3];
  uint32_t* q;   // [D][t] quantised lower bbox corner, 14 bits per axis
  int t;
  SKB_HD void operator()(int e) const {
    constexpr int K = D + 1;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double m = X[(size_t)T[e * K] * D + i];
#pragma unroll
      for (int a = 1; a < K; ++a) {
        const double v = X[(size_t)T[e * K + a] * D + i];
        m = v < m ? v : m;
      }
      double f = (m - lo[i]) * scale[i];
      f = f < 0.0 ? 0.0 : (f > 16383.0 ? 16383.0 : f);
      q[(size_t)i * t + e] = (uint32_t)f;
    }
  }
};
// key of the element at sorted position `pos`: segment (pos / seg_size) then the three axes in the order a0, a1, a2
struct StrKey {
  const int* order;
  const uint32_t* q;
  int t, D, seg_size, a0, a1, a2;
  SKB_HD uint64_t operator()(int pos) const {
    const int e = order[pos];
    uint64_t k = (uint64_t)(seg_size > 0 ? pos / seg_size : 0) << 42;
    k |= (uint64_t)q[(size_t)a0 * t + e] << 28;
    k |= (uint64_t)q[(size_t)a1 * t + e] << 14;
    if (D == 3) k |= (uint64_t)q[(size_t)a2 * t + e];
    return k;
  }
};
struct MinMaxCoord {
  const double* X;
  int D, axis;
  SKB_HD double operator()(int v) const { return X[(size_t)v * D + axis]; }
};
struct PermuteRows {
  const int* Tin;
  const int* order;
  int* Tout;
  int K;
  SKB_HD void operator()(int i) const {
    for (int a = 0; a < K; ++a) Tout[i * K + a] = Tin[order[i] * K + a];
  }
};

// order[i] = caller's index of the i-th element in the internal order, for the first t (active) elements of T
template <class B>
void str_element_order(const typename B::template vec<double>& X, const typename B::template vec<int>& T, int n, int t,
                       int dim, int tile_elems, typename B::template vec<int>& order) {
  auto pol = B::policy();
  using KV = typename B::template vec<uint64_t>;
  thrust::counting_iterator<int> it0(0);
  const int K = dim + 1;
  double lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
  const double* Xp = thrust::raw_pointer_cast(X.data());
  for (int a = 0; a < dim; ++a) {
    const double mn = thrust::transform_reduce(pol, it0, it0 + n, MinMaxCoord{Xp, dim, a}, 1e300, thrust::minimum<double>());
    const double mx = thrust::transform_reduce(pol, it0, it0 + n, MinMaxCoord{Xp, dim, a}, -1e300, thrust::maximum<double>());
    lo[a] = mn;
    ext[a] = (mx > mn) ? (mx - mn) : 1.0;
  }
  typename B::template vec<uint32_t> q((size_t)dim * t);
  if (dim == 3) {
    CornerKeyFunctor<3> f{Xp, thrust::raw_pointer_cast(T.data()), {lo[0], lo[1], lo[2]},
                          {16383.999 / ext[0], 16383.999 / ext[1], 16383.999 / ext[2]}, thrust::raw_pointer_cast(q.data()), t};
    thrust::for_each(pol, it0, it0 + t, f);
  } else {
    CornerKeyFunctor<2> f{Xp, thrust::raw_pointer_cast(T.data()), {lo[0], lo[1], 0.0},
                          {16383.999 / ext[0], 16383.999 / ext[1], 1.0}, thrust::raw_pointer_cast(q.data()), t};
    thrust::for_each(pol, it0, it0 + t, f);
  }
  // tile edge in units of the mean element spacing; slab / pencil sizes in elements, whole tiles
  double vol = 1.0;
  for (int a = 0; a < dim; ++a) vol *= ext[a];
  const double ell = pow(vol / (double)t, 1.0 / dim);
  const double s = pow((double)tile_elems, 1.0 / dim) * ell;
  long long pencil, slab;
  if (dim == 3) {
    pencil = (long long)floor((double)t * s * s / (ext[0] * ext[1]) / tile_elems + 0.5);
    if (pencil < 1) pencil = 1;
    pencil *= tile_elems;
    long long per = (long long)floor(ext[1] / s + 0.5);
    if (per < 1) per = 1;
    slab = per * pencil;
  } else {
    pencil = 0;
    slab = (long long)floor((double)t * s / ext[0] / tile_elems + 0.5);
    if (slab < 1) slab = 1;
    slab *= tile_elems;
  }
  if (slab > t) slab = t;
  if (pencil > t) pencil = t;
  order.resize(t);
  thrust::sequence(pol, order.begin(), order.end());
  KV key(t);
  const uint32_t* qp = thrust::raw_pointer_cast(q.data());
  auto pass = [&](int seg, int a0, int a1, int a2) {
    thrust::transform(pol, it0, it0 + t, key.begin(), StrKey{thrust::raw_pointer_cast(order.data()), qp, t, dim, seg, a0, a1, a2});
    thrust::stable_sort_by_key(pol, key.begin(), key.end(), order.begin());
  };
  if (dim == 3) {
    pass(0, 0, 1, 2);
    pass((int)slab, 1, 2, 0);
    pass((int)pencil, 2, 1, 0);
  } else {
    pass(0, 0, 1, 1);
    pass((int)slab, 1, 0, 0);
  }
}

struct AddOffset {
  int off;
  SKB_HD int operator()(int x) const { return x + off; }
};

// Reorders the first t (active) rows of T into the internal order and records it in p.eorder.  t_split in (0, t)
// orders the ranges [0, t_split) and [t_split, t) separately (a shard that also evaluates its lower neighbour's
// interface elements keeps its own elements first: the energy sums over those only).
template <class B>
void apply_element_order(PlanData<B>& p, typename B::template vec<int>& T, const typename B::template vec<double>& X, int n,
                         int t, int dim, int tile_elems, int t_split = 0) {
  auto pol = B::policy();
  const int K = dim + 1;
  using IV = typename B::template vec<int>;
  p.eorder.resize(t);
  int bounds[3] = {0, (t_split > 0 && t_split < t) ? t_split : t, t};
  for (int part = 0; part < 2; ++part) {
    const int a = bounds[part], b = bounds[part + 1];
    if (b <= a) continue;
    IV Tsub(T.begin() + (size_t)a * K, T.begin() + (size_t)b * K);
    IV ord;
    str_element_order<B>(X, Tsub, n, b - a, dim, tile_elems, ord);
    thrust::transform(pol, ord.begin(), ord.end(), p.eorder.begin() + a, AddOffset{a});
  }
  IV Tn(T);  // pattern-only rows (>= t) stay where they are
  thrust::counting_iterator<int> it0(0);
  thrust::for_each(pol, it0, it0 + t,
                   PermuteRows{thrust::raw_pointer_cast(T.data()), thrust::raw_pointer_cast(p.eorder.data()),
                               thrust::raw_pointer_cast(Tn.data()), K});
  T.swap(Tn);
}

// ---------------------------------------------------------------- geometry --
// D = (H (X_e^T H)^-1)^T in the CALLER's corner order (its corner-0 column is minus the sum of the
// others), stored for the internal corners 1..dim (internal corner 0 is again implied), and the
// rest quadrature weight (signed by the caller's orientation).
template <int D>
struct GeomFunctor {
  const double* X;
  const int* T;        // internal (sorted) corners
  const uint8_t* perm; // internal corner s = caller corner perm[s]
  double* Dm;
  double* vol;
  int t;
  SKB_HD void operator()(int e) const {
    constexpr int K = D + 1;
    int To[K];  // caller order
#pragma unroll
    for (int s = 0; s < K; ++s) To[perm[e * K + s]] = T[e * K + s];
    Mat<D> Ed;  // edge matrix: column k-1 = x_k - x_0
#pragma unroll
    for (int k = 1; k < K; ++k)
#pragma unroll
      for (int i = 0; i < D; ++i) Ed.m[i][k - 1] = X[(size_t)To[k] * D + i] - X[(size_t)To[0] * D + i];
    double dt = det(Ed);
    Mat<D> c = cofactor(Ed);  // inverse = cof^T / det
    // D[j][a] (a>=1) = cof[j][a-1] / det ; D[j][0] = -sum
    double inv = 1.0 / dt;
    double Df[D][K];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double s0 = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        Df[j][a + 1] = c.m[j][a] * inv;
        s0 -= Df[j][a + 1];
      }
      Df[j][0] = s0;
    }
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int s = 1; s < K; ++s) {
        const int a = perm[e * K + s];
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q)
          if (q == a) v = Df[j][q];
        Dm[(size_t)(j * D + (s - 1)) * t + e] = v;
      }
    if (D == 3) {
      // tetrahedron_volumes.py:26-27: det of rows (x_k - x_0) / 6  (= det Ed)
      vol[e] = dt / 6.0;
    } else {
      vol[e] = 0.5 * fabs(dt);  // triangle_areas.py: unsigned
    }
  }
};

// D given directly as AoS [e][j][a] (a = 0..dim, caller order): used when the plan is rebuilt from J
struct CopyDFunctor {
  const double* Din;
  const uint8_t* perm;
  double* Dm;
  int t, D;
  SKB_HD void operator()(int e) const {
    const int K = D + 1;
    for (int j = 0; j < D; ++j)
      for (int s = 1; s < K; ++s)
        Dm[(size_t)(j * D + (s - 1)) * t + e] = Din[((size_t)e * D + j) * K + perm[e * K + s]];
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
                                    thrust::raw_pointer_cast(p.perm.data()), thrust::raw_pointer_cast(p.Dm.data()),
                                    thrust::raw_pointer_cast(p.vol0.data()), t});
  } else {
    thrust::for_each(pol, it0, it0 + t,
                     GeomFunctor<2>{thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(p.T32.data()),
                                    thrust::raw_pointer_cast(p.perm.data()), thrust::raw_pointer_cast(p.Dm.data()),
                                    thrust::raw_pointer_cast(p.vol0.data()), t});
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
                   CopyDFunctor{thrust::raw_pointer_cast(Daos.data()), thrust::raw_pointer_cast(p.perm.data()),
                                thrust::raw_pointer_cast(p.Dm.data()), t, dim});
  p.has_vol0 = false;
}

struct MakeUpperPos {
  const uint64_t* ukey;   // upper slot keys (v, w), v <= w
  const uint64_t* fkey;   // full pattern keys, sorted
  const int* bptr;
  int nnzb, D;
  UpperPos* out;
  SKB_HD int find(uint64_t k) const {
    int lo = 0, hi = nnzb;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (fkey[mid] < k) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  SKB_HD void operator()(int u) const {
    const uint64_t k = ukey[u];
    const int v = (int)(k >> 32), w = (int)(k & 0xffffffffu);
    const int s = find(k), st = find(((uint64_t)(uint32_t)w << 32) | (uint32_t)v);
    UpperPos up;
    int b0 = bptr[v], nb = bptr[v + 1] - b0;
    up.base = b0 * D * D + (s - b0) * D;
    up.stride = nb * D;
    b0 = bptr[w];
    nb = bptr[w + 1] - b0;
    up.tbase = b0 * D * D + (st - b0) * D;
    up.tstride = nb * D;
    out[u] = up;
  }
};

struct LessThan {
  int bound;
  SKB_HD bool operator()(int c) const { return c < bound; }
};

// topology: pattern, upper-slot positions, reduction schedules.  Returns false if an element
// repeats a vertex.
// T holds t_total >= t elements: the first t are ACTIVE (evaluated, reduced); the remaining
// "pattern-only" elements (another rank's elements that touch vertices this rank owns, see
// simkit_b200/sharding.py) only reserve their slots in the CSR pattern so that the neighbour's
// interface contributions have somewhere to land.
template <class B>
bool build_plan(PlanData<B>& p, const typename B::template vec<int>& T, int n, int t, int dim, int tile_elems,
                int t_total = -1) {
  if (t_total < t) t_total = t;
  auto pol = B::policy();
  using IV = typename B::template vec<int>;
  using KV = typename B::template vec<uint64_t>;
  const int K = dim + 1;
  const int NP = K * (K + 1) / 2;
  p.dim = dim; p.K = K; p.n = n; p.t = t;
  if (p.t_energy <= 0 || p.t_energy > t) p.t_energy = t;
  p.tile_elems = tile_elems;
  p.n_tiles = (t + tile_elems - 1) / tile_elems;
  thrust::counting_iterator<int> it0(0);

  // ---- internal corner order: ascending vertex id, so that local a <= b implies row <= col
  p.T32.resize((size_t)t_total * K);
  p.perm.resize((size_t)t_total * K);
  {
    IV bad(1, 0);
    thrust::for_each(pol, it0, it0 + t_total,
                     SortCorners{thrust::raw_pointer_cast(T.data()), thrust::raw_pointer_cast(p.T32.data()),
                                 thrust::raw_pointer_cast(p.perm.data()), thrust::raw_pointer_cast(bad.data()), K});
    if ((int)bad[0] != 0) return false;
  }

  // ---- upper block slots, full pattern, block schedule
  {
    const int nc = t_total * NP;
    KV key(nc);
    IV c_sorted(nc);
    thrust::transform(pol, it0, it0 + nc, key.begin(), UpperKey{thrust::raw_pointer_cast(p.T32.data()), K, NP});
    thrust::sequence(pol, c_sorted.begin(), c_sorted.end());
    thrust::stable_sort_by_key(pol, key.begin(), key.end(), c_sorted.begin());
    IV flag(nc), slot_sorted(nc);
    thrust::transform(pol, it0, it0 + nc, flag.begin(), HeadFlag64{thrust::raw_pointer_cast(key.data())});
    thrust::inclusive_scan(pol, flag.begin(), flag.end(), slot_sorted.begin());
    thrust::transform(pol, slot_sorted.begin(), slot_sorted.end(), slot_sorted.begin(), MinusOne());
    p.nu = (int)slot_sorted.back() + 1;
    KV ukey(p.nu);
    thrust::unique_copy(pol, key.begin(), key.end(), ukey.begin());
    // full pattern = upper slots + transposes of the off-diagonal ones
    KV fkey(2 * (size_t)p.nu);
    thrust::copy(pol, ukey.begin(), ukey.end(), fkey.begin());
    auto mid = fkey.begin() + p.nu;
    auto last = thrust::copy_if(pol, ukey.begin(), ukey.end(), mid, IsOffDiag());
    thrust::transform(pol, mid, last, mid, SwapKey());
    fkey.resize(last - fkey.begin());
    thrust::sort(pol, fkey.begin(), fkey.end());
    p.nnzb = (int)fkey.size();
    p.brow.resize(p.nnzb);
    p.bcol.resize(p.nnzb);
    thrust::transform(pol, fkey.begin(), fkey.end(), p.brow.begin(), Hi32());
    thrust::transform(pol, fkey.begin(), fkey.end(), p.bcol.begin(), Lo32());
    p.bptr.resize(n + 1);
    thrust::lower_bound(pol, p.brow.begin(), p.brow.end(), it0, it0 + n + 1, p.bptr.begin());
    p.upos.resize(p.nu);
    thrust::for_each(pol, it0, it0 + p.nu,
                     MakeUpperPos{thrust::raw_pointer_cast(ukey.data()), thrust::raw_pointer_cast(fkey.data()),
                                  thrust::raw_pointer_cast(p.bptr.data()), p.nnzb, dim,
                                  thrust::raw_pointer_cast(p.upos.data())});
    if (t_total > t) {
      // the schedule covers active elements only (stable filter keeps ascending c inside a slot)
      const int nact = t * NP;
      IV s2(nact), c2(nact);
      thrust::copy_if(pol, slot_sorted.begin(), slot_sorted.end(), c_sorted.begin(), s2.begin(), LessThan{nact});
      thrust::copy_if(pol, c_sorted.begin(), c_sorted.end(), c2.begin(), LessThan{nact});
      build_sched<B>(p.blocks, p.nu, p.n_tiles, NP, tile_elems, s2, c2);
    } else {
      build_sched<B>(p.blocks, p.nu, p.n_tiles, NP, tile_elems, slot_sorted, c_sorted);
    }
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
  return true;
}

}  // namespace skb
