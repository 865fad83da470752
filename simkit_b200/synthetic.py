"""Synthetic meshes, materials and states for the BASELINE configs.

These generators are host-side numpy utilities shared by ``bench.py`` and the
tests. They have no reference counterpart (the reference's own generators,
``examples/interactive_demos/utils.py:33-90``, import polyscope and split cells
into 5 tets); the meshes here are the ones BASELINE.json's configs name:

* 3D: ``mx*my*mz`` cells, vertex id ``(i*(my+1)+j)*(mz+1)+k``, each cell split
  into 6 tets along the main diagonal (Kuhn / Freudenthal), all with positive
  signed volume.
* 2D: ``m*m`` cells, two triangles ``(v00,v10,v11),(v00,v11,v01)`` per cell.
"""

from itertools import permutations

import numpy as np

# name -> (dim, cells, extent)   
CONFIGS = {
    "C1": dict(dim=3, cells=(20, 20, 20), extent=(1.0, 1.0, 1.0)),
    "C2": dict(dim=2, cells=(316, 316), extent=(1.0, 1.0)),
    "C3": dict(dim=3, cells=(32, 32, 163), extent=(1.0, 1.0, 163.0 / 32.0)),
    "C4": dict(dim=3, cells=(88, 88, 88), extent=(1.0, 1.0, 1.0)),
    "C5": dict(dim=3, cells=(139, 139, 139), extent=(1.0, 1.0, 1.0)),
}


def kuhn_tet_grid(mx, my, mz, extent=(1.0, 1.0, 1.0), dtype_index=np.int64):
    """Kuhn 6-tet-per-cell grid.  Returns ``X (n,3) f64``, ``T (6*mx*my*mz,4)``."""
    nx, ny, nz = mx + 1, my + 1, mz + 1
    gi, gj, gk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    X = np.stack(
        [gi * (extent[0] / mx), gj * (extent[1] / my), gk * (extent[2] / mz)], axis=-1
    ).reshape(-1, 3).astype(np.float64)
    return X, kuhn_tets(mx, my, mz, 0, mx, dtype_index)


def kuhn_tets(mx, my, mz, i0, i1, dtype_index=np.int64):
    """Elements of the cell planes ``i0 <= i < i1`` of the Kuhn grid (global vertex ids, global element order)."""
    ny, nz = my + 1, mz + 1
    ci, cj, ck = np.meshgrid(np.arange(i0, i1), np.arange(my), np.arange(mz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()

    def vid(i, j, k):
        return (i * ny + j) * nz + k

    tets = []
    for perm in permutations(range(3)):
        # walk 000 -> 111 adding one unit step per axis in the order `perm`
        off = np.zeros(3, dtype=np.int64)
        corners = [vid(ci, cj, ck)]
        for ax in perm:
            off = off.copy()
            off[ax] = 1
            corners.append(vid(ci + off[0], cj + off[1], ck + off[2]))
        tet = np.stack(corners, axis=-1)
        # orientation of this permutation class is constant over the grid
        sign = np.linalg.det(np.eye(3)[list(perm)])
        if sign < 0:
            tet = tet[:, [0, 2, 1, 3]]
        tets.append(tet)
    # cell-major ordering: the 6 tets of a cell are contiguous
    return np.stack(tets, axis=1).reshape(-1, 4).astype(dtype_index)


def tri_grid(mx, my, extent=(1.0, 1.0), dtype_index=np.int64):
    """Two-triangle-per-cell grid.  Returns ``X (n,2) f64``, ``T (2*mx*my,3)``."""
    nx, ny = mx + 1, my + 1
    gi, gj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    X = np.stack([gi * (extent[0] / mx), gj * (extent[1] / my)], axis=-1).reshape(-1, 2)
    X = X.astype(np.float64)
    return X, grid_tris(mx, my, 0, mx, dtype_index)


def grid_tris(mx, my, i0, i1, dtype_index=np.int64):
    """Triangles of the cell columns ``i0 <= i < i1`` (global vertex ids, global element order)."""
    ny = my + 1
    ci, cj = np.meshgrid(np.arange(i0, i1), np.arange(my), indexing="ij")
    ci, cj = ci.ravel(), cj.ravel()
    v00 = ci * ny + cj
    v10 = (ci + 1) * ny + cj
    v11 = (ci + 1) * ny + cj + 1
    v01 = ci * ny + cj + 1
    return np.stack(
        [np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], axis=1
    ).reshape(-1, 3).astype(dtype_index)


def grid_elements(cells, i0, i1):
    """Elements of the cell planes ``[i0, i1)`` along the slowest axis of a synthetic grid."""
    cells = tuple(cells)
    return kuhn_tets(*cells, i0, i1) if len(cells) == 3 else grid_tris(*cells, i0, i1)


def grid_vertices(cells, extent, ids):
    """Rest positions of the vertices ``ids`` (global ids) of a synthetic grid, without building the whole grid."""
    cells = tuple(cells)
    ids = np.asarray(ids, dtype=np.int64)
    dims = [c + 1 for c in cells]
    idx = np.unravel_index(ids, dims)
    return np.stack([idx[a] * (extent[a] / cells[a]) for a in range(len(cells))], axis=-1).astype(np.float64)


def jittered_state_rows(cells, extent, ids, sigma=0.1, seed=0):
    """Rows ``ids`` of ``jittered_state`` of the full grid (same random stream), for one rank of a sharded run."""
    cells = tuple(cells)
    n = int(np.prod([c + 1 for c in cells]))
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((n, len(cells)))[np.asarray(ids, dtype=np.int64)]
    return grid_vertices(cells, extent, ids) + sigma * cell_size(cells, extent) * noise


def make_mesh(name_or_cells, extent=None):
    """Mesh for a named config (``"C1"``..``"C5"``) or an explicit cell tuple."""
    if isinstance(name_or_cells, str):
        cfg = CONFIGS[name_or_cells]
        cells, extent = cfg["cells"], cfg["extent"]
    else:
        cells = tuple(name_or_cells)
        if extent is None:
            extent = tuple(1.0 for _ in cells)
    if len(cells) == 3:
        return kuhn_tet_grid(*cells, extent=extent)
    return tri_grid(*cells, extent=extent)


def pencil_order(cells, extent, X, T, width=None):
    """Permutation of the elements that makes consecutive runs of 128 elements compact in space: cells are ordered in
    pencils of ``width x width`` cells (3 x 3 in 3D, 8 in 2D) running along the last axis, so a tile of the assembly kernel
    covers about 3 x 3 x 2.4 cells (8 x 8 in 2D) instead of a 21-cell line and merges more of its Hessian blocks
    before they leave the SM (2.33 partial records per tet instead of 3.12; `tests/test_hostsim.py`).  The mesh is
    the same mesh: only the order in which its elements are listed changes."""
    dim = len(cells)
    if width is None:
        width = 3 if dim == 3 else 8
    h = np.array([e / c for e, c in zip(extent, cells)])
    cen = np.asarray(X)[np.asarray(T)].mean(axis=1)
    ci = np.minimum(np.floor(cen / h[None, :]).astype(np.int64), np.array(cells)[None, :] - 1)
    if dim == 3:
        keys = (ci[:, 1] % width, ci[:, 0] % width, ci[:, 2], ci[:, 1] // width, ci[:, 0] // width)
    else:
        keys = (ci[:, 0] % width, ci[:, 1], ci[:, 0] // width)
    return np.lexsort(keys)


def cell_size(cells, extent):
    return min(e / c for e, c in zip(extent, cells))


def jittered_state(X, cells, extent, sigma=0.1, seed=0):
    """``U = X + sigma*(cell size)*N(0,1)`` with ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    return X + sigma * cell_size(cells, extent) * rng.standard_normal(X.shape)


def lame(ym=1e5, pr=0.45):
    """(mu, lam) from Young's modulus / Poisson ratio (reference ``ympr_to_lame.py:32-33``)."""
    mu = ym / (2 * (1 + pr))
    lam = ym * pr / ((1 + pr) * (1 - 2 * pr))
    return mu, lam


def heterogeneous_lame(t, ym=1e5, pr=0.45, seed=1):
    """Per-element ``ym*10**U(-1,1)`` (SURVEY §8d heterogeneous variant)."""
    rng = np.random.default_rng(seed)
    yme = ym * 10.0 ** rng.uniform(-1.0, 1.0, size=(t, 1))
    return lame(yme, pr)


def cos_modes(X, r, seed=2, lo=None, hi=None, n_total=None):
    """The same family of smooth modes as ``smooth_modes`` WITHOUT the QR, as a pure function of the vertex position:
    rows can be generated for any subset of the vertices (one rank of a sharded mesh passes the global bounding box
    ``lo, hi`` and vertex count ``n_total``) and agree with the rows of the global basis.  Columns are scaled by
    ``1/sqrt(n_total)`` (about unit norm)."""
    rng = np.random.default_rng(seed)
    n, dim = X.shape
    lo = X.min(0) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = X.max(0) if hi is None else np.asarray(hi, dtype=np.float64)
    n_total = n if n_total is None else int(n_total)
    Xn = (X - lo) / (hi - lo)
    B = np.empty((n * dim, r))
    for j in range(r):
        k = rng.integers(0, 4, size=dim)
        phase = rng.uniform(0, np.pi, size=dim)
        f = np.prod(np.cos(np.pi * k[None, :] * Xn + phase[None, :]), axis=1)
        w = rng.standard_normal(dim)
        B[:, j] = (f[:, None] * w[None, :]).reshape(-1) / np.sqrt(n_total)
    return B


def smooth_modes(X, r, seed=2):
    """Dense random smooth basis ``B (n*dim, r)`` for the reduced config (C4).

    Low-frequency cosine products with random wave vectors applied per
    coordinate, then orthonormalised (QR).  Stands in for the reference's
    skinning eigenmodes, whose eigen-solver needs cvxopt (absent).
    """
    rng = np.random.default_rng(seed)
    n, dim = X.shape
    Xn = (X - X.min(0)) / (X.max(0) - X.min(0))
    B = np.empty((n * dim, r))
    for j in range(r):
        k = rng.integers(0, 4, size=dim)
        phase = rng.uniform(0, np.pi, size=dim)
        f = np.prod(np.cos(np.pi * k[None, :] * Xn + phase[None, :]), axis=1)
        w = rng.standard_normal(dim)
        B[:, j] = (f[:, None] * w[None, :]).reshape(-1)
    Q, _ = np.linalg.qr(B)
    return np.ascontiguousarray(Q)
