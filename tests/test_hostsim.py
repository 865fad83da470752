"""CPU replay of the kernel phase functions vs the oracle (debug aid for the GPU-less
container; the GPU parity tests are in test_gpu_parity.py)."""
import numpy as np
import pytest

from oracle import elasticity as oe
from simkit_b200 import synthetic as syn
import hostsim

MAT_ID = {m: i for i, m in enumerate(oe.MATERIALS)}
TOL = 1e-10


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _mesh(dim, sigma, seed=0):
    cells = (4, 3, 3) if dim == 3 else (7, 5)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(X.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    X, T = X[perm], inv[T]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=sigma, seed=seed)
    mu, lam = syn.heterogeneous_lame(T.shape[0], seed=seed + 1)
    return X, T, U, mu, lam


@pytest.mark.parametrize("dim", [2, 3])
def test_svd_matches_convention(dim):
    rng = np.random.default_rng(3)
    F = rng.standard_normal((200, dim, dim))
    F[:20] = np.eye(dim) + 1e-9 * rng.standard_normal((20, dim, dim))   # near-degenerate
    F[20] = np.eye(dim)
    F[21] = 0.0
    F[22, :, -1] = 0.0                                                  # rank deficient
    U, S, V = hostsim.svd(F)
    Uo, So, Vo = oe.svd_rv(F)
    Sd = np.zeros_like(F)
    Sd[:, np.arange(dim), np.arange(dim)] = S
    assert rel(U @ Sd @ np.swapaxes(V, 1, 2), F) < 1e-13
    assert np.allclose(np.linalg.det(U), 1.0, atol=1e-13) and np.allclose(np.linalg.det(V), 1.0, atol=1e-13)
    assert np.abs(U @ np.swapaxes(U, 1, 2) - np.eye(dim)).max() < 1e-14
    assert rel(S, So[:, np.arange(dim), np.arange(dim)]) < 1e-13
    # polar factors are unique where F is well conditioned
    good = np.abs(So[:, np.arange(dim), np.arange(dim)]).min(axis=1) > 1e-6
    R = U @ np.swapaxes(V, 1, 2)
    Ro, _ = oe.polar_svd(F)
    assert rel(R[good], Ro[good]) < 1e-9


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", oe.MATERIALS)
def test_element_tier(dim, material):
    X, T, U, mu, lam = _mesh(dim, 0.4 if material != "neo_hookean" else 0.1)
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    F = np.asarray(J @ U.reshape(-1, 1)).reshape(-1, dim, dim)
    psi, P = hostsim.element_energy_gradient(MAT_ID[material], F, mu, lam)
    assert rel(psi, oe.energy_element_F(material, F, mu, lam).ravel()) < 1e-13
    assert rel(P, oe.gradient_element_F(material, F, mu, lam)) < 1e-12
    H = hostsim.element_hessian(MAT_ID[material], 0, F, mu, lam, np.ones(F.shape[0]))
    assert rel(H, oe.hessian_element_F(material, F, mu, lam)) < TOL
    for mode, before in ((1, False), (2, True)):
        Hp = hostsim.element_hessian(MAT_ID[material], mode, F, mu, lam, vol)
        if material == "linear_elasticity" and mode == 1:
            continue
        ref = oe.weighted_element_hessians(material, F, mu, lam, vol, psd=True, psd_before_vol=before)
        assert rel(Hp, ref) < TOL


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", oe.MATERIALS)
@pytest.mark.parametrize("sigma", [0.1, 0.4])
def test_assembly(dim, material, sigma):
    if material == "neo_hookean" and sigma > 0.2:
        pytest.skip("NaN by construction for inverted elements")
    X, T, U, mu, lam = _mesh(dim, sigma)
    n = X.shape[0]
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    psd_mode = 0 if material == "linear_elasticity" else 1
    out = hostsim.run(X, T, MAT_ID[material], psd_mode, U, mu, lam, vol, tile_elems=32)
    indptr, indices, bptr, bcol = oe.structural_pattern(T, n, dim)
    assert np.array_equal(out["bptr"], bptr) and np.array_equal(out["bcol"], bcol)
    assert rel(out["vol0"], vol.ravel()) < 1e-14
    E = oe.energy_x(material, U, J, mu, lam, vol)
    assert abs(out["energy"] - E) <= 1e-12 * abs(E)
    g = oe.gradient_x(material, U, J, mu, lam, vol)
    assert rel(out["g"], g.ravel()) < TOL
    Q = oe.hessian_x(material, U, J, mu, lam, vol, psd=True)
    Qours = hostsim.csr_from_blocks(out["bptr"], out["bcol"], out["vals"], n, dim)
    assert rel(Qours.toarray(), Q.toarray()) < TOL
    # the assembled matrix is exactly symmetric (upper blocks are mirrored, not recomputed)
    assert abs(Qours - Qours.T).max() == 0.0


@pytest.mark.parametrize("dim", [2, 3])
def test_pencil_order_merges_more_blocks_per_tile(dim):
    """``synthetic.pencil_order`` (bench.py --element-order pencil): a permutation of the same mesh that leaves fewer
    partial records per element at 128-element tiles and the same assembled Hessian and gradient."""
    cells = (18, 18, 18) if dim == 3 else (48, 48)
    ext = tuple(1.0 for _ in cells)
    X, T = syn.make_mesh(cells)
    order = syn.pencil_order(cells, ext, X, T)
    assert np.array_equal(np.sort(order), np.arange(T.shape[0]))
    U = syn.jittered_state(X, cells, ext, sigma=0.1)
    mu, lam = syn.lame()
    a = hostsim.run(X, T, 0, 1, U, mu, lam, tile_elems=128)
    b = hostsim.run(X, T[order], 0, 1, U, mu, lam, tile_elems=128)
    assert b["info"][1] < 0.85 * a["info"][1] and b["info"][2] < 0.85 * a["info"][2]
    assert np.array_equal(a["bptr"], b["bptr"]) and np.array_equal(a["bcol"], b["bcol"])
    assert rel(b["vals"], a["vals"]) < 1e-13 and rel(b["g"], a["g"]) < 1e-13


def _odd_meshes(dim):
    """Edge-case inputs: vertices no element references (at the end and in the middle of the numbering), one element,
    two disjoint elements, an element listed twice, elements shuffled with rotated corners."""
    rng = np.random.default_rng(5)
    cells = (3, 2, 2) if dim == 3 else (5, 4)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    out = {"plain": (X, T, U)}
    out["unreferenced_end"] = (np.vstack([X, rng.random((3, dim))]), T, np.vstack([U, rng.random((3, dim))]))
    ins = np.insert(np.arange(X.shape[0]), [2, 2, 7], -1)
    newid = np.full(X.shape[0], -1)
    newid[ins[ins >= 0]] = np.where(ins >= 0)[0]
    Xg, Ug = rng.random((ins.size, dim)), rng.random((ins.size, dim))
    Xg[newid], Ug[newid] = X, U
    out["unreferenced_middle"] = (Xg, newid[T], Ug)
    out["one_element"] = (X, T[:1], U)
    out["two_disjoint"] = (X, T[[0, -1]], U)
    out["listed_twice"] = (X, np.vstack([T, T[:2]]), U)
    rot = [1, 2, 0, 3] if dim == 3 else [1, 2, 0]            # even permutation: orientation kept
    out["shuffled"] = (X, T[rng.permutation(T.shape[0])][:, rot], U)
    return out


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("case", ["unreferenced_end", "unreferenced_middle", "one_element", "two_disjoint", "listed_twice",
                                  "shuffled"])
def test_edge_case_meshes(case, dim):
    X, T, U = _odd_meshes(dim)[case]
    n = X.shape[0]
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    g = oe.gradient_x("stable_neo_hookean", U, J, mu, lam, vol).ravel()
    H = oe.hessian_x("stable_neo_hookean", U, J, mu, lam, vol, psd=True)
    for tile in (32, 128):
        r = hostsim.run(X, T, MAT_ID["stable_neo_hookean"], 1, U, mu, lam, vol=vol, tile_elems=tile)
        Hh = hostsim.csr_from_blocks(r["bptr"], r["bcol"], r["vals"], n, dim)
        assert rel(r["g"], g) < TOL and abs(Hh - H).max() <= TOL * abs(H).max()
        assert abs(r["energy"] - oe.energy_x("stable_neo_hookean", U, J, mu, lam, vol)) <= 1e-12 * abs(r["energy"])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shuffle", [False, True])
def test_plan_element_order(dim, shuffle):
    """The plan's own spatial element order (plan.cuh ``str_element_order``, what ``skb_plan_create`` applies): fewer
    partial records per element than the generator's order -- also when the caller lists the elements in random order
    -- with the same pattern and the same assembled values; heterogeneous per-element materials, volumes and the
    ``_u`` tier's ``Jx_bar`` stay in the caller's order at the boundary."""
    cells = (18, 18, 18) if dim == 3 else (48, 48)
    ext = tuple(1.0 for _ in cells)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(11)
    if shuffle:
        T = T[rng.permutation(T.shape[0])]
    U = syn.jittered_state(X, cells, ext, sigma=0.1)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    vol = oe.volume(X, T)
    Fbar = 0.01 * rng.standard_normal((T.shape[0], dim, dim))
    a = hostsim.run(X, T, 0, 1, U, mu, lam, vol, Fbar=Fbar, tile_elems=128)
    b = hostsim.run(X, T, 0, 1, U, mu, lam, vol, Fbar=Fbar, tile_elems=128, reorder=True)
    lexi = hostsim.run(*syn.make_mesh(cells), 0, 1, U, *syn.lame(), tile_elems=128)["info"]
    assert b["info"][1] < 0.85 * lexi[1] and b["info"][2] < 0.85 * lexi[2]
    if shuffle:
        assert b["info"][1] < 0.5 * a["info"][1]
    assert np.array_equal(a["bptr"], b["bptr"]) and np.array_equal(a["bcol"], b["bcol"])
    assert rel(b["vals"], a["vals"]) < 1e-13 and rel(b["g"], a["g"]) < 1e-13
    assert abs(b["energy"] - a["energy"]) <= 1e-13 * abs(a["energy"])
    assert np.array_equal(a["Dm"], b["Dm"]) and np.array_equal(a["vol0"], b["vol0"])
