"""The lazy objects behind the reference's call pattern: ``deformation_jacobian`` returns a
``csc_matrix`` whose host arrays are built on first touch, the ``*_hessian_x`` functions a ``csr_matrix`` whose values
stay in HBM through ``H + S`` sums and the solve.  Every path must give what the plain host matrices give."""
import numpy as np
import pytest
import scipy.sparse as sps

import simkit_b200 as sk
from oracle import elasticity as oe
from simkit_b200 import synthetic as syn
from simkit_b200.device_csr import DeviceCSR

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _setup(cells=(6, 5, 4)):
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.2)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    return X, T, U, mu, lam


def test_lazy_jacobian_is_the_reference_matrix_when_touched():
    X, T, U, mu, lam = _setup()
    J = sk.deformation_jacobian(X, T)
    assert sps.issparse(J) and sps.isspmatrix_csc(J) and J.shape == (T.shape[0] * 9, X.shape[0] * 3)
    assert J._T is not None                                    # nothing built yet
    vol = sk.volume(X, T)
    sk.stable_neo_hookean_energy_x(U, J, mu, lam, vol)          # the energy tiers only read the attached plan
    assert J._T is not None
    Jo = oe.canonical_csr(oe.deformation_jacobian(X, T))
    F = (J @ U.reshape(-1, 1)).reshape(-1, 3, 3)                # first touch builds the host arrays
    assert J._T is None
    assert rel(F, (Jo @ U.reshape(-1, 1)).reshape(-1, 3, 3)) < 1e-13
    # same matrix; the stored patterns differ only where the oracle's SpGEMM leaves a rounding-size value that is an
    # exact (pruned) zero here -- on jittered meshes the patterns are identical (test_gpu_parity.py::test_golden_operators)
    assert abs(sps.csr_matrix(J) - Jo).max() <= 1e-12 * abs(Jo).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_lazy_hessian_sums_and_solves_on_the_device(dim):
    cells = (6, 5, 4) if dim == 3 else (14, 11)
    X, T, U, mu, lam = _setup(cells)
    n, nd = X.shape[0], X.size
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    H = sk.stable_neo_hookean_hessian_x(U, J, mu, lam, vol)
    assert isinstance(H, DeviceCSR) and isinstance(H, sps.csr_matrix) and sps.issparse(H) and H.on_device
    assert H.shape == (nd, nd) and H.dtype == np.float64 and H.format == "csr"
    Ho = oe.canonical_csr(oe.hessian_x("stable_neo_hookean", U, Jo, mu, lam, volo))
    assert H.on_device                                          # shape / dtype / nnz / format did not download
    # sums with the caller's matrices stay on the device: diagonal (csc, dia), block diagonal (contact-like), lazy + lazy
    rng = np.random.default_rng(3)
    mdiag = 1.0 + rng.random(nd)
    M = sps.diags(mdiag).tocsc()
    h = 1e-2
    A1 = H + M * (1.0 / h ** 2)
    A2 = H + sps.diags(mdiag)                                   # dia format
    blocks = [np.outer(v, v) for v in rng.standard_normal((n, dim))]
    C = sps.block_diag(blocks, format="csr")
    A3 = H + C
    A4 = C + H                                                  # csr on the left: Python tries DeviceCSR.__radd__ first
    A5 = 2.0 * H - H
    A6 = (H + H.copy()) / 2.0
    for A in (A1, A2, A3, A4, A5, A6):
        assert isinstance(A, DeviceCSR) and A.on_device
    assert H.on_device
    tol = 1e-10
    assert abs(sps.csr_matrix(A1.copy()) - (Ho + M / h ** 2)).max() / abs(Ho).max() < tol
    assert abs(sps.csr_matrix(A2.copy()) - (Ho + sps.diags(mdiag))).max() / abs(Ho).max() < tol
    assert abs(sps.csr_matrix(A3.copy()) - (Ho + C)).max() / abs(Ho).max() < tol
    assert abs(sps.csr_matrix(A4.copy()) - (Ho + C)).max() / abs(Ho).max() < tol
    assert abs(sps.csr_matrix(A5.copy()) - Ho).max() / abs(Ho).max() < tol
    assert abs(sps.csr_matrix(A6.copy()) - Ho).max() / abs(Ho).max() < tol
    # a term with an entry outside the mesh's pattern: scipy's host path, same sum
    far = sps.csr_matrix(([3.0], ([0], [nd - 1])), shape=(nd, nd))
    assert Ho[0, nd - 1] == 0.0
    A7 = H.copy() + far
    assert abs(A7 - (Ho + far)).max() / abs(Ho).max() < tol
    # the solve on resident values equals the solve on the host copy and the direct solve
    import scipy.sparse.linalg as spla
    rhs = rng.standard_normal(nd)
    A1h = sps.csr_matrix(A1.copy())
    x_dev = sk.solve_sparse(A1, rhs, rtol=1e-12)
    assert A1.on_device
    x_host = sk.solve_sparse(A1h, rhs, rtol=1e-12)
    x_dir = spla.spsolve((Ho + M / h ** 2).tocsc(), rhs)
    assert rel(x_dev, x_dir) < 1e-9 and rel(x_host, x_dir) < 1e-9 and rel(x_dev, x_host) < 1e-10
    # touching the data makes it an ordinary host matrix, once
    d = H.data
    assert not H.on_device and d.size == H.nnz and np.array_equal(H.indptr, sk.MeshPlan(X=X, T=T).csr_pattern()[0])
    assert abs(H - Ho).max() / abs(Ho).max() < tol
    assert rel(H @ rhs, Ho @ rhs) < 1e-10


def test_reference_style_closures_keep_the_hessian_on_the_device(monkeypatch):
    """energy / gradient / Hessian closures around the `*_x` functions + gravity + a plane contact term handed to
    backward_euler (examples/interactive_demos/010_interactive_contact_plane_3D.py:81-116): same iterate as the oracle's
    loop with SuperLU, and no Hessian value crosses PCIe on the way."""
    import simkit_b200.device_csr as dc
    cells = (6, 5, 4)
    X, T = syn.make_mesh(cells)
    dim = 3
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    M = sps.kron(oe.massmatrix(X, T, rho), sps.identity(dim)).tocsc()
    Mn = oe.massmatrix(X, T, rho)
    fg = oe.gravity_force(X, T, -9.8, rho).reshape(-1, 1)
    k, p0, nrm = 1e5, np.array([0.0, 0.15, 0.0]), np.array([0.0, 1.0, 0.0])
    m = "stable_neo_hookean"

    def closures(mod, Jm, volm, contact):
        e_x, g_x, h_x = (getattr(mod, f"{m}_{q}_x") if mod is sk else (lambda *a, q=q: getattr(oe, f"{q}_x")(m, *a))
                         for q in ("energy", "gradient", "hessian"))
        ce, cg, ch = contact

        def E(x):
            xn = x.reshape(-1, dim)
            return e_x(xn, Jm, mu, lam, volm) + ce(xn, k, p0, nrm, Mn) - float((fg.T @ x).item())

        def G(x):
            xn = x.reshape(-1, dim)
            return g_x(xn, Jm, mu, lam, volm) + cg(xn, k, p0, nrm, Mn) - fg

        def H(x):
            xn = x.reshape(-1, dim)
            return h_x(xn, Jm, mu, lam, volm) + ch(xn, k, p0, nrm, Mn)

        return E, G, H

    def oc(kind):
        return lambda xn, *a: oe.contact_springs_plane(xn, *a)[kind]

    Eo, Go, Ho = closures(oe, Jo, volo, (oc(0), oc(1), oc(2)))
    Es, Gs, Hs = closures(sk, J, vol, (sk.contact_springs_plane_energy, sk.contact_springs_plane_gradient,
                                       sk.contact_springs_plane_hessian))
    rng = np.random.default_rng(4)
    x_curr = X.reshape(-1, 1) + 1e-3 * rng.standard_normal((X.size, 1))
    x_prev = X.reshape(-1, 1) + 1e-3 * rng.standard_normal((X.size, 1))
    xo, io = oe.backward_euler(x_curr, x_prev, Eo, Go, Ho, M, h, max_iter=3, return_info=True)
    downloads = []
    orig = dc.DeviceCSR._materialize

    def spy(self):
        if self.on_device:
            downloads.append(1)
        return orig(self)

    monkeypatch.setattr(dc.DeviceCSR, "_materialize", spy)
    xs, is_ = sk.backward_euler(x_curr, x_prev, Es, Gs, Hs, M, h, max_iter=3, return_info=True, pcg_rtol=1e-13)
    assert downloads == []                                       # nothing touched the values on the host
    assert list(is_["alphas"]) == list(io["alphas"]) and is_["iters"] == io["iters"]
    assert rel(xs, xo) < 1e-8
