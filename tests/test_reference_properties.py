"""The reference's own test strategy, run against this library: the reference pins this path with
PROPERTIES, not numbers -- rest state stationary, analytic gradient vs central differences of the energy
(``FD_STEP=1e-6``, ``GRAD_TOL=1e-5``), analytic Hessian vs central differences of the gradient (``HESS_TOL=1e-4``,
``psd=False``), ``_u`` tier == ``_x`` tier (``TOL=1e-10``), operator identities, integrator order, FST bilinearity
(tests/test_stable_neo_hookean.py:35-91, test_displacement_u_tier.py:86-123, test_deformation_jacobian.py:27-64,
test_psd_project.py:10-24, test_polar_svd.py:10-24, test_svd_rv.py:10-20, test_integrators.py:30-219,
test_fast_sandwich_transform_clustered.py:33-58).

Every test body runs twice: against the oracle on the CPU (which validates the test itself in the GPU-less build
container) and, marked ``gpu``, against ``simkit_b200`` through its drop-in names.
"""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe
from simkit_b200 import synthetic as syn

MATERIALS = list(oe.MATERIALS)
IMPLS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="simkit_b200", marks=pytest.mark.gpu)]
FD_STEP, GRAD_TOL, HESS_TOL, TIER_TOL = 1e-6, 1e-5, 1e-4, 1e-10


class Oracle:
    name = "oracle"
    deformation_jacobian = staticmethod(oe.deformation_jacobian)
    volume = staticmethod(oe.volume)
    psd_project = staticmethod(oe.psd_project)
    polar_svd = staticmethod(oe.polar_svd)
    svd_rv = staticmethod(oe.svd_rv)
    backward_euler = staticmethod(oe.backward_euler)
    bdf2 = staticmethod(oe.bdf2)

    @staticmethod
    def element(kind, material, F, mu, lam):
        return getattr(oe, kind + "_element_F")(material, F, mu, lam)

    @staticmethod
    def assembled(kind, material, x, J, mu, lam, vol, Jx_bar=None, psd=True):
        if kind == "hessian":
            return oe.hessian_x(material, x, J, mu, lam, vol, psd=psd, Jx_bar=Jx_bar)
        return getattr(oe, kind + "_x")(material, x, J, mu, lam, vol, Jx_bar=Jx_bar)

    @staticmethod
    def fst(A, B, l, dim):
        ARBs = oe.fst_precompute(A, B, l, dim=dim)
        return lambda r: oe.fst_eval(ARBs, r, dim)


class Gpu:
    """``simkit_b200`` through the reference's per-material names (ARAP takes no ``lam``)."""
    name = "gpu"

    def __init__(self):
        import simkit_b200 as sk
        self.sk = sk
        for f in ("deformation_jacobian", "volume", "psd_project", "polar_svd", "svd_rv", "backward_euler", "bdf2"):
            setattr(self, f, getattr(sk, f))

    def element(self, kind, material, F, mu, lam):
        fn = getattr(self.sk, "%s_%s_element_F" % (material, kind))
        return fn(F, mu) if material == "arap" else fn(F, mu, lam)

    def assembled(self, kind, material, x, J, mu, lam, vol, Jx_bar=None, psd=True):
        fn = getattr(self.sk, "%s_%s_%s" % (material, kind, "x" if Jx_bar is None else "u"))
        args = [x, J] + ([] if Jx_bar is None else [Jx_bar]) + ([mu] if material == "arap" else [mu, lam]) + [vol]
        return fn(*args, psd=psd) if kind == "hessian" else fn(*args)

    def fst(self, A, B, l, dim):
        return self.sk.fast_sandwich_transform_clustered(A, B, l, dim=dim)


@pytest.fixture
def impl(request):
    return Oracle() if request.param == "oracle" else Gpu()


def _elements(seed, t, dim):
    rng = np.random.default_rng(seed)
    F = np.eye(dim)[None] + 0.05 * rng.standard_normal((t, dim, dim))
    return F, rng.uniform(0.5, 2.0, (t, 1)), rng.uniform(0.5, 2.0, (t, 1))


def _central(f, x, h=FD_STEP):
    """Central differences of a vector-valued ``f`` at ``x`` (1-D): the Jacobian, one column per coordinate."""
    cols = []
    for k in range(x.size):
        e = np.zeros_like(x)
        e[k] = h
        cols.append((np.asarray(f(x + e), dtype=np.float64).ravel() - np.asarray(f(x - e), dtype=np.float64).ravel()) / (2 * h))
    return np.stack(cols, axis=-1)


# ---------------------------------------------------------------------------------------------- element tier
@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", MATERIALS)
def test_rest_state_is_stationary(impl, material, dim):
    _, mu, lam = _elements(0, 4, dim)
    rest = np.tile(np.eye(dim)[None], (4, 1, 1))
    assert np.abs(impl.element("gradient", material, rest, mu, lam)).max() < 1e-10
    assert np.isfinite(np.asarray(impl.element("energy", material, rest, mu, lam))).all()


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", MATERIALS)
def test_element_gradient_matches_central_differences(impl, material, dim):
    F, mu, lam = _elements(1, 3, dim)
    t = F.shape[0]
    energy = lambda f: np.array([float(np.asarray(impl.element("energy", material, f.reshape(t, dim, dim), mu, lam)).sum())])  # noqa: E731
    g_fd = _central(energy, F.ravel()).reshape(t, dim, dim)
    assert np.allclose(impl.element("gradient", material, F, mu, lam), g_fd, atol=GRAD_TOL)


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", MATERIALS)
def test_element_hessian_matches_central_differences(impl, material, dim):
    F, mu, lam = _elements(2, 2, dim)
    t, b = F.shape[0], dim * dim
    grad = lambda f: impl.element("gradient", material, f.reshape(t, dim, dim), mu, lam)  # noqa: E731
    H_fd = _central(grad, F.ravel())                                   # (t*b, t*b), block diagonal
    H = np.asarray(impl.element("hessian", material, F, mu, lam))
    assert H.shape == (t, b, b)
    assert np.allclose(sps.block_diag(list(H)).toarray(), H_fd, atol=HESS_TOL)


# -------------------------------------------------------------------------------------------- assembled tiers
def _mesh(dim, seed=3):
    cells = (2, 1, 1) if dim == 3 else (2, 2)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(seed)
    U = X + 0.05 * rng.standard_normal(X.shape)
    t = T.shape[0]
    return X, T, U, rng.uniform(0.5, 2.0, (t, 1)), rng.uniform(0.5, 2.0, (t, 1))


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", MATERIALS)
def test_displacement_tier_equals_position_tier(impl, material, dim):
    X, T, U, mu, lam = _mesh(dim)
    J, vol = impl.deformation_jacobian(X, T), impl.volume(X, T)
    u = U - X                                   # (n, dim), as the reference takes it (stable_neo_hookean.py:544-576)
    Jx_bar = np.asarray(J @ X.reshape(-1, 1))
    ex = impl.assembled("energy", material, U, J, mu, lam, vol)
    eu = impl.assembled("energy", material, u, J, mu, lam, vol, Jx_bar=Jx_bar)
    assert isinstance(eu, float) and abs(ex - eu) <= TIER_TOL * max(1.0, abs(ex))
    gx = impl.assembled("gradient", material, U, J, mu, lam, vol)
    gu = impl.assembled("gradient", material, u, J, mu, lam, vol, Jx_bar=Jx_bar)
    assert gu.shape == (X.size, 1) and np.abs(gx - gu).max() <= TIER_TOL * max(1.0, np.abs(gx).max())
    for psd in (False, True):
        Hx = impl.assembled("hessian", material, U, J, mu, lam, vol, psd=psd)
        Hu = impl.assembled("hessian", material, u, J, mu, lam, vol, Jx_bar=Jx_bar, psd=psd)
        assert sps.issparse(Hu) and abs(Hx - Hu).max() <= TIER_TOL * max(1.0, abs(Hx).max())


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("material", MATERIALS)
def test_assembled_derivatives_and_invariances(impl, material, dim):
    """gradient_x vs central differences of energy_x, hessian_x(psd=False) vs central differences of gradient_x;
    translations cost nothing: the gradient sums to zero per coordinate and H maps a translation to zero; the
    projected Hessian has no negative eigenvalue."""
    X, T, U, mu, lam = _mesh(dim)
    J, vol = impl.deformation_jacobian(X, T), impl.volume(X, T)
    n = X.shape[0]
    x0 = U.ravel().copy()
    g = impl.assembled("gradient", material, U, J, mu, lam, vol)
    g_fd = _central(lambda x: np.array([impl.assembled("energy", material, x.reshape(n, dim), J, mu, lam, vol)]), x0)
    assert np.allclose(g.ravel(), g_fd.ravel(), atol=GRAD_TOL)
    H = impl.assembled("hessian", material, U, J, mu, lam, vol, psd=False)
    H_fd = _central(lambda x: impl.assembled("gradient", material, x.reshape(n, dim), J, mu, lam, vol), x0)
    assert np.allclose(H.toarray(), H_fd, atol=HESS_TOL)
    scale = max(1.0, abs(H).max())
    assert np.abs(g.reshape(n, dim).sum(axis=0)).max() < 1e-10 * max(1.0, np.abs(g).max())
    for i in range(dim):
        shift = np.zeros((n, dim))
        shift[:, i] = 1.0
        assert np.abs(H @ shift.reshape(-1, 1)).max() < 1e-10 * scale
    Hp = impl.assembled("hessian", material, U, J, mu, lam, vol, psd=True)
    assert np.abs(Hp.toarray() - Hp.toarray().T).max() < 1e-12 * scale
    if material != "linear_elasticity":       # its own module ignores psd (linear_elasticity.py:199-230)
        assert np.linalg.eigvalsh(Hp.toarray()).min() > -1e-9 * scale


# --------------------------------------------------------------------------------------------------- operators
@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
def test_deformation_jacobian_recovers_linear_maps(impl, dim):
    X, T, _, _, _ = _mesh(dim)
    rng = np.random.default_rng(4)
    A = rng.standard_normal((dim, dim))
    J = impl.deformation_jacobian(X, T)
    assert sps.issparse(J) and J.shape == (T.shape[0] * dim * dim, X.size)
    F = np.asarray(J @ (X @ A.T + rng.standard_normal(dim)).reshape(-1, 1)).reshape(-1, dim, dim)
    assert np.abs(F - A[None]).max() < 1e-10
    vol = np.asarray(impl.volume(X, T))
    assert vol.shape == (T.shape[0], 1) and abs(vol.sum() - 1.0) < 1e-12          # unit square / cube


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("b", [4, 9])
def test_psd_project_floors_the_spectrum(impl, b):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((30, b, b))
    A = A + np.swapaxes(A, 1, 2)
    P = impl.psd_project(A)
    assert P.shape == A.shape and np.linalg.eigvalsh(P).min() >= 1e-6 * (1 - 1e-6)
    w = np.linalg.eigvalsh(A)
    assert np.allclose(np.linalg.eigvalsh(impl.psd_project(A, "abs")), np.sort(np.abs(w), axis=1), atol=1e-10)
    assert impl.psd_project(A[0]).shape == (1, b, b)                               # 2-D input is promoted and stays 3-D
    spd = A[:3] @ np.swapaxes(A[:3], 1, 2) + np.eye(b)[None]
    assert np.allclose(impl.psd_project(spd), spd, rtol=1e-11, atol=1e-11)         # nothing to floor: unchanged


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
def test_polar_and_rotation_variant_svd(impl, dim):
    rng = np.random.default_rng(6)
    F = rng.standard_normal((40, dim, dim))
    F[:10] = np.eye(dim)[None] + 0.3 * rng.standard_normal((10, dim, dim))
    I = np.eye(dim)[None]
    R, S = impl.polar_svd(F)
    assert np.abs(np.swapaxes(R, 1, 2) @ R - I).max() < 1e-10 and np.abs(np.linalg.det(R) - 1.0).max() < 1e-10
    assert np.abs(S - np.swapaxes(S, 1, 2)).max() < 1e-10 and np.abs(R @ S - F).max() < 1e-10
    U, Sg, V = impl.svd_rv(F)
    assert np.abs(U @ Sg @ np.swapaxes(V, 1, 2) - F).max() < 1e-10
    Ruv = U @ np.swapaxes(V, 1, 2)                                     # test_svd_rv.py:10-20: the rotation factor is proper
    assert np.abs(np.swapaxes(Ruv, 1, 2) @ Ruv - I).max() < 1e-10 and np.abs(np.linalg.det(Ruv) - 1.0).max() < 1e-10
    off = Sg.copy()
    off[:, np.arange(dim), np.arange(dim)] = 0.0
    assert np.abs(off).max() < 1e-14                                   # S is a diagonal matrix
    with pytest.raises(NameError):
        impl.polar_svd(F, flip=False)


# ------------------------------------------------------------------------------------------------- integrators
def _oscillator():
    """Three unit masses on a line joined by springs to each other and to a wall: E = 1/2 x^T K x (dense K)."""
    K = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]]) * 4.0
    M = sps.identity(3, format="csc")
    x0 = np.array([[1.0], [0.0], [-0.5]])
    return K, M, x0


def _exact(K, x0, T):
    w2, V = np.linalg.eigh(K)
    return V @ (np.cos(np.sqrt(w2) * T)[:, None] * (V.T @ x0))


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
def test_backward_euler_dissipates_and_is_first_order(impl):
    K, M, x0 = _oscillator()
    E, G, H = (lambda x: 0.5 * float((x.T @ K @ x).item())), (lambda x: K @ x), (lambda x: K)
    Tend, errs = 0.4, []
    for steps in (20, 40):
        h = Tend / steps
        xp, xc = x0.copy(), x0.copy()              # starts at rest
        total = [E(xc)]
        for _ in range(steps):
            xn = impl.backward_euler(xc, xp, E, G, H, M, h, tolerance=1e-12, max_iter=3)
            v = (xn - xc) / h
            total.append(E(xn) + 0.5 * float((v.T @ v).item()))
            xp, xc = xc, xn
        assert all(b <= a + 1e-12 for a, b in zip(total[1:], total[2:]))     # numerical damping: energy never grows
        errs.append(np.abs(xc - _exact(K, x0, Tend)).max())
    assert 1.5 < errs[0] / errs[1] < 2.6                                      # halving h halves the error


@pytest.mark.parametrize("impl", IMPLS, indirect=True)
def test_bdf2_is_second_order(impl):
    K, M, x0 = _oscillator()
    E, G, H = (lambda x: 0.5 * float((x.T @ K @ x).item())), (lambda x: K @ x), (lambda x: K)
    Tend, errs = 0.4, []
    for steps in (20, 40):
        h = Tend / steps
        hist = [_exact(K, x0, -k * h) for k in range(4)]        # x_curr, x_prev, x_prev2, x_prev3 from the exact solution
        for _ in range(steps):
            xn = impl.bdf2(hist[0], hist[1], hist[2], hist[3], E, G, H, M, h, tolerance=1e-12, max_iter=3)
            hist = [xn] + hist[:3]
        errs.append(np.abs(hist[0] - _exact(K, x0, Tend)).max())
    assert 3.0 < errs[0] / errs[1] < 5.2                                      # halving h quarters the error


# --------------------------------------------------------------------------------------------------------- FST
@pytest.mark.parametrize("impl", IMPLS, indirect=True)
@pytest.mark.parametrize("dim", [2, 3])
def test_fast_sandwich_transform_is_linear_in_r(impl, dim):
    rng = np.random.default_rng(7)
    t, m1, m2, nc = 12, 5, 4, 3
    A = rng.standard_normal((m1, dim * dim * t))
    B = sps.random(dim * dim * t, m2, density=0.4, random_state=8, format="csr")
    l = rng.integers(0, nc, size=t)
    l[:nc] = np.arange(nc)
    f = impl.fst(A, B, l, dim)
    r1, r2 = rng.standard_normal((nc, dim, dim)), rng.standard_normal((nc, dim, dim))
    out = f(r1)
    assert out.shape == (m1, m2)
    assert np.abs(f(r1 + 2.0 * r2) - (out + 2.0 * f(r2))).max() < 1e-12 * max(1.0, np.abs(out).max())
