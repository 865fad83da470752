"""CPU checks of the boundary: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), and the Python surface mirrors the reference's names."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "simkit_b200.h")
LIB = os.path.join(ROOT, "simkit_b200", "libsimkit_b200.so")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z_0-9A-Z]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from simkit_b200 import build
        build.build()
    return ctypes.CDLL(LIB)


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from simkit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_no_gpu_fails_loudly(lib):
    """Without a device the product raises instead of falling back to any CPU path."""
    import simkit_b200 as sk
    from simkit_b200._lib import SimkitB200Error
    lib.skb_device_count.restype = ctypes.c_int
    if lib.skb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(SimkitB200Error):
        sk.psd_project(np.eye(3))
    with pytest.raises(SimkitB200Error):
        sk.deformation_jacobian(np.eye(3)[:, :2] * 1.0, np.array([[0, 1, 2]]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert not re.search(r"^\s*(import|from)\s+simkit(\s|\.)", txt, flags=re.M), f
                assert "sys.path" not in txt, f


def test_python_surface_mirrors_reference_names():
    import simkit_b200 as sk
    import inspect
    mats = ["stable_neo_hookean", "neo_hookean", "stvk", "linear_elasticity", "arap"]
    for m in mats:
        for kind in ("energy", "gradient", "hessian"):
            for tier in ("_element_F", "_x", "_u", ""):
                assert hasattr(sk.energies, f"{m}_{kind}{tier}"), f"{m}_{kind}{tier}"
    # positional signatures of the reference
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(sk.stable_neo_hookean_hessian_x) == ["X", "J", "mu", "lam", "vol", "psd"]
    assert sig(sk.stable_neo_hookean_hessian_u) == ["u", "J", "Jx_bar", "mu", "lam", "vol", "psd"]
    assert sig(sk.arap_hessian_x) == ["X", "J", "mu", "vol", "psd"]
    assert sig(sk.stable_neo_hookean_hessian) == ["X", "T", "mu", "lam", "U", "psd"]
    assert sig(sk.newton_solver)[:8] == ["x0", "energy_func", "gradient_func", "hessian_func", "tolerance",
                                         "max_iter", "do_line_search", "return_info"]
    assert sig(sk.backtracking_line_search) == ["f", "x0", "g", "dx", "alpha", "beta", "max_iter", "threshold"]
    assert sig(sk.backward_euler)[:11] == ["x_curr", "x_prev", "energy_func", "gradient_func", "hessian_func", "M",
                                           "h", "tolerance", "max_iter", "do_line_search", "return_info"]
    assert sig(sk.bdf2)[:4] == ["x_curr", "x_prev", "x_prev2", "x_prev3"]
    for name in ("deformation_jacobian", "volume", "massmatrix", "psd_project", "polar_svd",
                 "fast_sandwich_transform_clustered", "elastic_hessian_z", "ElasticEnergyZPrecomp"):
        assert hasattr(sk, name), name


def test_host_logic_line_search_and_newton_on_quadratic():
    """Host control flow (no GPU): Armijo on a quadratic, reference test_backtracking_line_search.py:15-52."""
    import simkit_b200 as sk
    A = np.diag([1.0, 4.0, 9.0])
    b = np.array([[1.0], [2.0], [3.0]])
    f = lambda x: float(0.5 * x.T @ A @ x - b.T @ x)
    x0 = np.zeros((3, 1))
    g = A @ x0 - b
    dx = np.linalg.solve(A, -g)
    t, x, fx = sk.backtracking_line_search(f, x0, g, dx)
    assert t == 1.0 and np.allclose(x, dx) and fx == f(dx)
    t, x, fx = sk.backtracking_line_search(f, x0, g, -dx, max_iter=5)
    assert t == 0.0 and x is x0
    with pytest.raises(AssertionError):
        sk.backtracking_line_search(f, x0, g, dx, alpha=0.9)


def test_native_nccl_plumbing():
    """capi_nccl.cu takes NCCL with dlopen(RTLD_NOLOAD) from the process (torch's copy): the argument block has the
    layout the C side asserts, and an id can be made without a GPU once torch is imported."""
    import ctypes
    import numpy as np
    import torch  # noqa: F401  (loads libnccl.so.2)
    from simkit_b200 import _lib
    assert ctypes.sizeof(_lib.DistPcgArgs) == 152
    lib = _lib.load(require_gpu=False)
    idbuf = np.zeros(128, dtype=np.uint8)
    rc = lib.skb_nccl_unique_id(_lib.ptr(idbuf), idbuf.size)
    if rc != 0:   # a torch build with NCCL linked statically: the native path reports it instead of loading another NCCL
        assert b"libnccl.so.2" in lib.skb_last_error()
        return
    assert idbuf.any()
    assert lib.skb_nccl_unique_id(_lib.ptr(idbuf), 64) != 0          # buffer too small
    assert lib.skb_nccl_set_halo(None, 0, None, None, None, None, None, None, None) != 0   # no plan / no communicator


def test_potential_uploads_materials_once_per_change():
    """the device-resident step re-uploaded mu / lam / vol (390 MB at 16 M tets) on every call.  The upload is
    now tagged with an owner token on the plan; anything else that writes the plan's materials resets the token."""
    import numpy as np
    from simkit_b200.potential import ElasticPotential

    class FakePlan:
        ndof, n, dim, t, COARSE_MIN_ITERS = 6, 2, 3, 1, 10 ** 9
        uploads = 0

        def volume(self):
            return np.ones(1)

        def set_materials(self, mu, lam, vol, owner=None):
            self.uploads += 1
            self._mat_owner = owner

        def set_contact_plane(self, *a):
            pass

        def set_contact_sphere(self, *a):
            pass

        def newton(self, *a, **k):
            return np.zeros(6), dict(pcg_iters=0, iters=0)

    plan = FakePlan()
    pot = ElasticPotential("stable_neo_hookean", 1.0, 2.0, plan=plan)
    other = ElasticPotential("arap", 3.0, 0.0, plan=plan)
    pot.newton(np.zeros(6))
    pot.newton(np.zeros(6))
    assert plan.uploads == 1
    other.newton(np.zeros(6))          # another potential on the same plan takes the device copies over
    pot.newton(np.zeros(6))
    assert plan.uploads == 3
    plan._mat_owner = None             # what MeshPlan._mats does on any direct call that passes materials
    pot.newton(np.zeros(6))
    assert plan.uploads == 4
    pot.update_materials(mu=5.0)
    pot.newton(np.zeros(6))
    pot.newton(np.zeros(6))
    assert plan.uploads == 5 and pot.mu == 5.0
