"""TEST INFRASTRUCTURE ONLY: builds and wraps tests/host_harness.cu (CPU replay of the
kernel phase functions).  Never imported by ``simkit_b200``."""

import ctypes
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
# SKB_HOSTSIM_FLAGS="-DSKB_EXP_..." replays an A/B kernel variant (same switches as SKB_BUILD_FLAGS in build.py)
_FLAGS = os.environ.get("SKB_HOSTSIM_FLAGS", "").split()
_SUFFIX = "".join(c if c.isalnum() else "_" for c in "".join(_FLAGS))
LIB = os.path.join(BUILD, "libskb_hostsim%s.so" % ("_" + _SUFFIX if _SUFFIX else ""))
SRC = os.path.join(HERE, "host_harness.cu")
CSRC = os.path.join(os.path.dirname(HERE), "simkit_b200", "csrc")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)


def _stale():
    if not os.path.exists(LIB):
        return True
    tl = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > tl for d in deps)


def load():
    if _stale():
        os.makedirs(BUILD, exist_ok=True)
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcudafe", "--diag_suppress=177"] + _FLAGS + [
               "-o", LIB, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host harness build failed:\n" + r.stderr)
    return ctypes.CDLL(LIB)


def _p(a, ty=_dp):
    return None if a is None else a.ctypes.data_as(ty)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def run(X, T, material, psd_mode, x, mu, lam, vol=None, Fbar=None, tile_elems=32, want_g=True, want_h=True,
        t_active=None, reorder=False):
    lib = load()
    X = _c(X)
    T = np.ascontiguousarray(T, dtype=np.int64)
    n, dim = X.shape
    t = T.shape[0]
    t_active = t if t_active is None else int(t_active)
    K = dim + 1
    x = _c(x)
    mu = _c(np.asarray(mu, dtype=np.float64).reshape(-1))
    lam = _c(np.asarray(lam, dtype=np.float64).reshape(-1))
    vol = _c(np.asarray(vol, dtype=np.float64).reshape(-1)) if vol is not None else None
    Fbar = _c(Fbar)
    info = np.zeros(3, dtype=np.int64)
    args = [_p(X), _p(T, _lp), n, t, t_active, dim, tile_elems, material, psd_mode, _p(x), _p(Fbar), _p(mu), mu.size,
            _p(lam), lam.size, _p(vol), 0 if vol is None else vol.size, _p(info, _lp)]
    lib.hs_run.argtypes = [_dp, _lp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                           ctypes.c_int, _dp, _dp, _dp, ctypes.c_int64, _dp, ctypes.c_int64, _dp, ctypes.c_int64,
                           _lp, _ip, _ip, _ip, _dp, _dp, _dp, _dp, _dp, ctypes.c_int]
    lib.hs_run(*args, None, None, None, None, None, None, None, None, int(bool(reorder)))
    nnzb = int(info[0])
    bptr = np.zeros(n + 1, dtype=np.int32)
    bcol = np.zeros(nnzb, dtype=np.int32)
    bslot = np.zeros(t * K * K, dtype=np.int32)
    Dm = np.zeros(dim * dim * t_active)
    vol0 = np.zeros(t_active)
    g = np.zeros(n * dim) if want_g else None
    vals = np.zeros(nnzb * dim * dim) if want_h else None
    energy = np.zeros(1)
    lib.hs_run(*args, _p(bptr, _ip), _p(bcol, _ip), _p(bslot, _ip), _p(Dm), _p(vol0), _p(g), _p(vals), _p(energy),
               int(bool(reorder)))
    return dict(bptr=bptr, bcol=bcol, bslot=bslot.reshape(t, K, K), Dm=Dm.reshape(dim, dim, t_active), vol0=vol0,
                g=g, vals=vals, energy=float(energy[0]), info=info)


def csr_from_blocks(bptr, bcol, vals, n, dim):
    """scipy CSR from the canonical scalar layout (rows of a block row are contiguous runs)."""
    import scipy.sparse as sps
    nb = np.diff(bptr)
    indptr = np.zeros(n * dim + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.repeat(nb * dim, dim))
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for v in range(n):
        cols = (bcol[bptr[v]:bptr[v + 1]][:, None] * dim + np.arange(dim)[None, :]).ravel()
        for i in range(dim):
            r = v * dim + i
            indices[indptr[r]:indptr[r + 1]] = cols
    return sps.csr_matrix((vals, indices, indptr), shape=(n * dim, n * dim))


def value_positions(bptr, bcol, rows, cols, dim):
    """csr_value_position of kernels.cuh for scalar entries (rows[k], cols[k]); -1 = outside the pattern."""
    lib = load()
    bptr = np.ascontiguousarray(bptr, dtype=np.int32)
    bcol = np.ascontiguousarray(bcol, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    out = np.empty(rows.size, dtype=np.int32)
    lib.hs_value_positions.argtypes = [ctypes.c_int, _ip, _ip, ctypes.c_int64, _ip, _ip, _ip]
    lib.hs_value_positions(dim, _p(bptr, _ip), _p(bcol, _ip), rows.size, _p(rows, _ip), _p(cols, _ip), _p(out, _ip))
    return out


def svd(F):
    lib = load()
    F = _c(F)
    t, dim, _ = F.shape
    U = np.zeros_like(F)
    V = np.zeros_like(F)
    S = np.zeros((t, dim))
    lib.hs_svd.argtypes = [ctypes.c_int, ctypes.c_int64, _dp, _dp, _dp, _dp]
    lib.hs_svd(dim, t, _p(F), _p(U), _p(S), _p(V))
    return U, S, V


def element_hessian(material, psd_mode, F, mu, lam, vol):
    lib = load()
    F = _c(F)
    t, dim, _ = F.shape
    b = dim * dim
    mu = _c(np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (t,)))
    lam = _c(np.broadcast_to(np.asarray(lam, dtype=np.float64).reshape(-1), (t,)))
    vol = _c(np.broadcast_to(np.asarray(vol, dtype=np.float64).reshape(-1), (t,)))
    H = np.zeros((t, b, b))
    lib.hs_element_hessian.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, _dp, _dp, _dp, _dp, _dp]
    lib.hs_element_hessian(material, psd_mode, dim, t, _p(F), _p(mu), _p(lam), _p(vol), _p(H))
    return H


def element_energy_gradient(material, F, mu, lam):
    lib = load()
    F = _c(F)
    t, dim, _ = F.shape
    mu = _c(np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (t,)))
    lam = _c(np.broadcast_to(np.asarray(lam, dtype=np.float64).reshape(-1), (t,)))
    psi = np.zeros(t)
    P = np.zeros_like(F)
    lib.hs_element_energy_gradient.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _dp, _dp, _dp, _dp, _dp]
    lib.hs_element_energy_gradient(material, dim, t, _p(F), _p(mu), _p(lam), _p(psi), _p(P))
    return psi, P
