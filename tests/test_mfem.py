"""MFEM blocks: the oracle replays the frozen reference outputs on the CPU; the GPU test runs
the same small mixed problem through simkit_b200 (stretch, dS/dF, ds/dz, symmetric stretch map, the `_S` tier of
the dispatcher, sqp_mfem with the GPU linear solves)."""
import os

import numpy as np
import pytest

from oracle import elasticity as oe
from oracle.mfem_problem import mfem_problem

TAGS = ["mfem_tri", "mfem_tet"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


def _dense(a):
    return a.toarray() if hasattr(a, "toarray") else np.asarray(a)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_mfem_blocks(golden_dir, tag):
    g = load(golden_dir, tag)
    dim = int(g["dim"])
    assert rel(oe.stretch(g["F"]), g["stretch"]) < 1e-12
    assert rel(oe.stretch_gradient_dF(g["F"]), g["dSdF"]) < 1e-11
    prob = mfem_problem(oe.MfemSurface, g["X"], g["T"], float(g["rho_aug"]))
    p0 = g["p0"]
    assert rel(prob["p0"], p0) == 0.0
    dsdz = oe.stretch_gradient_dz(p0[:prob["nz"]], prob["GJB"], dim, Ci=prob["Ci"], GJq=prob["GJq"])
    assert rel(_dense(dsdz), g["dsdz"]) < 1e-11
    assert abs(prob["energy"](p0) - float(g["energy0"])) <= 1e-12 * abs(float(g["energy0"]))
    fu, fz, fmu = prob["grad_blocks"](p0)
    assert rel(fu, g["f_u0"]) < 1e-11 and rel(fz, g["f_z0"]) < 1e-11 and rel(fmu, g["f_mu0"]) < 1e-11
    p3 = oe.sqp_mfem(p0, prob["energy"], prob["hess_blocks"], prob["grad_blocks"], tolerance=1e-12, max_iter=3)
    assert rel(p3, g["p3"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_mfem_blocks(golden_dir, tag):
    import simkit_b200 as sk
    g = load(golden_dir, tag)
    dim = int(g["dim"])
    s = sk.stretch(g["F"])
    assert s.shape == g["stretch"].shape and rel(s, g["stretch"]) < 1e-10
    d = sk.stretch_gradient_dF(g["F"])
    assert d.shape == g["dSdF"].shape and rel(d, g["dSdF"]) < 1e-10
    assert rel(sk.stretch_gradient(g["F"]), g["dSdF"]) < 1e-10
    Se, Sei = sk.symmetric_stretch_map(5, dim)
    So, Soi = oe.symmetric_stretch_map(5, dim)
    assert abs(Se - So).max() == 0.0 and abs(Sei - Soi).max() == 0.0
    prob = mfem_problem(sk, g["X"], g["T"], float(g["rho_aug"]))
    p0 = g["p0"]
    dsdz = sk.stretch_gradient_dz(p0[:prob["nz"]], prob["GJB"], dim, Ci=prob["Ci"], GJq=prob["GJq"])
    assert rel(_dense(dsdz), g["dsdz"]) < 1e-10
    assert abs(prob["energy"](p0) - float(g["energy0"])) <= 1e-11 * abs(float(g["energy0"]))
    fu, fz, fmu = prob["grad_blocks"](p0)
    assert rel(fu, g["f_u0"]) < 1e-10 and rel(fz, g["f_z0"]) < 1e-10 and rel(fmu, g["f_mu0"]) < 1e-10
    p3 = sk.sqp_mfem(p0, prob["energy"], prob["hess_blocks"], prob["grad_blocks"], tolerance=1e-12, max_iter=3)
    assert rel(p3, g["p3"]) < 1e-8
