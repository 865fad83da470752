"""Plane contact springs: oracle vs the frozen reference outputs on the CPU; on the GPU the
drop-in functions and a backward-Euler step of stable neo-Hookean + contact, host-callable and device-resident."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe

TAGS = ["contact_tet", "contact_tri"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_contact(golden_dir, tag):
    g = load(golden_dir, tag)
    M = sps.diags(g["mass"])
    E, gr, H, inds = oe.contact_springs_plane(g["U"], float(g["k"]), g["p"], g["n"], M)
    assert abs(E - float(g["E"])) <= 1e-13 * abs(float(g["E"])) and np.array_equal(inds, g["inds"])
    assert rel(gr, g["g"]) < 1e-13 and rel(H.toarray(), g["H"]) < 1e-13
    E2, g2, _, _ = oe.contact_springs_plane(g["U"], float(g["k"]), g["p"], g["n"])
    assert abs(E2 - float(g["E_noM"])) <= 1e-13 * abs(float(g["E_noM"])) and rel(g2, g["g_noM"]) < 1e-13
    assert oe.contact_springs_plane(g["U"] + 10.0 * g["n"], float(g["k"]), g["p"], g["n"], M)[0] == float(g["E_above"]) == 0.0


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_contact_sphere(golden_dir, tag):
    g = load(golden_dir, tag)
    M = sps.diags(g["mass"])
    E, gr, H = oe.contact_springs_sphere(g["U"], float(g["k"]), g["s_p"], float(g["s_r"]), M)
    assert abs(E - float(g["s_E"])) <= 1e-13 * abs(float(g["s_E"]))
    assert rel(gr, g["s_g"]) < 1e-13 and rel(H.toarray(), g["s_H"]) < 1e-13
    E2 = oe.contact_springs_sphere(g["U"], float(g["k"]), g["s_p"], float(g["s_r"]))[0]
    assert abs(E2 - float(g["s_E_noM"])) <= 1e-13 * abs(float(g["s_E_noM"]))
    assert oe.contact_springs_sphere(g["U"] + 10.0, float(g["k"]), g["s_p"], float(g["s_r"]), M)[0] == float(g["s_E_far"]) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_contact_sphere(golden_dir, tag):
    import simkit_b200 as sk
    g = load(golden_dir, tag)
    X, T, U = g["X"], g["T"], g["U"]
    dim = int(g["dim"])
    k, p, n = float(g["k"]), g["p"], g["n"]
    sc, sr = g["s_p"], float(g["s_r"])
    M = sps.diags(g["mass"])
    E = sk.contact_springs_sphere_energy(U, k, sc, sr, M)
    assert isinstance(E, float) and abs(E - float(g["s_E"])) <= 1e-12 * abs(float(g["s_E"]))
    assert rel(sk.contact_springs_sphere_gradient(U, k, sc, sr, M), g["s_g"]) < 1e-12
    H = sk.contact_springs_sphere_hessian(U, k, sc, sr, M)
    assert sps.issparse(H) and rel(H.toarray(), g["s_H"]) < 1e-12
    assert abs(sk.contact_springs_sphere_energy(U, k, sc, sr) - float(g["s_E_noM"])) <= 1e-12 * abs(float(g["s_E_noM"]))
    assert sk.contact_springs_sphere_energy(U + 10.0, k, sc, sr, M) == 0.0
    assert sk.contact_springs_sphere_hessian(U + 10.0, k, sc, sr, M).nnz == 0
    # plane + sphere together inside the device-resident backward-Euler step
    mu, lam, h = float(g["mu"]), float(g["lam"]), float(g["h"])
    Md = sps.kron(M, sps.identity(dim)).tocsc()
    pot = sk.ElasticPotential("stable_neo_hookean", mu, lam, X=X, T=T, f_ext=g["fg"], contact_plane=dict(k=k, p=p, n=n, M=M),
                              contact_sphere=dict(k=k, p=sc, r=sr, M=M))
    x_curr, x_prev = U.reshape(-1, 1), X.reshape(-1, 1)
    x1, info = sk.backward_euler(x_curr, x_prev, pot.energy, pot.gradient, pot.hessian, Md, h, max_iter=3, return_info=True,
                                 pcg_rtol=1e-13)
    assert list(info["alphas"]) == list(g["be2_alphas"]) and rel(x1, g["be2_x"]) < 1e-8
    x2 = sk.backward_euler(x_curr, x_prev, lambda x: pot.energy(x), lambda x: pot.gradient(x), lambda x: pot.hessian(x), Md, h,
                           max_iter=3, pcg_rtol=1e-13)
    assert rel(x2, g["be2_x"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_contact(golden_dir, tag):
    import simkit_b200 as sk
    g = load(golden_dir, tag)
    X, T, U = g["X"], g["T"], g["U"]
    dim = int(g["dim"])
    k, p, n = float(g["k"]), g["p"], g["n"]
    M = sps.diags(g["mass"])
    E, inds = sk.contact_springs_plane_energy(U, k, p, n, M, return_contact_inds=True)
    assert isinstance(E, float) and abs(E - float(g["E"])) <= 1e-12 * abs(float(g["E"]))
    assert np.array_equal(inds.ravel(), g["inds"])
    gr = sk.contact_springs_plane_gradient(U, k, p, n, M)
    assert gr.shape == g["g"].shape and rel(gr, g["g"]) < 1e-12
    H = sk.contact_springs_plane_hessian(U, k, p, n, M)
    assert sps.issparse(H) and rel(H.toarray(), g["H"]) < 1e-12
    assert abs(sk.contact_springs_plane_energy(U, k, p, n) - float(g["E_noM"])) <= 1e-12 * abs(float(g["E_noM"]))
    assert rel(sk.contact_springs_plane_gradient(U, k, p, n), g["g_noM"]) < 1e-12
    assert sk.contact_springs_plane_energy(U + 10.0 * n, k, p, n, M) == 0.0
    assert sk.contact_springs_plane_hessian(U + 10.0 * n, k, p, n, M).nnz == 0
    # backward Euler with contact: device-resident step (ElasticPotential) and the host-callable path
    mu, lam, h = float(g["mu"]), float(g["lam"]), float(g["h"])
    Md = sps.kron(M, sps.identity(dim)).tocsc()
    pot = sk.ElasticPotential("stable_neo_hookean", mu, lam, X=X, T=T, f_ext=g["fg"], contact_plane=dict(k=k, p=p, n=n, M=M))
    x_curr, x_prev = U.reshape(-1, 1), X.reshape(-1, 1)
    x1, info = sk.backward_euler(x_curr, x_prev, pot.energy, pot.gradient, pot.hessian, Md, h, max_iter=3, return_info=True,
                                 pcg_rtol=1e-13)
    assert list(info["alphas"]) == list(g["be_alphas"]) and rel(x1, g["be_x"]) < 1e-8
    x2 = sk.backward_euler(x_curr, x_prev, lambda x: pot.energy(x), lambda x: pot.gradient(x), lambda x: pot.hessian(x), Md, h,
                           max_iter=3, pcg_rtol=1e-13)
    assert rel(x2, g["be_x"]) < 1e-8
