"""The assembly kernels (one CTA per tile, pipelined persistent, warp-specialised persistent: csrc/kernels.cuh) run the
same per-tile schedule in the same summation order.  The two persistent kernels instantiate the same per-material phase
functions, so their results agree to rounding (bit for bit
wherever the compiler contracted the arithmetic the same way), and two runs of one kernel are bit-identical.  The kernel is chosen once per process
(SKB_ASSEMBLE), so each one runs in a child process (with a time limit: the warp-specialised kernel synchronises its
warpgroups with named barriers) and the parent compares the arrays."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = r"""
import sys, numpy as np
import simkit_b200 as sk
from simkit_b200 import synthetic as syn
from simkit_b200._lib import MATERIAL_IDS
out = sys.argv[1]
res = {}
for cells in ((9, 8, 7), (37, 29), (3, 3, 2), (24, 20, 18)):
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(5)
    U = X + 0.2 * syn.cell_size(cells, tuple(1.0 for _ in cells)) * rng.standard_normal(X.shape)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    big = T.shape[0] > 10000   # many tiles per CTA and many level-2 pieces: one material is enough
    for mat in (["stable_neo_hookean"] if big else sorted(MATERIAL_IDS)):
        for psd in ((1,) if big else (0, 1)):
            g, v = plan.gradient_hessian(mat, U, mu, lam, vol, psd)
            res["%s_%d_%s_g_%d" % (mat, psd, "x".join(map(str, cells)), dim)] = g
            res["%s_%d_%s_v_%d" % (mat, psd, "x".join(map(str, cells)), dim)] = v
    # gradient-only and Hessian-only calls, and the _u tier's per-element offset
    res["gonly_%d" % dim] = plan.gradient("stable_neo_hookean", U, mu, lam, vol)
    res["honly_%d" % dim] = plan.hessian_values("arap", U, mu, lam, vol, 1)
    Fbar = 0.05 * rng.standard_normal((T.shape[0], dim, dim)) + np.eye(dim)
    g, v = plan.gradient_hessian("stable_neo_hookean", U - X, mu, lam, vol, 1, Fbar=Fbar)
    res["fbar_g_%d" % dim], res["fbar_v_%d" % dim] = g, v
np.savez(out, **res)
"""


def _run(kernel, tmp_path, tag=""):
    out = str(tmp_path / ("vals_%s%s.npz" % (kernel, tag)))
    env = dict(os.environ, SKB_ASSEMBLE=kernel)
    r = subprocess.run([sys.executable, "-c", CHILD, out], cwd=os.path.dirname(HERE), env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out)


def test_assembly_kernels_agree(tmp_path):
    tile, pipe, ws = (_run(k, tmp_path) for k in ("tile", "pipe", "ws"))
    ws2 = _run("ws", tmp_path, tag="_again")
    assert sorted(pipe.files) == sorted(ws.files) == sorted(tile.files)
    worst = 0.0
    for k in pipe.files:
        # no atomics, fixed summation order: two runs of the warp-specialised kernel are bit-identical (a race in the
        # warpgroup hand-over would show here).  Plain neo-Hookean is NaN on inverted elements (log J), like the
        # reference: the NaNs must sit in the same places in every kernel.
        assert np.array_equal(ws[k], ws2[k], equal_nan=True), "%s differs between two runs of the warp-specialised kernel" % k
        ok = np.isfinite(pipe[k])
        assert np.array_equal(ok, np.isfinite(ws[k])) and np.array_equal(ok, np.isfinite(tile[k])), k
        if "neo_hookean" not in k or "stable" in k:
            assert ok.all(), k
        scale = max(np.abs(pipe[k][ok]).max(), 1e-300)
        # same schedule and summation order; the compiler may contract a*b+c differently in the two kernel bodies
        d = np.abs(pipe[k][ok] - ws[k][ok]).max() / scale
        worst = max(worst, d)
        assert d <= 1e-12, "%s: pipelined vs warp-specialised kernel, %g" % (k, d)
        assert np.abs(tile[k][ok] - pipe[k][ok]).max() <= 1e-12 * scale, "%s: tile kernel vs pipelined kernel" % k
    print("largest pipelined / warp-specialised difference: %g" % worst)


FST_CHILD = r"""
import sys, numpy as np
import simkit_b200 as sk
out = sys.argv[1]
res = {}
rng = np.random.default_rng(17)
for dim, t, m1, m2, ncl in ((3, 700, 37, 150, 9), (2, 900, 13, 40, 5), (3, 50, 3, 5, 40)):
    b = dim * dim
    A = rng.standard_normal((m1, b * t))
    B = rng.standard_normal((b * t, m2))
    l = rng.integers(0, ncl, size=t)
    l[:ncl] = np.arange(ncl)
    f = sk.fast_sandwich_transform_clustered(A, B, l, dim=dim)
    res["ARBs_%d_%d" % (dim, t)] = f.ARBs
    ref = np.zeros_like(f.ARBs)                      # fast_sandwich_transform_clustered.py:66-93, written as one einsum
    A4 = A.reshape(m1, t, dim, dim)
    B4 = B.reshape(t, dim, dim, m2)
    for c in range(ncl):
        ref[:, :, c] = np.einsum("peik,ejkq->pqij", A4[:, l == c], B4[l == c])
    assert np.abs(f.ARBs - ref).max() <= 1e-12 * np.abs(ref).max()
np.savez(out, **res)
"""


def test_fst_precompute_kernels_bit_identical(tmp_path):
    """The register-tiled FST kernel (default) sums every entry over its cluster's elements in the same order as the
    one-row-per-CTA kernel (SKB_FST=simple): identical bits; both equal the definition (einsum) to rounding.  Ragged
    sizes: rows not a multiple of the 8-row tile, columns not a multiple of 128, clusters with a single element."""
    outs = []
    for mode in ("tiled", "simple"):
        out = str(tmp_path / ("fst_%s.npz" % mode))
        env = dict(os.environ, SKB_FST=mode)
        r = subprocess.run([sys.executable, "-c", FST_CHILD, out], cwd=os.path.dirname(HERE), env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        outs.append(np.load(out))
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]), k
