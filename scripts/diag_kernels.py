#!/usr/bin/env python
"""Drives one group of kernels at bench size so that ncu can list / capture them (scripts/gpu_profile_r02.sh):

    python scripts/diag_kernels.py newton   [C5]      device-resident Newton step: energy, PCG (SpMV, update, direction,
                                                       coarse restrict / GEMV / add) kernels
    python scripts/diag_kernels.py reduced  [C4] [r]  reduced Hessian B^T H B with r modes
    python scripts/diag_kernels.py fst      [t] [r]   fast_sandwich_transform_clustered precompute + eval on t cubature
                                                       elements and r modes
Prints wall-clock timings of the same calls (meaningless under ncu)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simkit_b200 as sk
from simkit_b200 import synthetic as syn

what = sys.argv[1] if len(sys.argv) > 1 else "newton"
if what == "newton":
    wl = sys.argv[2] if len(sys.argv) > 2 else "C5"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    cfg = syn.CONFIGS[wl]
    X, T = syn.make_mesh(wl)
    dim = X.shape[1]
    U = syn.jittered_state(X, cfg["cells"], cfg["extent"], sigma=0.1).reshape(-1)
    plan = sk.MeshPlan(X=X, T=T)
    mu, lam = syn.lame()
    plan.set_materials(mu, lam, plan.volume())
    mass = np.repeat(plan.vertex_masses(1e3), dim)
    fext = np.zeros((plan.n, dim)); fext[:, 1] = -9.8
    fext = fext.reshape(-1) * mass
    plan.set_coarse_space(X, plan.auto_aggregates())
    for s in range(2):
        t0 = time.perf_counter()
        x, info = plan.newton("stable_neo_hookean", U, x_tilde=U, mass=mass, kin_scale=1e4, f_ext=fext, max_iter=1, pcg_rtol=1e-10,
                              pcg_max_iter=iters)
        print("newton step %.1f ms, %d PCG iterations" % ((time.perf_counter() - t0) * 1e3, info["pcg_iters"]), flush=True)
elif what == "reduced":
    wl = sys.argv[2] if len(sys.argv) > 2 else "C4"
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    X, T = syn.make_mesh(wl)
    plan = sk.MeshPlan(X=X, T=T)
    mu, lam = syn.lame()
    plan.set_materials(mu, lam, plan.volume())
    B = syn.cos_modes(X, r, seed=2)
    z = 0.02 * np.random.default_rng(3).standard_normal(r)
    plan.set_basis(B)
    from simkit_b200._lib import check, load, ptr
    times = np.zeros(3)
    for s in range(int(sys.argv[4]) if len(sys.argv) > 4 else 3):
        t0 = time.perf_counter()
        E, g, H = plan.reduced("stable_neo_hookean", None, z, x0=X.reshape(-1))
        check(load().skb_reduced_last_times(ptr(times)))
        print("reduced r=%d: api %.1f ms, element pass %.2f ms, contraction %.2f ms" % (r, (time.perf_counter() - t0) * 1e3, times[0], times[1]), flush=True)
else:
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    rng = np.random.default_rng(0)
    dim = 3
    A = rng.standard_normal((r, 9 * t))
    B = rng.standard_normal((9 * t, r))
    ncl = 100
    l = rng.integers(0, ncl, size=t)
    for s in range(2):
        t0 = time.perf_counter()
        f = sk.fast_sandwich_transform_clustered(A, B, l, dim=dim)
        t1 = time.perf_counter()
        out = f(rng.standard_normal((ncl, 3, 3)))
        t2 = time.perf_counter()
        print("fst t=%d r=%d clusters=%d: precompute %.1f ms (%.2f TFLOP/s of 54 m1 m2 t flop, incl. host copies), eval %.2f ms"
              % (t, r, ncl, (t1 - t0) * 1e3, 54.0 * r * r * t / (t1 - t0) / 1e12, (t2 - t1) * 1e3), flush=True)
