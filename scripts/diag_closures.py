#!/usr/bin/env python
"""cProfile of one backward-Euler Newton step through reference-style closures at a bench workload (where does the host
side of the lazy-Hessian path spend its time?)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simkit_b200 as sk
from simkit_b200 import synthetic as syn

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
cfg = syn.CONFIGS[wl]
X, T = syn.make_mesh(wl)
dim = X.shape[1]
U = syn.jittered_state(X, cfg["cells"], cfg["extent"], sigma=0.1)
mu, lam = syn.lame()
J = sk.deformation_jacobian(X, T)
plan = J._skb_plan
vol = plan.volume()
mass = np.repeat(plan.vertex_masses(1e3), dim)
fg = np.zeros((plan.n, dim)); fg[:, 1] = -9.8
fg = (fg.reshape(-1) * mass).reshape(-1, 1)
M = sps.diags(mass).tocsc()
h = 1e-2
E = lambda x: sk.stable_neo_hookean_energy_x(x.reshape(-1, dim), J, mu, lam, vol) - float((fg.T @ x.reshape(-1, 1)).item())
G = lambda x: sk.stable_neo_hookean_gradient_x(x.reshape(-1, dim), J, mu, lam, vol) - fg
H = lambda x: sk.stable_neo_hookean_hessian_x(x.reshape(-1, dim), J, mu, lam, vol)
x0 = np.ascontiguousarray(U.reshape(-1, 1))
for s in range(2):
    t0 = time.perf_counter(); sk.backward_euler(x0, x0, E, G, H, M, h, max_iter=1, pcg_rtol=1e-10); print("warm-up %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter(); sk.backward_euler(x0, x0, E, G, H, M, h, max_iter=1, pcg_rtol=1e-10); dt = time.perf_counter() - t0
pr.disable()
print("profiled step %.1f ms" % (dt * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
