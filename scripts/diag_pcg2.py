#!/usr/bin/env python
"""One rank, whole mesh: the single-reduction distributed solve (csrc/capi_pcg2.cu) with a world of one, so that ncu can
list / capture its kernels at full size (no halo, the all-reduce is a no-op).  Used for the launch list and the full
captures under profiles/ (scripts/gpu_r02g.sh).

    python scripts/diag_pcg2.py [--workload C5] [--aggregates 343] [--solver pcg2_eager] [--steps 1]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C5")
    ap.add_argument("--aggregates", type=int, default=343)
    ap.add_argument("--solver", default="pcg2_eager")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--max-pcg", type=int, default=20000)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from simkit_b200 import sharding
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    sh = sharding.make_shard(a.workload, 0, 1, device=0)
    sh.solver = a.solver
    from simkit_b200 import synthetic as syn
    mu, lam = syn.lame()
    sh.set_materials(mu, lam)
    dim = sh.layout.dim
    x_d = torch.from_numpy(np.ascontiguousarray(sh.U_local.reshape(-1))).to(dev)
    mass_d = sh.lumped_mass_dofs(1e3)
    fext_d = torch.zeros(sh.plan.n, dim, dtype=torch.float64, device=dev)
    fext_d[:, 1] = -9.8
    fext_d = fext_d.reshape(-1) * mass_d
    n_agg = sh.set_coarse_space(a.aggregates)
    from simkit_b200._lib import check, load, ptr
    lib = load()
    names = ["assemble", "finalize_blocks", "finalize_verts", "energy", "spmv", "pcg_vector", "other", "?"]
    for s in range(a.steps):
        xs = x_d.clone()
        lib.skb_kernel_timing(sh.plan._h, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = sh.newton_step("stable_neo_hookean", xs, x_tilde_d=x_d, mass_d=mass_d, kin_scale=1e4, fext_d=fext_d, max_iter=1,
                              pcg_rtol=1e-10, pcg_max_iter=a.max_pcg)
        torch.cuda.synchronize()
        print("step %d: %.1f ms, %d PCG iterations, n_agg %d, solve %s" % (s, (time.perf_counter() - t0) * 1e3, info["pcg_iters"],
                                                                        n_agg, getattr(sh, "last_solve_ms", None)), flush=True)
        kms, kc = np.zeros(8), np.zeros(8, dtype=np.int64)
        check(lib.skb_kernel_times(sh.plan._h, ptr(kms), ptr(kc)))
        lib.skb_kernel_timing(sh.plan._h, 0)
        print("   CUDA-event time per kernel kind (eager launches only): " + ", ".join(
            "%s %.1f ms / %d = %.1f us" % (names[k], kms[k], kc[k], 1e3 * kms[k] / max(kc[k], 1)) for k in range(7) if kc[k]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
