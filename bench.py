#!/usr/bin/env python
"""bench.py -- the hot-path benchmark of simkit_b200 (contract: task brief + SURVEY.md §8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: the fused stable
neo-Hookean gradient + PSD-projected Hessian + CSR assembly of every element of the workload mesh
(BASELINE.json `metric`, quoted on the 16M-tet config C5).  One JSON line is printed by rank 0:

  value     tets/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public host-pointer API (pinned host buffers, H2D of the state and
            D2H of gradient + CSR values inside the timed region)
  roofline  dominant kernel (assemble_ws_kernel) against the measured HBM peak; algorithmic bytes per
            tet from SURVEY.md §8(d); the FP64 fraction is reported beside it
  cpu_baseline  the numpy/scipy oracle (a port of the reference CPU path) on a bounded sample
  newton    one backward-Euler Newton step (assembly + block-Jacobi PCG + line search), steps/s

`--impl reference` times the reference algorithm's CPU path (the oracle port: the reference is pure
Python and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class StdoutToStderr:
    """rank 0 prints ONE JSON line on stdout.  NCCL writes its version banner ("NCCL version ...") to file
    descriptor 1 when the first communicator is created, so for N > 1 everything before the final print runs
    with fd 1 pointed at stderr; restore() puts stdout back for the JSON line."""

    def __init__(self, active):
        self.saved = None
        if active:
            sys.stdout.flush()
            self.saved = os.dup(1)
            os.dup2(2, 1)

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

MATERIAL = "stable_neo_hookean"
METRIC = "tets/s for stable-NH grad+PSD Hessian+CSR assembly"
CPU_SAMPLE = "C1"          # 20^3-cell cube, 48,000 tets: the reference's own CPU-runnable case


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_desc(name, t, n, nnz):
    from simkit_b200 import synthetic as syn
    cfg = syn.CONFIGS[name]
    cells = "x".join(str(c) for c in cfg["cells"])
    kind = "Kuhn 6-tet" if cfg["dim"] == 3 else "2-triangle"
    return "%s: %s-cell %s grid, %d elements, %d vertices, %d CSR non-zeros; state U = X + 0.1*cell*N(0,1) seed 0; ym=1e5 pr=0.45" % (
        name, cells, kind, t, n, nnz)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, active=True):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self.nv = None
        if not active:      # N > 1: rank 0 samples its GPU; eight sampler threads only compete with the launch threads
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_pass(sample=CPU_SAMPLE, reps=3):
    """gradient_x + hessian_x of the reference algorithm (oracle port) on a bounded sample mesh.
    Returns (tets/s best of reps, seconds per pass, t)."""
    from oracle import elasticity as oe
    from simkit_b200 import synthetic as syn
    cfg = syn.CONFIGS[sample]
    X, T = syn.make_mesh(sample)
    U = syn.jittered_state(X, cfg["cells"], cfg["extent"], sigma=0.1)
    mu, lam = syn.lame()
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    t = T.shape[0]
    best = float("inf")
    for r in range(reps + 1):                      # first pass is the warm-up
        t0 = time.perf_counter()
        oe.gradient_x(MATERIAL, U, J, mu, lam, vol)
        oe.hessian_x(MATERIAL, U, J, mu, lam, vol, psd=True)
        dt = time.perf_counter() - t0
        if r > 0:
            best = min(best, dt)
    return t / best, best, t


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info()]
        return max(n) if n else 1
    except Exception:
        return 1


def run_reference(args, rank):
    """`--impl reference`: the reference algorithm's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    from simkit_b200 import synthetic as syn
    times = []
    tps, sec, t = None, None, None
    for s in range(args.warmup + args.steps):
        tps, sec, t = cpu_reference_pass(CPU_SAMPLE, reps=1)
        if s >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    value = t / sec
    cfg = syn.CONFIGS[args.workload]
    sample = ("each step = gradient_x + hessian_x(psd) of the %s mesh (%d tets) in numpy/scipy: SpMV, batched element "
              "formulas, LAPACK eigh per 9x9, block_diag, two SpGEMMs -- the reference's algorithm; it is linear in t "
              "(SURVEY 6), so tets/s carries to %s" % (CPU_SAMPLE, t, args.workload))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s (%s cells), timed on a bounded sample" % (args.workload, "x".join(map(str, cfg["cells"]))),
                   "material": MATERIAL},
        "cpu_baseline": {"value": value, "unit": "tets/s", "cores": 1, "kind": "port", "sample": sample,
                         "blas_threads": host_threads(), "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "tets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def load_traffic(kernel_key):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def parity_assembly(workload, g_rows, Q_rows, col_ids, n_layers=2):
    """Checker (not timed, not part of the product path): gradient rows and Hessian block rows of the first ``n_layers``
    vertex planes of the workload mesh, as computed on the device, against the oracle on the closed sub-mesh made of
    the first ``n_layers`` cell layers (the rows of those vertices receive contributions from no other element).
    ``g_rows``: device gradient of the first rows; ``Q_rows``: scipy CSR of the same rows with local column numbering,
    ``col_ids`` the global vertex id of every local column vertex.  Returns (max rel gradient, max rel Hessian)."""
    import warnings
    warnings.filterwarnings("ignore")
    import scipy.sparse as sps
    from oracle import elasticity as oe
    from simkit_b200 import synthetic as syn
    cfg = syn.CONFIGS[workload]
    cells, dim = cfg["cells"], cfg["dim"]
    Ts = syn.grid_elements(cells, 0, n_layers)
    nsub = int(Ts.max()) + 1
    Xs = syn.grid_vertices(cells, cfg["extent"], np.arange(nsub))
    Us = syn.jittered_state_rows(cells, cfg["extent"], np.arange(nsub), sigma=0.1)
    mu, lam = syn.lame()
    Jo, volo = oe.deformation_jacobian(Xs, Ts), oe.volume(Xs, Ts)
    nfull = n_layers * int(np.prod([c + 1 for c in cells[1:]]))
    nr = nfull * dim
    go = np.asarray(oe.gradient_x(MATERIAL, Us, Jo, mu, lam, volo)).ravel()[:nr]
    Qo = oe.canonical_csr(oe.hessian_x(MATERIAL, Us, Jo, mu, lam, volo, psd=True))[:nr]
    rel_g = float(np.abs(np.asarray(g_rows).ravel()[:nr] - go).max() / np.abs(go).max())
    # map the device rows' local columns to global dof ids, then compare on the sub-mesh's columns
    gdof = (np.asarray(col_ids, dtype=np.int64)[:, None] * dim + np.arange(dim)[None, :]).ravel()
    Q = sps.csr_matrix(Q_rows)[:nr]
    Qg = sps.csr_matrix((Q.data, gdof[Q.indices], Q.indptr), shape=(nr, max(int(gdof.max()) + 1, nsub * dim)))
    d = Qg[:, : nsub * dim] - Qo
    rel_h = float(abs(d).max() / abs(Qo).max())
    outside = Qg[:, nsub * dim:]
    if outside.nnz:
        rel_h = max(rel_h, float(abs(outside).max() / abs(Qo).max()))
    return rel_g, rel_h


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import simkit_b200 as sk
    from simkit_b200 import _lib, synthetic as syn
    from simkit_b200._lib import MATERIAL_IDS, PSD_AFTER_VOL, check, ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; simkit_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    guard = StdoutToStderr(world > 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    cfg = syn.CONFIGS[args.workload]
    if world > 1:
        from simkit_b200 import sharding
        shard = sharding.make_shard(args.workload, rank, world, device=local_rank, interface=args.interface or None)
        plan, U = shard.plan, shard.U_local
        tt = torch.tensor([shard.nnz_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        t_total, n_total, nnz_total = shard.t_total, shard.n_total, int(tt.item())
    else:
        shard = None
        X, T = syn.make_mesh(args.workload)
        if args.element_order == "pencil":     # experiment: same mesh, elements listed in spatially compact runs
            T = np.ascontiguousarray(T[syn.pencil_order(cfg["cells"], cfg["extent"], X, T)])
        U = syn.jittered_state(X, cfg["cells"], cfg["extent"], sigma=0.1)
        if args.shuffle != "none":             # the caller lists the same mesh in random order
            rng = np.random.default_rng(17)
            T = np.ascontiguousarray(T[rng.permutation(T.shape[0])])
            if args.shuffle == "both":
                vp = rng.permutation(X.shape[0])
                inv = np.empty_like(vp)
                inv[vp] = np.arange(vp.size)
                X, U, T = np.ascontiguousarray(X[vp]), np.ascontiguousarray(U[vp]), np.ascontiguousarray(inv[T])
        # the reference's own set-up call (deformation_jacobian.py:9-87); the drop-in builds the device plan and hands it
        # on through J, the way every `*_x` function of the reference receives its mesh
        Jop = sk.deformation_jacobian(X, T)
        plan = Jop._skb_plan
        t_total, n_total, nnz_total = plan.t, plan.n, plan.nnz
    mu, lam = syn.lame()
    vol = plan.volume()
    mat = MATERIAL_IDS[MATERIAL]
    plan.set_materials(mu, lam, vol)

    f64 = torch.float64
    x_d = torch.from_numpy(np.ascontiguousarray(U.reshape(-1))).to(dev)
    g_d = torch.empty(plan.ndof, dtype=f64, device=dev)
    vals_d = torch.empty(plan.nnz, dtype=f64, device=dev)
    stream = torch.cuda.current_stream()

    def step_dev():
        if shard is not None:
            shard.gradient_hessian_dev(MATERIAL, PSD_AFTER_VOL, x_d, g_d, vals_d)
        else:
            check(lib.skb_gradient_hessian_dev(plan._h, mat, PSD_AFTER_VOL, x_d.data_ptr(), None, g_d.data_ptr(),
                                               vals_d.data_ptr(), stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=f64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident timing -------------------------------------------------------------
    # everything slow on the host (NVML initialisation of the clock sampler, its thread start) happens BEFORE the
    # barrier that precedes the first timed launch: a rank that enters the timed loop late makes its neighbours wait
    # in the interface exchange, and over K short steps that start-up skew would be billed to every step
    # (scripts/diag_exchange.py: 1.19 ms per step free-running at 8 GPUs)
    sampler = ClockSampler(local_rank, active=(rank == 0))
    for _ in range(args.warmup):
        step_dev()
    barrier()
    lib.skb_kernel_timing(plan._h, 1)
    l0 = plan.last_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_dev()
        e1.record(stream)
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = plan.last_launch_count() - l0
    kms = np.zeros(8)
    kcount = np.zeros(8, dtype=np.int64)
    check(lib.skb_kernel_times(plan._h, ptr(kms), ptr(kcount)))
    lib.skb_kernel_timing(plan._h, 0)
    ms_step = ms_total / args.steps
    value = t_total / (ms_step * 1e-3)
    if shard is not None:
        launches += shard.exchange_launches * args.steps

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tf = _lib.ctypes.c_double(0.0)
    check(lib.skb_fp64_peak(local_rank, _lib.ctypes.byref(tf)))
    fp64_peak = float(tf.value)
    dim = plan.dim
    # SURVEY 8(d) algorithmic bytes per element: T (int32 x K) + D (dim*dim f64) + (mu, lam, vol) + block slot map
    # (K*K int32) + x read and g written once per vertex + CSR values written once
    K = dim + 1
    per_elem = 4 * K + 8 * dim * dim + 24 + 4 * K * K
    alg_bytes = per_elem * plan.t + 2 * 8 * dim * plan.n + 8 * plan.nnz
    alg_flops = (2600.0 if dim == 3 else 700.0) * plan.t
    k_ms = kms[0] / max(int(kcount[0]), 1)
    # the library picks the assembly kernel (csrc/capi.cu launch_assemble_t): warp-specialised persistent kernel by
    # default, SKB_ASSEMBLE=pipe|tile for the round-1 kernels
    which = os.environ.get("SKB_ASSEMBLE", "ws")
    kernel_key = {"pipe": "assemble_pipelined_kernel<%d>", "tile": "assemble_tile_kernel<%d>"}.get(which, "assemble_ws_kernel<%d>") % dim
    # the algorithmic bytes (366 B/tet at C5) are what the STEP has to move: the CSR values among them are written by
    # finalize_blocks, not by the assembly kernel, so the fraction is quoted for the chain of kernels that makes up the
    # step (0.127, not the 0.178 that credited the assembly kernel with bytes it does not move)
    chain_ms = float(sum(kms[i] / max(int(kcount[i]), 1) for i in range(3)))
    achieved = alg_bytes / (chain_ms * 1e-3) / 1e9
    traffic = (load_traffic("step_ws" if which not in ("pipe", "tile") else "step")
               if (world == 1 and args.workload == "C5" and args.shuffle == "none") else None)
    roofline = {
        "kernel": "assembly step = %s + finalize_blocks_kernel + finalize_verts_kernel" % kernel_key,
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "peak_source": peak_src,
        "traffic": traffic,
        "traffic_source": "profiles/roofline_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the step's kernels, ncu --set full",
        "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_elem": alg_bytes / plan.t,
        "kernel_ms": chain_ms, "dominant_kernel": kernel_key, "dominant_kernel_ms": k_ms,
        "kernel_share_of_step": float(kms[0] / max(kms[:3].sum(), 1e-30)),
        "step_kernels_ms": {"assemble": kms[0] / max(int(kcount[0]), 1), "finalize_blocks": kms[1] / max(int(kcount[1]), 1),
                            "finalize_verts": kms[2] / max(int(kcount[2]), 1)},
        "step_frac_hbm": alg_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak if world == 1 else None,
        "fp64": {"achieved_tflops": alg_flops / (k_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                 "frac": alg_flops / (k_ms * 1e-3) / 1e12 / max(fp64_peak, 1e-30), "flops_per_elem": alg_flops / plan.t,
                 "peak_source": "measured here (skb_fp64_peak: 8 DFMA chains/thread, 2048 threads/SM)"},
    }

    # ---- end to end through the public host API -----------------------------------------------
    def pinned(n):
        return torch.empty(n, dtype=f64, pin_memory=True).numpy()

    e2e_steps = max(2, min(args.steps, 5))
    if args.no_e2e:          # development A/B runs only (scripts/ab.sh): the line then carries no e2e number
        print(json.dumps({"ms_per_step": ms_step, "roofline": roofline, "dev_only": True}), flush=True)
        return
    x_h, vol_h = pinned(plan.ndof), pinned(plan.t)
    x_h[:] = U.reshape(-1)
    vol_h[:] = vol.reshape(-1)
    e2e_out = {}
    if shard is None:
        # the reference's call pattern (energies/stable_neo_hookean.py:477-538): two calls with host arrays, a host
        # gradient and a host scipy matrix back.  `.data` forces the values of the lazy Hessian onto the host.
        grad_x, hess_x = getattr(sk, MATERIAL + "_gradient_x"), getattr(sk, MATERIAL + "_hessian_x")
        margs = (mu,) if MATERIAL == "arap" else (mu, lam)
        U_h = x_h.reshape(-1, dim)
        api = "sk.%s_gradient_x(U, J, mu, lam, vol) + sk.%s_hessian_x(U, J, mu, lam, vol).data (host arrays in, host ndarray + host csr_matrix out)" % (MATERIAL, MATERIAL)

        def step_e2e(touch=True):
            e2e_out["g"] = grad_x(U_h, Jop, *margs, vol_h)
            H = hess_x(U_h, Jop, *margs, vol_h)
            e2e_out["vals"] = H.data if touch else None
            e2e_out["H"] = H
    else:
        v0, v1 = shard.owned_value_range()
        o0, o1 = shard.layout.own_lo * dim, shard.layout.own_hi * dim
        g_h, vals_h = pinned(o1 - o0), pinned(v1 - v0)
        api = ("Shard.gradient_hessian per rank: H2D state, fused assembly (interface layer %s), D2H owned rows"
               % ("recomputed, no communication" if shard.interface == "recompute" else "exchanged over NCCL"))

        def step_e2e():
            shard.gradient_hessian(MATERIAL, x_h, mu, lam, vol_h, PSD_AFTER_VOL, g_out=g_h, vals_out=vals_h)

    for _ in range(2):
        step_e2e()
    barrier()
    with sampler:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    if shard is None:
        g_h, vals_h = e2e_out["g"].ravel(), e2e_out["vals"]
        h2d = 2 * 8 * (plan.ndof + plan.t + 2)       # each of the two calls uploads the state and the per-element weights
    else:
        h2d = 8 * (plan.ndof + plan.t + 2)
    d2h = 8 * (g_h.size + vals_h.size)
    if world > 1:
        tt = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        h2d, d2h = int(tt[0].item()), int(tt[1].item())
    e2e = {"value": t_total / (dt / e2e_steps), "unit": "tets/s", "ms_per_step": dt / e2e_steps * 1e3,
           "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": api}
    if shard is None:
        # the same two calls when the caller does not look at the values on the host (what a Newton loop does: the lazy
        # Hessian goes into `H + M/h**2` and the solve on the device); reported beside the contract number, not as it
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e(touch=False)
        torch.cuda.synchronize()
        e2e["values_left_on_device_ms_per_step"] = (time.perf_counter() - t0) / e2e_steps * 1e3
        e2e_out.clear()
    if shard is None:
        # the e2e and device-resident paths must agree bit for bit (same kernels, same reduction order)
        assert np.array_equal(vals_h, vals_d.cpu().numpy()) and np.array_equal(g_h, g_d.cpu().numpy())
    else:
        assert np.array_equal(vals_h, vals_d[v0:v1].cpu().numpy()) and np.array_equal(g_h, g_d[o0:o1].cpu().numpy())

    # ---- parity of the benched result against the oracle (checker; outside every timed region) ----------------
    parity = None
    if not args.no_parity and args.shuffle == "none" and args.element_order == "input":
        nl = 2
        rel_g = rel_h = None
        if rank == 0:
            cells = cfg["cells"]
            nfull = nl * int(np.prod([c + 1 for c in cells[1:]]))
            bptr, bcol = plan.block_pattern()
            nbr = np.diff(bptr[: nfull + 1]).astype(np.int64)
            nvals = int(bptr[nfull]) * dim * dim
            indptr = np.concatenate([[0], np.cumsum(np.repeat(nbr * dim, dim))])
            indices = np.empty(nvals, dtype=np.int64)
            pos = 0
            for v in range(nfull):
                cols = (bcol[bptr[v]:bptr[v + 1]].astype(np.int64)[:, None] * dim + np.arange(dim)[None, :]).ravel()
                for _i in range(dim):
                    indices[pos:pos + cols.size] = cols
                    pos += cols.size
            import scipy.sparse as sps
            Qrows = sps.csr_matrix((vals_d[:nvals].cpu().numpy(), indices, indptr), shape=(nfull * dim, plan.ndof))
            col_ids = np.arange(plan.n) if shard is None else shard.layout.l2g
            rel_g, rel_h = parity_assembly(args.workload, g_d[: nfull * dim].cpu().numpy(), Qrows, col_ids, nl)
        parity = {"assembly": {"what": "gradient rows and Hessian block rows of the first %d vertex planes (rank 0's rows) vs the "
                                       "oracle on the closed sub-mesh of the first %d cell layers" % (nl, nl),
                               "max_rel_gradient": rel_g, "max_rel_hessian": rel_h, "tol": 1e-10}}
        if rank == 0:
            parity["ok"] = bool(rel_g < 1e-10 and rel_h < 1e-10)
            assert parity["ok"], "parity check failed: %r" % (parity,)
    barrier()

    # ---- Newton step (assembly + PCG + line search), device-resident -------------------------------
    newton = None
    rho, h = 1e3, 1e-2
    if args.newton and shard is None:
        mass = np.repeat(plan.vertex_masses(rho), dim)
        fext = np.zeros((plan.n, dim))
        fext[:, 1] = -9.8
        fext = fext.reshape(-1) * mass
        xc = U.reshape(-1)
        nsteps = max(1, min(args.steps, args.newton_steps))
        spmv_bytes = (8 * dim * dim + 4) * plan.nnzb + 8 * plan.n + 3 * 8 * plan.ndof

        def newton_leg(label):
            lib.skb_kernel_timing(plan._h, 1)
            tt, info, xn = [], None, None
            for s in range(1 + nsteps):
                t0 = time.perf_counter()
                xn, info = plan.newton(MATERIAL, xc, x_tilde=xc, mass=mass, kin_scale=1.0 / h ** 2, f_ext=fext, max_iter=1,
                                       pcg_rtol=args.pcg_rtol, pcg_max_iter=20000)
                if s > 0:
                    tt.append(time.perf_counter() - t0)
                else:
                    check(lib.skb_kernel_times(plan._h, ptr(kms), ptr(kcount)))   # drop the warm-up record
            nl = plan.last_launch_count()
            check(lib.skb_kernel_times(plan._h, ptr(kms), ptr(kcount)))
            lib.skb_kernel_timing(plan._h, 0)
            sec = float(np.mean(tt))
            spmv_ms = kms[4] / max(int(kcount[4]), 1)
            return xn, {"preconditioner": label, "steps_per_s": 1.0 / sec, "ms_per_step": sec * 1e3, "pcg_iters": info["pcg_iters"],
                        "pcg_rtol": args.pcg_rtol, "pcg_relres": info["pcg_relres"], "alpha": info["alphas"][:1],
                        "launches_per_step": nl, "includes": "host->device upload of state and device->host read of x_next",
                        "spmv": {"ms": spmv_ms, "achieved_gbs": spmv_bytes / (spmv_ms * 1e-3) / 1e9,
                                 "frac_hbm": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / hbm_peak, "bytes": spmv_bytes},
                        "pcg_ms_per_iter": (kms[4] + kms[5]) / max(info["pcg_iters"], 1) / nsteps}

        # block-Jacobi alone, then block-Jacobi + rigid-mode coarse correction (csrc/coarse.cuh): same Newton iterate
        x_bj, newton_bj = newton_leg("%dx%d block-Jacobi" % (dim, dim))
        # same rule as ElasticPotential(coarse="auto"): the coarse correction is switched on when block-Jacobi needed
        # more than MeshPlan.COARSE_MIN_ITERS iterations
        if args.aggregates < 0:
            want_agg = plan.auto_aggregates() if newton_bj["pcg_iters"] > plan.COARSE_MIN_ITERS else 0
        else:
            want_agg = args.aggregates
        if want_agg:
            n_agg = plan.set_coarse_space(X, want_agg)
            x_tl, newton = newton_leg("%dx%d block-Jacobi + rigid-body modes of %d vertex aggregates (two-level, additive)"
                                      % (dim, dim, n_agg))
            plan.set_coarse_space(None)
            newton["block_jacobi_only"] = {k: newton_bj[k] for k in ("steps_per_s", "ms_per_step", "pcg_iters", "pcg_ms_per_iter")}
            newton["iterate_difference_vs_block_jacobi"] = float(np.abs(x_tl - x_bj).max() / np.abs(x_bj).max())
        else:
            newton = newton_bj
        if not args.no_closures:
            # the same step through the reference's own call pattern: energy / gradient / Hessian closures around the
            # `*_x` functions plus gravity, handed to backward_euler (examples/interactive_demos/010_..._3D.py:81-116,
            # integrators/backward_euler.py:27-91).  The Hessian stays on the device through `H + M/h**2` and the solve.
            import scipy.sparse as sps
            Mv = sps.diags(mass).tocsc()
            fg = fext.reshape(-1, 1)
            e_x = getattr(sk, MATERIAL + "_energy_x")

            def E_cl(x):
                return e_x(x.reshape(-1, dim), Jop, *margs, vol_h) - float((fg.T @ x.reshape(-1, 1)).item())

            def G_cl(x):
                return grad_x(x.reshape(-1, dim), Jop, *margs, vol_h) - fg

            def H_cl(x):
                return hess_x(x.reshape(-1, dim), Jop, *margs, vol_h)

            x0c = np.ascontiguousarray(xc.reshape(-1, 1))
            tcl, xcl, icl = [], None, None
            for s in range(2 + nsteps):          # the first solve runs block-Jacobi and switches the coarse space on
                t0 = time.perf_counter()
                xcl, icl = sk.backward_euler(x0c, x0c, E_cl, G_cl, H_cl, Mv, h, max_iter=1, return_info=True,
                                             pcg_rtol=args.pcg_rtol)
                if s > 1:
                    tcl.append(time.perf_counter() - t0)
            plan.set_coarse_space(None)
            sec_cl = float(np.mean(tcl))
            xref = x_tl if want_agg else x_bj
            newton["through_reference_closures"] = {
                "api": "sk.backward_euler(x, x_prev, E, G, H, M, h) with E/G/H closures around sk.%s_{energy,gradient,hessian}_x - gravity" % MATERIAL,
                "steps_per_s": 1.0 / sec_cl, "ms_per_step": sec_cl * 1e3, "alpha": [float(a_) for a_ in icl["alphas"][:1]],
                "vs_device_resident_step": sec_cl * 1e3 / newton["ms_per_step"],
                "iterate_difference_vs_device_resident_step": float(np.abs(xcl.ravel() - xref.ravel()).max() / np.abs(xref).max())}
    elif args.newton:
        # sharded implicit step: device-resident state, distributed single-reduction PCG (peer memory or NCCL; see
        # newton.collectives_per_iter in the line)
        mass_d = shard.lumped_mass_dofs(rho)
        fext_d = torch.zeros(plan.n, dim, dtype=f64, device=dev)
        fext_d[:, 1] = -9.8
        fext_d = fext_d.reshape(-1) * mass_d
        nsteps = max(1, min(args.steps, args.newton_steps))
        if args.dist_solver:
            shard.solver = args.dist_solver
        n_agg = shard.set_coarse_space(min(729, max(8, n_total // 1000)) if args.aggregates < 0 else args.aggregates)
        tt, info = [], None
        for s in range(1 + nsteps):
            xs = x_d.clone()
            barrier()
            t0 = time.perf_counter()
            info = shard.newton_step(MATERIAL, xs, x_tilde_d=x_d, mass_d=mass_d, kin_scale=1.0 / h ** 2, fext_d=fext_d,
                                     max_iter=1, pcg_rtol=args.pcg_rtol)
            barrier()
            if s > 0:
                tt.append(time.perf_counter() - t0)
        sec = max_over_ranks(float(np.mean(tt)))
        if parity is not None:
            # the same step on ONE GPU (rank 0 builds the whole mesh's plan; the mesh fits) -> every rank compares its
            # owned rows of x_next; tolerance of north_star for Newton iterates: 1e-8 relative
            xref = torch.empty(n_total * dim, dtype=f64, device=dev)
            if rank == 0:
                Xf, Tf = syn.make_mesh(args.workload)
                Uf = syn.jittered_state(Xf, cfg["cells"], cfg["extent"], sigma=0.1).reshape(-1)
                pf = sk.MeshPlan(X=Xf, T=Tf, device=local_rank)
                pf.set_materials(mu, lam, pf.volume())
                massf = np.repeat(pf.vertex_masses(rho), dim)
                fextf = np.zeros((pf.n, dim))
                fextf[:, 1] = -9.8
                fextf = fextf.reshape(-1) * massf
                if n_agg:
                    pf.set_coarse_space(Xf, n_agg)
                xn, inf1 = pf.newton(MATERIAL, Uf, x_tilde=Uf, mass=massf, kin_scale=1.0 / h ** 2, f_ext=fextf, max_iter=1,
                                     pcg_rtol=args.pcg_rtol, pcg_max_iter=20000)
                xref.copy_(torch.from_numpy(np.ascontiguousarray(xn.reshape(-1))))
                del pf
            dist.broadcast(xref, 0)
            lay = shard.layout
            o0, o1 = lay.own_lo * dim, lay.own_hi * dim
            mine = xs[o0:o1]
            ref = xref[lay.v_lo * dim: lay.v_hi * dim]
            err = torch.stack([(mine - ref).abs().max(), ref.abs().max()])
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            rel_x = float(err[0].item() / err[1].item())
            parity["newton"] = {"what": "x_next of the sharded step (owned rows of every rank) vs the same step on one GPU",
                                "max_rel": rel_x, "tol": 1e-8}
            if rank == 0:
                parity["ok"] = bool(parity["ok"] and rel_x < 1e-8)
                assert parity["ok"], "parity check failed: %r" % (parity,)
            del xref
        newton = {"steps_per_s": 1.0 / sec, "ms_per_step": sec * 1e3, "pcg_iters": info["pcg_iters"],
                  "pcg_rtol": args.pcg_rtol, "pcg_relres": info["pcg_relres"], "alpha": info["alphas"][:1],
                  "preconditioner": "3x3 block-Jacobi + rigid-body modes of %d vertex aggregates (two-level, additive)" % n_agg,
                  "includes": "device-resident state per rank; assembly + interface exchange + distributed PCG + line search",
                  "pcg_ms_per_iter": sec * 1e3 / max(info["pcg_iters"], 1)}
        solver = getattr(shard, "solver", None) or os.environ.get("SKB_DIST_PCG", "peer")
        if solver.startswith("peer") and not getattr(shard, "_peer", False):
            solver = "pcg2 (NCCL; peer-memory set-up failed: %s)" % getattr(shard, "_peer_error", "?")
        nc_ = (6 if dim == 3 else 3) * n_agg
        newton["solver"] = solver
        if getattr(shard, "last_solve_ms", None):
            newton["solve_ms"] = shard.last_solve_ms      # rank 0's host-clock breakdown of the last solve
        newton["collectives_per_iter"] = (
            {"all_reduce": 1, "all_reduce_doubles": 4 + nc_, "grouped_send_recv": 1,
             "note": "single-reduction PCG (csrc/capi_pcg2.cu): gamma, delta, r.r and the restricted vector of the coarse "
                     "space share one ncclAllReduce; halo of u by one grouped ncclSend/ncclRecv; CUDA-graph replay"}
            if solver.startswith("pcg2") else
            {"all_reduce": 0, "grouped_send_recv": 0, "nccl_calls": 0, "peer_store_doubles_reduction": (4 + nc_) * world,
             "note": "single-reduction PCG over peer memory (csrc/capi_pcg2.cu, transport 1): halo values of u and the 4 + 6 n_agg "
                     "reduction partials are stored by the producing kernels straight into the other ranks' HBM over NVLink "
                     "(CUDA IPC mappings, flags); every rank sums the partials itself in rank order; CUDA-graph replay"}
            if solver.startswith("peer") else
            {"all_reduce": 3 if n_agg else 2, "all_reduce_doubles": 3 + nc_, "grouped_send_recv": 1,
             "note": "textbook PCG loop (%s)" % solver})

    # ---- reduced (subspace) Hessian B^T H B, BASELINE config 4 (C4 mesh, r = 200) --------------------
    reduced = None

    def run_reduced(plan, X, r, Bm):
        z = 0.02 * np.random.default_rng(3).standard_normal(r)
        tt, tt_res = [], []
        times = np.zeros(3)
        for s_ in range(2):
            t0 = time.perf_counter()
            Er, gr_, Hr0 = plan.reduced(MATERIAL, Bm, z, x0=X.reshape(-1), psd_mode=PSD_AFTER_VOL)
            tt.append(time.perf_counter() - t0)
        plan.set_basis(Bm)                      # basis resident on the device, as ElasticEnergyZPrecomp keeps JB
        for s_ in range(3):
            t0 = time.perf_counter()
            Er, gr_, Hr = plan.reduced(MATERIAL, None, z, x0=X.reshape(-1), psd_mode=PSD_AFTER_VOL)
            tt_res.append(time.perf_counter() - t0)
            check(lib.skb_reduced_last_times(ptr(times)))
        assert np.array_equal(Hr, Hr0)
        plan.set_basis(None)
        b = dim * dim
        flops = (2.0 * b * b * r + 2.0 * b * r * r) * plan.t        # Y = He JB, Hr += JB^T Y (no symmetry assumed)
        tfd = _lib.ctypes.c_double(0.0)
        check(lib.skb_dmma_peak(local_rank, _lib.ctypes.byref(tfd)))
        dmma_peak = float(tfd.value)
        # what the kernel executes: Y on the FMA pipe, the block pairs bi <= bj of 5x5 tiles on DMMA
        rt_ = (r + 7) // 8
        nblk_ = (rt_ + 4) // 5
        exec_flops = (2.0 * b * b * r + 2.0 * 4 * ((2 * b + 3) // 4) * (nblk_ * (nblk_ + 1) // 2) * 25 * 64 / 2.0) * plan.t
        out = {"r": r, "elements": plan.t, "api_ms": min(tt) * 1e3,
                   "api_includes": "host->device copy of the basis B (%.2f GB) and of z, device->host copy of Hr" % (Bm.nbytes / 1e9),
                   "api_resident_basis_ms": min(tt_res) * 1e3,
                   "element_pass_ms": float(times[0]), "contraction_ms": float(times[1]), "device_ms": float(times[2]),
                   "contraction_tflops": flops / (times[1] * 1e-3) / 1e12, "algorithmic_flops": flops,
                   "frac_fp64_peak": flops / (times[1] * 1e-3) / 1e12 / max(fp64_peak, 1e-30),
                   "dmma_peak_tflops": dmma_peak, "executed_flops": exec_flops,
                   "executed_tflops": exec_flops / (times[1] * 1e-3) / 1e12,
                   "roofline": {"bound": "tensor", "achieved": exec_flops / (times[1] * 1e-3) / 1e12, "peak": dmma_peak,
                                "unit": "TFLOP/s", "frac": exec_flops / (times[1] * 1e-3) / 1e12 / max(dmma_peak, 1e-30)},
                   "peak_source": "FP64 FMA peak (skb_fp64_peak) and DMMA.8x8x4 peak (skb_dmma_peak) measured here; "
                                  "algorithmic_flops counts the full product, executed_flops the 15 of 25 block pairs "
                                  "the symmetric kernel computes (rows padded to the DMMA k = 4)",
                   "hr_symmetry_defect": float(np.abs(Hr - Hr.T).max() / np.abs(Hr).max())}
        return out, z, Hr

    if args.reduced and shard is None:
        reduced, _, _ = run_reduced(plan, X, args.reduced, syn.smooth_modes(X, args.reduced, seed=2))
    elif (shard is None and world == 1 and args.workload == "C5" and not args.no_reduced and args.shuffle == "none"
          and args.element_order == "input"):
        # BASELINE config 4 inside the default run, so that the driver's own bench measures the reduced tier: the C4
        # mesh (88^3 cells, 4,088,832 tets), r = 200 smooth modes, B^T H B through the plan-resident basis; parity of a
        # closed element sub-block against the oracle (the per-element weight is zero elsewhere: the `_z` tier floors
        # before the weight, elastic.py:663-664), outside every timed region
        c4 = syn.CONFIGS["C4"]
        X4, T4 = syn.make_mesh("C4")
        plan4 = sk.MeshPlan(X=X4, T=T4, device=local_rank)
        vol4 = plan4.volume()
        plan4.set_materials(mu, lam, vol4)
        r4 = 200
        B4 = syn.cos_modes(X4, r4, seed=2)
        reduced, z4, _ = run_reduced(plan4, X4, r4, B4)
        reduced["workload"] = "C4: %dx%dx%d-cell Kuhn grid, %d tets, r = %d cosine modes" % (tuple(c4["cells"]) + (plan4.t, r4))
        if not args.no_parity:
            nl = 2
            tsub = 6 * nl * int(np.prod(c4["cells"][1:]))
            w4 = np.zeros_like(vol4)
            w4[:tsub] = vol4[:tsub]
            plan4.set_materials(mu, lam, w4)
            E4, g4, H4 = plan4.reduced(MATERIAL, B4, z4, x0=X4.reshape(-1), psd_mode=2)
            Ts = T4[:tsub]
            nsub = int(Ts.max()) + 1
            Bs = B4[: nsub * 3]
            from oracle import elasticity as oe
            Jo, volo = oe.deformation_jacobian(X4[:nsub], Ts), oe.volume(X4[:nsub], Ts)
            xo = (Bs @ z4).reshape(-1, 3) + X4[:nsub]
            Ho = Bs.T @ (oe.hessian_x(MATERIAL, xo, Jo, mu, lam, volo, psd_before_vol=True) @ Bs)
            rel_r = float(np.abs(H4 - Ho).max() / np.abs(Ho).max())
            reduced["parity"] = {"what": "B^T H B of the first %d cell layers (%d tets) vs B^T Q_oracle B" % (nl, tsub),
                                 "max_rel": rel_r, "tol": 1e-10, "ok": bool(rel_r < 1e-10)}
            assert reduced["parity"]["ok"], "reduced parity check failed: %r" % (reduced["parity"],)
        del plan4, B4

    if args.reduced and shard is not None:
        # sharded reduced Hessian: every rank contracts its own elements, one all-reduce of 1 + r + r^2 doubles
        r = args.reduced
        ext = np.asarray(cfg["extent"], dtype=np.float64)
        Bl = syn.cos_modes(shard.X_local, r, seed=2, lo=np.zeros(dim), hi=ext, n_total=n_total)
        z = 0.02 * np.random.default_rng(3).standard_normal(r)
        plan.set_basis(Bl)
        x0l = shard.X_local.reshape(-1)
        tt_r, times = [], np.zeros(3)
        for s_ in range(4):
            barrier()
            t0 = time.perf_counter()
            Er, gr_, Hr = shard.reduced(MATERIAL, None, z, x0_local=x0l, psd_mode=PSD_AFTER_VOL)
            barrier()
            if s_ > 0:
                tt_r.append(time.perf_counter() - t0)
            check(lib.skb_reduced_last_times(ptr(times)))
        plan.set_basis(None)
        sec_r = max_over_ranks(min(tt_r))
        contr = max_over_ranks(float(times[1]))
        b = dim * dim
        flops = (2.0 * b * b * r + 2.0 * b * r * r) * t_total
        reduced = {"r": r, "elements": t_total, "api_resident_basis_ms": sec_r * 1e3, "contraction_ms_max_over_ranks": contr,
                   "contraction_tflops_all_ranks": flops / (contr * 1e-3) / 1e12, "algorithmic_flops": flops,
                   "all_reduce_bytes": shard.reduced_allreduce_bytes,
                   "hr_symmetry_defect": float(np.abs(Hr - Hr.T).max() / np.abs(Hr).max()),
                   "what": "Shard.reduced: per-rank B^T H B over the rank's own elements (basis rows of its local vertices "
                           "resident), one all-reduce of 1 + r + r^2 doubles"}

    # ---- CPU baseline (rank 0, N = 1) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import warnings
        warnings.filterwarnings("ignore")
        tps, sec, tt_ = cpu_reference_pass(CPU_SAMPLE, reps=3)
        cpu = {"value": tps, "unit": "tets/s", "cores": 1, "kind": "port",
               "sample": "oracle gradient_x + hessian_x(psd) on the %s mesh (%d tets), best of 3 after a warm-up, %.2f s per pass; "
                         "numpy/scipy is effectively single-threaded on this path" % (CPU_SAMPLE, tt_, sec),
               "blas_threads": host_threads(), "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "tets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, t_total, n_total, nnz_total), "material": MATERIAL,
                       "psd": "analytic eigensystem, floor 1e-6 after vol",
                       "l2": "no flush needed: per-step inputs+outputs (>= %.1f GB) exceed the 126 MB L2" % (alg_bytes / 1e9),
                       "sharding": "none" if world == 1 else (
                           "contiguous element slabs; interface = %s" % (
                               "exchange: partial gradient rows and Hessian block-rows of the ghost vertices sent to their owner, "
                               "one grouped NCCL send/recv per assembly" if shard.interface == "exchange" else
                               "recompute: every rank also evaluates the one layer of its lower neighbour's elements that touches "
                               "its rows (ghost elements), no communication in the assembly; NVLink traffic only in the solve")),
                       "element_order": "caller: %s%s; plan: internal sort-tile-recursive order (SKB_ELEMENT_ORDER=%s)" % (
                           args.element_order if world == 1 else "input",
                           "" if args.shuffle == "none" else ", shuffled " + args.shuffle,
                           os.environ.get("SKB_ELEMENT_ORDER", "str"))},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "newton": newton, "parity_check": parity,
        }
        if reduced is not None:
            line["reduced"] = reduced
        guard.restore()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--material", default="stable_neo_hookean", help="constitutive model of the step (default: the headline metric's "
                    "stable_neo_hookean; BASELINE configs 2 and 3 also name neo_hookean and arap)")
    ap.add_argument("--newton", type=int, default=1, help="also time one backward-Euler Newton step (0 = skip)")
    ap.add_argument("--newton-steps", type=int, default=2)
    ap.add_argument("--pcg-rtol", type=float, default=1e-10)
    ap.add_argument("--aggregates", type=int, default=-1, help="vertex aggregates of the two-level PCG preconditioner "
                    "(-1: MeshPlan.auto_aggregates when block-Jacobi needs > 300 iterations, 729 at C5; 0: block-Jacobi only)")
    ap.add_argument("--interface", default="", choices=["", "exchange", "recompute"],
                    help="N > 1: how owned rows get the lower neighbour's interface contributions (Shard docstring); default recompute")
    ap.add_argument("--dist-solver", default="", choices=["", "python", "native", "native_graph", "pcg2", "pcg2_eager", "peer", "peer_eager"],
                    help="N > 1: distributed PCG variant (default pcg2: single-reduction, C++-driven, CUDA graph)")
    ap.add_argument("--no-reduced", action="store_true", help="skip the C4 reduced-Hessian leg of the default run")
    ap.add_argument("--reduced", type=int, default=0, help="also time the reduced Hessian B^T H B with this many modes (config 4: --workload C4 --reduced 200)")
    ap.add_argument("--element-order", default="input", choices=["input", "pencil"],
                    help="experiment (1 GPU): list the mesh's elements in 3x3-cell pencils (synthetic.pencil_order) instead of the "
                         "generator's cell-major order; fewer partial records per element")
    ap.add_argument("--shuffle", default="none", choices=["none", "elements", "both"],
                    help="list the mesh's elements (and vertices) in random order: the plan's internal spatial element order "
                         "keeps the step within a few percent of the generator's order (1 GPU)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check of the benched result")
    ap.add_argument("--no-closures", action="store_true", help="skip the Newton step through reference-style closures")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="development only: stop after the device-resident timing")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    globals()["MATERIAL"] = args.material
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
