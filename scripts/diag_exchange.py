#!/usr/bin/env python
"""Where a sharded assembly step spends its time (run under torchrun): GPU time of the local kernel chain, the pack
kernels, the grouped NCCL send/recv and the scatter-adds (CUDA events on the current stream), and the HOST time
the Python side needs to issue one step.  Prints one line per rank."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from simkit_b200 import sharding, synthetic as syn
    from simkit_b200._lib import MATERIAL_IDS, PSD_AFTER_VOL, check, load
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = load()
    wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
    sh = sharding.make_shard(wl, rank, world, device=local)
    plan = sh.plan
    mu, lam = syn.lame()
    plan.set_materials(mu, lam, plan.volume())
    f64 = torch.float64
    x_d = torch.from_numpy(np.ascontiguousarray(sh.U_local.reshape(-1))).to(dev)
    g_d = torch.empty(plan.ndof, dtype=f64, device=dev)
    v_d = torch.empty(plan.nnz, dtype=f64, device=dev)
    mat = MATERIAL_IDS["stable_neo_hookean"]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    st = torch.cuda.current_stream()

    def step(timed):
        marks = [ev() for _ in range(5)] if timed else None
        t0 = time.perf_counter()
        if timed:
            marks[0].record(st)
        check(lib.skb_gradient_hessian_dev(plan._h, mat, PSD_AFTER_VOL, x_d.data_ptr(), None, g_d.data_ptr(), v_d.data_ptr(), st.cuda_stream))
        if timed:
            marks[1].record(st)
        ops = []
        for q, (gi, hi) in sorted(sh.send.items()):
            buf = sh.sbuf[q]
            check(lib.skb_gather_dev(g_d.data_ptr(), gi.data_ptr(), gi.numel(), buf.data_ptr(), st.cuda_stream))
            check(lib.skb_gather_dev(v_d.data_ptr(), hi.data_ptr(), hi.numel(), buf.data_ptr() + 8 * gi.numel(), st.cuda_stream))
            ops.append(dist.P2POp(dist.isend, buf, q))
        for p in sorted(sh.recv):
            ops.append(dist.P2POp(dist.irecv, sh.rbuf[p], p))
        if timed:
            marks[2].record(st)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if timed:
            marks[3].record(st)
        for p, (gi, hi) in sorted(sh.recv.items()):
            buf = sh.rbuf[p]
            check(lib.skb_scatter_add_dev(g_d.data_ptr(), gi.data_ptr(), gi.numel(), buf.data_ptr(), st.cuda_stream))
            check(lib.skb_scatter_add_dev(v_d.data_ptr(), hi.data_ptr(), hi.numel(), buf.data_ptr() + 8 * gi.numel(), st.cuda_stream))
        if timed:
            marks[4].record(st)
        return time.perf_counter() - t0, marks

    for _ in range(5):
        step(False)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    host, segs = [], []
    for _ in range(20):
        h, m = step(True)
        host.append(h)
        segs.append(m)
    torch.cuda.synchronize()
    seg = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in segs]).mean(axis=0)
    # free-running throughput (what bench.py times)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(st)
    t0 = time.perf_counter()
    for _ in range(20):
        step(False)
    issue = (time.perf_counter() - t0) / 20
    e1.record(st)
    torch.cuda.synchronize()
    line = ("rank %d: GPU ms  kernels %.3f  pack %.3f  nccl %.3f  scatter %.3f | host issue %.3f ms/step (timed loop %.3f) | "
            "free-running %.3f ms/step | send %.1f MB" % (rank, seg[0], seg[1], seg[2], seg[3], issue * 1e3, np.mean(host) * 1e3,
                                                         e0.elapsed_time(e1) / 20, sh.exchange_bytes / 1e6))
    gathered = [None] * world
    dist.all_gather_object(gathered, line)
    os.dup2(saved, 1)
    if rank == 0:
        print("\n".join(gathered), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
