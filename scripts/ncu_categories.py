#!/usr/bin/env python
"""Executed warp-instructions, FP64 instructions and stall samples of the pipelined assembly kernel per code region,
from the source page of a full ncu capture (no GPU needed):
    ncu -i gpurun_out/r03_full.ncu-rep --page source --csv --print-source sass,cuda > /tmp/src.csv
    python scripts/ncu_categories.py /tmp/src.csv
The source page counts every warp instruction 2.08 times in this capture (4,424 per element-warp after the
correction = smsp__inst_executed.sum / element-warps of the same launch); the line ranges follow csrc/kernels.cuh,
smallmat.cuh and materials.cuh as of the capture."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
NW = 503554.0 * 2.08
tot = collections.Counter(); smp = collections.Counter(); f64 = collections.Counter()
fn=None; hdr=None; cur=None; fpath=None
def cat(f, l):
    if f == "kernels.cuh":
        if l < 100: return "phase2 index helpers (pair_base...)"
        if l < 136: return "load/gather/F"
        if l < 260: return "element_math glue (P, grad, W)"
        if l < 345: return "local_block"
        if l < 440: return "K blocks + staging stores"
        if l < 520: return "phase 2 (blocks+verts)"
        if l < 600: return "finalize"
        return "pipeline control / prefetch"
    if f == "smallmat.cuh":
        if l < 130: return "matmul etc (smallmat<130)"
        if l < 182: return "rsqrt / schur2 rotation params"
        if l < 290: return "jacobi_eig (3x3 sym eig)"
        if l < 520: return "svd_rv (fp32 warm + fp64 sweeps)"
        return "smallmat >520 (psd/cholesky/etc)"
    if f == "materials.cuh": return "materials (principal hessian, pk1)"
    return f
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; hdr=None; continue
    if r[0] == "Line No": hdr=r; iI=r.index("Instructions Executed"); iS=r.index("# Samples"); continue
    if hdr is None or "assemble_pipelined" not in fn: continue
    if r[0] != "": cur=(fpath,int(r[0])); continue
    ins = r[3].strip().split()
    if not ins: continue
    if ins[0].startswith("@"): ins=ins[1:]
    op=ins[0].split(".")[0]
    try: n=int(r[iI]); s_=int(r[iS])
    except ValueError: continue
    c=cat(*cur); tot[c]+=n; smp[c]+=s_
    if op in ("DFMA","DMUL","DADD","DSETP"): f64[c]+=n
T=sum(tot.values()); S=sum(smp.values())
print("%-45s %9s %9s %7s" % ("category","instr/el","fp64/el","stall%"))
for c,v in sorted(tot.items(), key=lambda kv:-smp[kv[0]]):
    print("%-45s %9.0f %9.0f %6.1f%%" % (c, v/NW, f64[c]/NW, 100.0*smp[c]/S))
print("%-45s %9.0f %9.0f" % ("total", T/NW, sum(f64.values())/NW))
