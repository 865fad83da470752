#!/usr/bin/env python
"""Opcode histogram per phase from an .ncu-rep source page. usage: ncu_opcodes.py rep [n_elem_warps]"""
import csv, collections, io, subprocess, sys
rep = sys.argv[1]
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 503554.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
iI = h.index('Instructions Executed'); iSrc = h.index('Source'); iS = h.index('# Samples')
bars = [k for k, r in enumerate(data) if 'BAR.SYNC' in r[iSrc]]
cut = bars[-1] if bars else len(data)
def op(src):
    t = src.strip().split()
    if t[0].startswith('@'): t = t[1:]
    return t[0].split('.')[0]
for name, lo, hi in (('phase1', 0, cut), ('phase2', cut, len(data))):
    c = collections.Counter(); sm = collections.Counter()
    for r in data[lo:hi]:
        c[op(r[iSrc])] += int(r[iI]); sm[op(r[iSrc])] += int(r[iS])
    tot = sum(c.values())
    print(name, 'total warp-instr', tot, 'per element-warp %.0f' % (tot / nw))
    for k, v in c.most_common(16):
        print('   %-10s %12d  %7.1f/elemwarp  samples %d' % (k, v, v / nw, sm[k]))
