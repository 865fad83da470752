#!/usr/bin/env python
"""Per-source-line stall-sample attribution from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] == "": continue                     # SASS rows; the line row already aggregates them
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    s = int(r[iS]); ins = int(r[iI])
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += ins; tot += s
print("total samples", tot)
for (f, ln, src), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%%  %9d smp %12d inst  %s:%d  %s" % (100.0 * s / max(tot, 1), s, ins, f, ln, src))
byfile = collections.Counter()
for (f, ln, src), (s, ins) in agg.items(): byfile[f] += s
print({k: "%.1f%%" % (100.0 * v / tot) for k, v in byfile.items()})
