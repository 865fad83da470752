#!/usr/bin/env python
"""Static SASS opcode histogram of the kernels whose mangled name contains a pattern (no GPU needed):
    python scripts/sass_opcodes.py simkit_b200/libsimkit_b200.so assemble_pipelined_kernelILi3ELi3ELi2ELi0ELi128
Used next to `ptxas -v` (python -m simkit_b200.build --force --ptxas) to compare A/B builds before spending GPU time."""
import subprocess, sys, re, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; counts = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur and pat in cur:
        counts[cur][m.group(1).split(".")[0]] += 1
for f, c in counts.items():
    tot = sum(c.values())
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f[:90], "total", tot, "fp64", fp64)
    print("  ", ", ".join("%s %d" % kv for kv in c.most_common(22)))
