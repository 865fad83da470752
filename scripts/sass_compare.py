#!/usr/bin/env python
"""Per-kernel comparison of the SASS of two builds of the library (no GPU needed):
    python scripts/sass_compare.py old/libsimkit_b200.so simkit_b200/libsimkit_b200.so
Used to show that changes made without a GPU left every GPU-verified kernel byte-identical."""
import subprocess, re, sys, hashlib, collections
def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d = collections.OrderedDict(); cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m: cur = m.group(1); d[cur] = []; continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur: d[cur].append(re.sub(r"\s+", " ", m.group(1)))
    return {k: (len(v), hashlib.md5("\n".join(v).encode()).hexdigest()) for k, v in d.items()}
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
print("kernels old %d new %d | identical SASS %d | changed %d | only old %d | only new %d" % (len(a), len(b), len(same), len(diff), len([k for k in a if k not in b]), len([k for k in b if k not in a])))
for k in diff: print("CHANGED", a[k][0], "->", b[k][0], k[:110])
for k in b:
    if k not in a: print("NEW", b[k][0], k[:110])
for k in a:
    if k not in b: print("GONE", a[k][0], k[:110])
