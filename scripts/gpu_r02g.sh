#!/bin/bash
# r02g (1 GPU): GPU sharding tests that run on one GPU; per-kernel launch list of the single-reduction solve at full size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; tail -3 gpurun_out/r02g_pytest.log
timeout 300 python scripts/diag_pcg2.py --steps 2 --solver pcg2 > gpurun_out/r02g_diag_plain.log 2>&1; tail -3 gpurun_out/r02g_diag_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pcg2_|pcg3_|coarse_' -s 8 -c 60 --csv \
    --log-file gpurun_out/r02g_pcg2_launches.csv python scripts/diag_pcg2.py --max-pcg 12 > gpurun_out/r02g_diag_ncu.log 2>&1
tail -2 gpurun_out/r02g_diag_ncu.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02g_pcg2_launches.csv")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    k = d["Kernel Name"].split("(")[0][:60]
    m = d["Metric Name"]; v = float(d["Metric Value"].replace(",", ""))
    agg.setdefault(k, collections.defaultdict(list))[m].append(v)
for k, ms in agg.items():
    t = ms.get("gpu__time_duration.sum", [0]); rd = ms.get("dram__bytes_read.sum", [0]); wr = ms.get("dram__bytes_write.sum", [0])
    print("%-60s n=%2d  time %s  read %s  write %s" % (k, len(t), ["%.1f" % x for x in t[:3]], ["%.1f" % x for x in rd[:2]], ["%.1f" % x for x in wr[:2]]))
PY
