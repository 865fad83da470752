#!/bin/bash
# r02q: warp-specialised assembly kernel (SKB_ASSEMBLE=ws): parity tests first (under a timeout: new barrier protocol),
# then the same bench with the pipelined (default) and the warp-specialised kernel
mkdir -p gpurun_out
SKB_ASSEMBLE=ws timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02q_pytest_ws.log 2>&1
echo "pytest ws rc=$?"; tail -3 gpurun_out/r02q_pytest_ws.log
for k in pipe ws; do
  SKB_ASSEMBLE=$k timeout 300 python bench.py --newton 0 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r02q_bench_$k.json 2> gpurun_out/r02q_bench_$k.err
  echo "$k rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02q_bench_$k.json")); r = d["roofline"]
    print("$k", "step %.3f ms" % d["ms_per_step"], r["step_kernels_ms"], (d.get("parity_check") or {}).get("ok"))
except Exception as ex:
    print("$k FAILED", ex)
PY
  tail -3 gpurun_out/r02q_bench_$k.err | cut -c1-300
done
