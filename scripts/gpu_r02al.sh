#!/bin/bash
# r02al: verification of the round's final build: smoke, the default bench line (now with the C4 reduced leg), the ncu
# launch list of the bench command, the whole GPU suite
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02al_bench.json 2> gpurun_out/r02al_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02al_bench.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/r02al_bench.json")); r = d["roofline"]; n = d["newton"]; rd = d.get("reduced") or {}
print("step %.3f ms" % d["ms_per_step"], r["step_kernels_ms"], "frac", r["frac"], "traffic", r["traffic"], r["dominant_kernel"])
print("e2e", d["e2e"]["ms_per_step"], "newton", n["ms_per_step"], n["pcg_iters"], "closures", n.get("through_reference_closures", {}).get("ms_per_step"), "parity", d["parity_check"]["ok"])
print("reduced", rd.get("contraction_ms"), rd.get("api_resident_basis_ms"), rd.get("frac_fp64_peak"), (rd.get("roofline") or {}).get("frac"), rd.get("parity"))
PY
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02al_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-closures --no-reduced > gpurun_out/r02al_launches.log 2>&1
echo "launch list rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02al_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/r02al_pytest_gpu.log
