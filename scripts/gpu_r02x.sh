#!/bin/bash
# r02x: default bench line with the warp-specialised kernel + a full ncu capture of the step's kernels
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/r02x_bench.json")); r = d["roofline"]; n = d["newton"]
print("step %.3f ms" % d["ms_per_step"], r["step_kernels_ms"], "frac", r["frac"], "e2e", d["e2e"]["ms_per_step"], "newton", n["ms_per_step"], n["pcg_iters"], "closures", n.get("through_reference_closures", {}).get("ms_per_step"), d["parity_check"]["ok"])
PY
timeout 600 ncu --clock-control none --set full --import-source on -k regex:'assemble_ws|finalize_blocks|finalize_verts' -s 6 -c 3 -f \
    -o gpurun_out/r02x_assemble python bench.py --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/r02x_assemble.log 2>&1
echo "ncu rc=$?"
