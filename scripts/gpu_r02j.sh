#!/bin/bash
# r02j (1 GPU): lazy / solver tests after the integrator + diagonal-add changes; closure profile; default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_solvers.py tests/test_contact.py tests/test_zz_quadratic.py -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02j_pytest_gpu.log
timeout 600 python scripts/diag_closures.py C5 > gpurun_out/r02j_closures.log 2>&1; head -30 gpurun_out/r02j_closures.log
timeout 900 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -3 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench.json"))
print("step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("values_left_on_device_ms_per_step"))
print("newton", d["newton"]["ms_per_step"], d["newton"].get("through_reference_closures"))
PY
