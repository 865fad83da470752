#!/bin/bash
# r02ak: reduced tier with its work arrays kept in the plan between calls
mkdir -p gpurun_out
timeout 600 python scripts/diag_kernels.py reduced C4 200 > gpurun_out/r02ak_reduced_clean.log 2>&1; tail -3 gpurun_out/r02ak_reduced_clean.log
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_configs.py -m gpu -x -q -k "reduced or z_ or c4" > gpurun_out/r02ak_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02ak_pytest.log
