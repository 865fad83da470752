#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_solvers.py -m gpu -x -q > gpurun_out/r02ah_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02ah_pytest.log
