#!/bin/bash
# r02aj: phase-2 unroll factor 2 (default) / 3 / 4; kernel-variant tests with the new default
mkdir -p gpurun_out
AB_SUFFIX=_r02aj bash scripts/ab.sh main u3 u4
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02aj_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02aj_pytest.log
