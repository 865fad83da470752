#!/bin/bash
# r02ab: spectral clustering / cubature on the GPU; subspace tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace.py -m gpu -x -q > gpurun_out/r02ab_pytest_subspace.log 2>&1
echo "pytest subspace rc=$?"; tail -15 gpurun_out/r02ab_pytest_subspace.log | cut -c1-200
