#!/bin/bash
# r02ae: element_math<GRAD_LAST>: the compute groups' wait for the staging area moved behind the Hessian part
mkdir -p gpurun_out
AB_SUFFIX=_r02ae timeout 600 bash scripts/ab.sh main gl
SKB_LIB_TAG=gl timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02ae_pytest_variants_gl.log 2>&1
echo "pytest variants (gl) rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02ae_pytest_variants_gl.log | tail -4
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02ae_pytest_variants.log 2>&1
echo "pytest variants (main) rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02ae_pytest_variants.log | tail -4
