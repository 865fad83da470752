#!/bin/bash
# r02z: phase 2 with one lane per (entry, block row) in the warp-specialised kernel; row-padded records
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02z_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02z_pytest_variants.log | tail -4
AB_SUFFIX=_r02z timeout 900 bash scripts/ab.sh main p2ru2
