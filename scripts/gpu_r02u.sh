#!/bin/bash
# r02u: ws kernel as the default; TMA-streamed level 2 (SKB_FINALIZE=bulk) vs thread-per-entry; variants test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02u_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02u_pytest_variants.log | tail -4
AB_SUFFIX=_r02u_items bash scripts/ab.sh main
SKB_FINALIZE=bulk AB_SUFFIX=_r02u_bulk timeout 300 bash scripts/ab.sh main
