#!/bin/bash
# r02t: rolled sweep loops (code 99 KB -> 52 KB); unpooled + late hook (wsA), + corner-0 pairs on the reducers (wsB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02t_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02t_pytest_variants.log | tail -4
SKB_ASSEMBLE=pipe AB_SUFFIX=_r02t_pipe bash scripts/ab.sh main
SKB_ASSEMBLE=ws AB_SUFFIX=_r02t_ws bash scripts/ab.sh main
AB_SUFFIX=_r02t bash scripts/ab.sh wsA wsB
