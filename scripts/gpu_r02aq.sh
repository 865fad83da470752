#!/bin/bash
# r02aq: level-2 mirror mode (lower blocks filled in destination order) on shuffled / ordered vertex numberings
mkdir -p gpurun_out
SKB_FINALIZE_MIRROR=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02aq_pytest_mirror.log 2>&1
echo "pytest (mirror forced) rc=$?"; tail -1 gpurun_out/r02aq_pytest_mirror.log
AB_ARGS="--shuffle both" AB_SUFFIX=_r02aq_shufB_auto bash scripts/ab.sh main
SKB_FINALIZE_MIRROR=0 AB_ARGS="--shuffle both" AB_SUFFIX=_r02aq_shufB_off bash scripts/ab.sh main
SKB_FINALIZE_MIRROR=1 AB_SUFFIX=_r02aq_ordered_on bash scripts/ab.sh main
AB_SUFFIX=_r02aq_ordered_auto bash scripts/ab.sh main
