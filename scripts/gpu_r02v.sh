#!/bin/bash
# r02v: smaller level-2 pieces (64 slots); the whole GPU suite with the warp-specialised kernel as the default
mkdir -p gpurun_out
AB_SUFFIX=_r02v bash scripts/ab.sh fb64 fb64n6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/r02v_pytest_gpu.log
