#!/bin/bash
# r02w: TMA-streamed level 2 with the metadata fetched by the TMA unit two pieces ahead
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -s > gpurun_out/r02w_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; grep -h "largest\|passed\|failed\|Error" gpurun_out/r02w_pytest_variants.log | tail -4
AB_SUFFIX=_r02w timeout 600 bash scripts/ab.sh fbm fbm6
SKB_LIB_TAG=fbm timeout 600 ncu --clock-control none --set full --import-source on -k regex:'finalize_blocks_bulk' -s 2 -c 1 -f \
    -o gpurun_out/r02w_fin python bench.py --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/r02w_fin_ncu.log 2>&1
echo "ncu rc=$?"
