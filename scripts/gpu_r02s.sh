#!/bin/bash
# r02s: pooled reducers (default build, SKB_ASSEMBLE=ws), late hook, register split; bitwise test; ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q > gpurun_out/r02s_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; tail -3 gpurun_out/r02s_pytest_variants.log
SKB_ASSEMBLE=ws AB_SUFFIX=_r02s_ws bash scripts/ab.sh main
AB_SUFFIX=_r02s bash scripts/ab.sh wsl wsnpl ws192u2
SKB_ASSEMBLE=ws timeout 600 ncu --clock-control none --set full --import-source on -k regex:'assemble_ws|finalize_blocks' -s 4 -c 2 -f \
    -o gpurun_out/r02s_ws python bench.py --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/r02s_ws_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02s_ws.ncu-rep
