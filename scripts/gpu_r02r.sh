#!/bin/bash
# r02r: bitwise agreement of the three assembly kernels; A/B of the warp-specialised kernel (reducers alone / compute
# alone / register split / slot-per-thread level 2); one full ncu capture of assemble_ws_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q > gpurun_out/r02r_pytest_variants.log 2>&1
echo "pytest variants rc=$?"; tail -3 gpurun_out/r02r_pytest_variants.log
AB_SUFFIX=_r02r bash scripts/ab.sh wsnop2 wsnomath wsfs wsfs256 ws208
SKB_ASSEMBLE=ws timeout 600 ncu --clock-control none --set full --import-source on -k regex:'assemble_ws|finalize_blocks' -s 8 -c 2 -f \
    -o gpurun_out/r02r_ws python bench.py --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/r02r_ws_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02r_ws.ncu-rep
