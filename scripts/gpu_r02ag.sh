#!/bin/bash
# r02ag: register-tiled FST precompute kernel vs the simple one (bit identity test, FST parity tests, kernel times by ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py::test_fst_precompute_kernels_bit_identical tests/test_gpu_solvers.py -m gpu -x -q -k "fst or sandwich or reduced" > gpurun_out/r02ag_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02ag_pytest.log
for mode in simple tiled; do
  SKB_FST=$mode timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active -k regex:fst_precompute --csv \
      --log-file gpurun_out/r02ag_fst_$mode.csv python scripts/diag_kernels.py fst 20000 200 > gpurun_out/r02ag_fst_$mode.log 2>&1
  grep -h "fst_precompute" gpurun_out/r02ag_fst_$mode.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200 | head -4
done
