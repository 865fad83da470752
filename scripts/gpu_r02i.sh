#!/bin/bash
# r02i (1 GPU): rest of the GPU suite from the lazy tests on; closure-path profile; per-kind timing of the solve after
# the SpMV spill fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_parity.py tests/test_gpu_sharding.py tests/test_gpu_solvers.py tests/test_mfem.py tests/test_reference_properties.py tests/test_zz_quadratic.py -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02i_pytest_gpu.log
timeout 300 python scripts/diag_pcg2.py --steps 2 --solver pcg2_eager > gpurun_out/r02i_diag_eager.log 2>&1; tail -2 gpurun_out/r02i_diag_eager.log
timeout 600 python scripts/diag_closures.py C5 > gpurun_out/r02i_closures.log 2>&1; head -60 gpurun_out/r02i_closures.log
