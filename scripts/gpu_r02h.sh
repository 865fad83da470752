#!/bin/bash
# r02h (1 GPU): whole GPU suite (lazy Hessian / Jacobian, recompute interface, energy count fix); default bench line
# (e2e through the reference names, closure-based Newton leg); eager per-kind timing of the single-reduction solve
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02h_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 2500 gpurun_out/r02h_bench.json; tail -5 gpurun_out/r02h_bench.err
timeout 300 python scripts/diag_pcg2.py --steps 2 --solver pcg2_eager > gpurun_out/r02h_diag_eager.log 2>&1; tail -4 gpurun_out/r02h_diag_eager.log
