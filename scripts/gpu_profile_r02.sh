#!/bin/bash
# r02 profiles (1 GPU): launch list of the bench command + full captures of every hot kernel.
# Summaries are made on the CPU box from the .ncu-rep files (scripts/ncu_summary.py) and copied to profiles/.
mkdir -p gpurun_out
TAG=${1:-r02p}
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (cold-cache, serialised: shares of the step)
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-closures > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log | cut -c1-300
# 2. the assembly step
timeout 900 $NCU --set full --import-source on -k regex:'assemble_pipelined|finalize_blocks|finalize_verts' -s 9 -c 3 -f \
    -o gpurun_out/${TAG}_assemble python bench.py --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/${TAG}_assemble.log 2>&1
# 3. single-GPU Newton kernels (first PCG iterations of the second step)
timeout 900 $NCU --set full --import-source on -k regex:'energy_kernel|pcg_spmv_dot|pcg_update|pcg_direction|coarse_restrict|coarse_gemv|coarse_add' \
    -s 20 -c 8 -f -o gpurun_out/${TAG}_newton python scripts/diag_kernels.py newton C5 40 > gpurun_out/${TAG}_newton.log 2>&1
# 4. single-reduction distributed solve at full size on one rank
timeout 900 $NCU --set full --import-source on -k regex:'pcg2_' -s 12 -c 5 -f \
    -o gpurun_out/${TAG}_pcg2 python scripts/diag_pcg2.py --max-pcg 8 --solver pcg2_eager > gpurun_out/${TAG}_pcg2.log 2>&1
# 5. reduced Hessian, C4, r = 200
timeout 900 $NCU --set full --import-source on -k regex:'reduced_pass' -s 2 -c 2 -f \
    -o gpurun_out/${TAG}_reduced python scripts/diag_kernels.py reduced C4 200 > gpurun_out/${TAG}_reduced.log 2>&1
# 6. FST
timeout 600 $NCU --set full --import-source on -k regex:'fst_precompute|gemv_rows' -c 2 -f \
    -o gpurun_out/${TAG}_fst python scripts/diag_kernels.py fst 20000 200 > gpurun_out/${TAG}_fst.log 2>&1
# clean timings of the same drivers (no profiler)
timeout 300 python scripts/diag_kernels.py reduced C4 200 > gpurun_out/${TAG}_reduced_clean.log 2>&1; tail -2 gpurun_out/${TAG}_reduced_clean.log
timeout 300 python scripts/diag_kernels.py fst 20000 200 > gpurun_out/${TAG}_fst_clean.log 2>&1; tail -2 gpurun_out/${TAG}_fst_clean.log
# summaries on the box (the .ncu-rep files together exceed what gpurun brings back); the assembly capture is kept
for k in assemble newton pcg2 reduced fst; do
  if [ -f gpurun_out/${TAG}_$k.ncu-rep ]; then
    python scripts/ncu_summary.py gpurun_out/${TAG}_$k.ncu-rep gpurun_out/${TAG}_${k}_ncu_summary.txt > /dev/null 2>&1 || echo "summary of $k failed"
    [ $k = assemble ] || rm -f gpurun_out/${TAG}_$k.ncu-rep
  fi
done
ls -la gpurun_out/ | grep ${TAG}
