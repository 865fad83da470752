#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full
# capture of the dominant kernel.  Outputs land in gpurun_out/ (copy the summaries to profiles/).
set -x
TAG=${1:-r01}
WL=${2:-C5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.csv
timeout 900 python bench.py --workload $WL > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
# launch list (cold-cache, serialised): shares of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'assemble_|finalize_' -c 30 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --workload $WL --steps 3 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -12 gpurun_out/${TAG}_launches.csv
# full capture of the two kernels of the step (after the warm-up launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'assemble_pipelined|finalize_blocks' -s 6 -c 2 \
    -o gpurun_out/${TAG}_full -f python bench.py --workload $WL --steps 1 --warmup 3 --newton 0 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
