#!/bin/bash
# r02c: GPU suite incl. the new config-size tests; default bench line; A/B of element orders and the combined switches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 1800 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
export AB_ARGS="--no-parity"
bash scripts/ab.sh main combo
SKB_ELEMENT_ORDER=input AB_SUFFIX=_input bash scripts/ab.sh main
AB_ARGS="--no-parity --shuffle elements" AB_SUFFIX=_shufE bash scripts/ab.sh main combo
AB_ARGS="--no-parity --shuffle both" AB_SUFFIX=_shufB bash scripts/ab.sh main
