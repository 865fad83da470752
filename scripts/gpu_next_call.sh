#!/bin/bash
# First GPU call after a GPU-less stretch (run under gpurun, ONE GPU):
#   here, before the call:   SKB_BUILD_TAG=sum0 SKB_BUILD_FLAGS=-DSKB_EXP_SUM0 python -m simkit_b200.build
#                            SKB_BUILD_TAG=fin2 SKB_BUILD_FLAGS=-DSKB_FIN_ITEMS=2 python -m simkit_b200.build   (and fin4)
#                            SKB_BUILD_TAG=srcbase SKB_BUILD_FLAGS=-DSKB_EXP_SRCBASE python -m simkit_b200.build
#   then:                    gpurun --timeout 1500 -- 'bash scripts/gpu_next_call.sh r02a'
# 1. the GPU tests that were written without a GPU (quadratic term, Dirichlet Laplacian), then the whole GPU suite;
# 2. the default bench line;  3. the prepared A/B experiments on the clock:
#    corner-0 pairs by read-back (sum0), elements listed in 3x3-cell pencils (pencil), both together.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_quadratic.py tests/test_reference_properties.py -m gpu -q > gpurun_out/${TAG}_pytest_new.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_new.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json
TAGS="main"
for t in sum0 srcbase fin2 fin4; do [ -f simkit_b200/libsimkit_b200_$t.so ] && TAGS="$TAGS $t"; done
bash scripts/ab.sh $TAGS
AB_ARGS="--element-order pencil" AB_SUFFIX=_pencil bash scripts/ab.sh $TAGS
