#!/bin/bash
# r02am: reduced API time with the work arrays kept in the plan (default) vs allocated per call
mkdir -p gpurun_out
for k in 1 0 1 0; do
  echo "SKB_REDUCED_KEEP=$k"
  SKB_REDUCED_KEEP=$k timeout 300 python scripts/diag_kernels.py reduced C4 200 8 2>&1 | grep "reduced r=" | awk '{print $4}' | tr '\n' ' '; echo
done
