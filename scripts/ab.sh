#!/bin/bash
# A/B timing of development builds: scripts/ab.sh tag1 tag2 ...  ("main" = the default library)
# AB_ARGS="--element-order pencil" adds bench arguments; AB_SUFFIX names the output files apart
for tag in "$@"; do
  if [ "$tag" = main ]; then unset SKB_LIB_TAG; else export SKB_LIB_TAG=$tag; fi
  SKB_VERBOSE=1 python bench.py --newton 0 --no-cpu --no-e2e --steps 10 --warmup 3 $AB_ARGS > gpurun_out/ab_$tag$AB_SUFFIX.json 2> gpurun_out/ab_$tag$AB_SUFFIX.err
  grep simkit_b200: gpurun_out/ab_$tag$AB_SUFFIX.err | head -1
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$tag$AB_SUFFIX.json")); r=d["roofline"]
    print("$tag$AB_SUFFIX", "step %.3f ms" % d["ms_per_step"], {k: round(v,3) for k,v in r["step_kernels_ms"].items()})
except Exception as ex:
    print("$tag$AB_SUFFIX", "FAILED", ex)
PY
done
