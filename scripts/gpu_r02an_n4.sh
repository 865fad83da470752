#!/bin/bash
# r02m (N GPUs): default configuration after the sliced coarse solve, at three coarse-space sizes
N=${1:-8}
mkdir -p gpurun_out
run() {  # tag, extra args
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --no-cpu "$@" > gpurun_out/r02an_bench_n${N}_$tag.json 2> gpurun_out/r02an_bench_n${N}_$tag.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02an_bench_n${N}_$tag.json")); nw = d.get("newton", {})
    print("$tag", "step %.3f ms" % d["ms_per_step"], "newton %.2f steps/s, %s PCG iterations, %.3f ms per iteration" % (
        nw.get("steps_per_s", float("nan")), nw.get("pcg_iters"), nw.get("pcg_ms_per_iter", float("nan"))), nw.get("solver"), nw.get("solve_ms"), (d.get("parity_check") or {}).get("ok"), (d.get("parity_check") or {}).get("newton"))
except Exception as ex:
    print("$tag", "FAILED", ex)
PY
  tail -2 gpurun_out/r02an_bench_n${N}_$tag.err | cut -c1-300
}
run default

