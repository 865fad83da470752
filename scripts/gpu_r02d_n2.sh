#!/bin/bash
# r02d (2 GPUs): the sharding tests incl. the single-reduction solve; bench at N with each distributed solver
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/r02d_pytest_n2.log 2>&1; tail -8 gpurun_out/r02d_pytest_n2.log
run() {  # tag, extra args
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --no-cpu "$@" > gpurun_out/r02d_bench_n${N}_$tag.json 2> gpurun_out/r02d_bench_n${N}_$tag.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02d_bench_n${N}_$tag.json")); nw = d.get("newton", {})
    print("$tag", "step %.3f ms" % d["ms_per_step"], "newton %.2f steps/s, %s PCG iterations, %.3f ms per iteration" % (
        nw.get("steps_per_s", float("nan")), nw.get("pcg_iters"), nw.get("pcg_ms_per_iter", float("nan"))), d.get("parity_check"))
except Exception as ex:
    print("$tag", "FAILED", ex)
PY
  tail -3 gpurun_out/r02d_bench_n${N}_$tag.err
}
run python --dist-solver python --no-parity
run pcg2
run pcg2_343 --aggregates 343 --no-parity
run pcg2_eager --dist-solver pcg2_eager --no-parity
