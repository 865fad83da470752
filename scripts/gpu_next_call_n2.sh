#!/bin/bash
# First multi-GPU call after a GPU-less stretch:  gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_next_call_n2.sh r02b 2'
# 1. the two-GPU tests (distributed Newton; the C++-driven PCG of capi_nccl.cu against the Python-driven loop);
# 2. bench at N GPUs with the Python-driven PCG and with SKB_NATIVE_NCCL=1 (same iterates, fewer host round trips).
TAG=${1:-r02b}
N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/${TAG}_pytest_n2.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_n2.log
for mode in python native graph; do
  unset SKB_NATIVE_NCCL
  [ $mode = native ] && export SKB_NATIVE_NCCL=1
  [ $mode = graph ] && export SKB_NATIVE_NCCL=2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --no-cpu > gpurun_out/${TAG}_bench_n${N}_$mode.json 2> gpurun_out/${TAG}_bench_n${N}_$mode.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_n${N}_$mode.json")); nw = d.get("newton", {})
    print("$mode", "step %.3f ms" % d["ms_per_step"], "newton %.2f steps/s, %s PCG iterations, %.3f ms per iteration" % (
        nw.get("steps_per_s", float("nan")), nw.get("pcg_iters"), nw.get("pcg_ms_per_iter", float("nan"))))
except Exception as ex:
    print("$mode", "FAILED", ex)
PY
done
