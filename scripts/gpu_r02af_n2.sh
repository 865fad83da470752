#!/bin/bash
# r02af (N GPUs): BASELINE config 4 sharded: C4 mesh, reduced Hessian r = 200, one all-reduce of 1 + r + r^2 doubles
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload C4 --reduced 200 --newton 0 --no-cpu > gpurun_out/r02af_bench_c4_n${N}.json 2> gpurun_out/r02af_bench_c4_n${N}.err
echo "rc=$?"; tail -3 gpurun_out/r02af_bench_c4_n${N}.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/r02af_bench_c4_n${N}.json")); print("step %.3f ms" % d["ms_per_step"], json.dumps(d.get("reduced"))[:900], d.get("parity_check"))
PY
