#!/bin/bash
# C3 (64-view scene, strong scaling) at the GPU count of this box: tools/gpu_c3.sh <N> [extra pytest -k expression]
N=$1
if [ -n "$2" ]; then python -m pytest tests -m gpu -q -k "$2" 2>&1 | tail -25 > gpurun_out/r2h_pytest_n$N.log; tail -6 gpurun_out/r2h_pytest_n$N.log; fi
if [ "$N" = 1 ]; then
  python bench.py --config C3 --steps 1 --warmup 3 > gpurun_out/r2h_bench_c3_n$N.json 2> gpurun_out/r2h_bench_c3_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --config C3 --steps 1 --warmup 3 > gpurun_out/r2h_bench_c3_n$N.json 2> gpurun_out/r2h_bench_c3_n$N.err
fi
echo "c3 n$N rc=$?"; grep -v Warning gpurun_out/r2h_bench_c3_n$N.err | tail -4 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_bench_c3_n$N.json").read().replace("NaN","null"))
    print("c3 n$N", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, "e2e", d.get("e2e",{}).get("value"), d.get("nccl_allgather"), d.get("depth_within_1pct_of_ground_truth"), d.get("prior_host_s_not_hidden_per_step"), d.get("clocks"))
except Exception as e: print("unreadable", e)
PY
