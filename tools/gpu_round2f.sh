#!/bin/bash
# 2-GPU call: C++ driver --gpus 2, fusion, driver tests; default bench at N=2; C3 (64-view scene) at N=2
python -m pytest tests -m gpu -q -k "cpp_driver or fusion or pruning" 2>&1 | tail -25 > gpurun_out/r2f_pytest_gpu.log; tail -8 gpurun_out/r2f_pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2f_bench_c2_n2.json 2> gpurun_out/r2f_bench_c2_n2.err; echo "c2 n2 rc=$?"; tail -2 gpurun_out/r2f_bench_c2_n2.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config C3 --steps 1 --warmup 3 > gpurun_out/r2f_bench_c3_n2.json 2> gpurun_out/r2f_bench_c3_n2.err; echo "c3 n2 rc=$?"; tail -3 gpurun_out/r2f_bench_c3_n2.err | cut -c1-300
for f in c2_n2 c3_n2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench_$f.json").read().replace("NaN","null"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, d.get("e2e",{}).get("value"), d.get("nccl_allgather"), d.get("depth_within_1pct_of_ground_truth"), d.get("prior_host_s_not_hidden_per_step"))
except Exception as e: print("$f", "unreadable", e)
PY
done
