#!/bin/bash
# 8 GPUs: default bench (C2, weak scaling) and C3 (64-view scene, strong scaling) after the pool / Park changes
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2s_bench_c2_n$N.json 2> gpurun_out/r2s_bench_c2_n$N.err
echo "c2 n$N rc=$?"; grep -v Warning gpurun_out/r2s_bench_c2_n$N.err | tail -3 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --config C3 --steps 1 --warmup 3 > gpurun_out/r2s_bench_c3_n$N.json 2> gpurun_out/r2s_bench_c3_n$N.err
echo "c3 n$N rc=$?"; grep -v Warning gpurun_out/r2s_bench_c3_n$N.err | tail -3 | cut -c1-300
python - <<PY
import json
for f in ("c2","c3"):
    try:
        d=json.loads(open("gpurun_out/r2s_bench_%s_n$N.json" % f).read().replace("NaN","null"))
        print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, "e2e", d.get("e2e",{}).get("value"), d.get("nccl_allgather"), d.get("clocks"))
    except Exception as e: print(f, "unreadable", e)
PY
