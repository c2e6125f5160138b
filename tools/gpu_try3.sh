#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final_bench_c2_n2.json 2> gpurun_out/final_bench_c2_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench_c2_n2.json").read().replace("NaN","null"))
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, "e2e", d["e2e"]["value"], d.get("nccl_allgather"), d.get("clocks"))
PY
