#!/bin/bash
# bash scripts/run_bench_n.sh N   -- the driver's multi-GPU launch of bench.py (one rank per GPU, NCCL), log kept under gpurun_out/
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], j["e2e"].get("numa"), "e2e_sync", j["e2e_sync"]["value"])
print("c4", j["c4_sweep"]["seconds"], j["c4_sweep"]["value"], j["c4_sweep"]["rmse_checksum"], j["c4_sweep"]["model_build_s"])
PY
