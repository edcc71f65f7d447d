#!/usr/bin/env bash
# 2 GPUs: 1024^3 + 10 000 RBCs with in-place streaming (512 planes per GPU: two buffers would need 176 GB per GPU), parity gate in place too
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --global-nx 1024 --global-rbcs 10000 --inplace 17 > gpurun_out/r2s_1024cubed_n$N.json 2> gpurun_out/r2s_1024cubed_n$N.err
echo "rc=$?"; tail -c 800 gpurun_out/r2s_1024cubed_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2s_1024cubed_n$N.json").read().strip().splitlines()[-1])
    print("1024^3:", d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["device_bytes"], d["e2e"]["ms_per_step"], d["phases_ms"], d["parity_check"]["max_rel_err"], d["parity_check"]["pass"])
except Exception as e:
    print("no bench line", e)
PY
