#!/usr/bin/env bash
# 2 GPUs, one process per GPU: pull halo over IPC-mapped peer memory. Weak bench (two buffers, parity gate) and 1024^3 in place.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2u_weak_n$N.json 2> gpurun_out/r2u_weak_n$N.err
echo "weak rc=$?"; tail -c 600 gpurun_out/r2u_weak_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --global-nx 1024 --global-rbcs 10000 --inplace 17 --no-parity > gpurun_out/r2u_1024cubed_n$N.json 2> gpurun_out/r2u_1024cubed_n$N.err
echo "1024^3 rc=$?"; tail -c 600 gpurun_out/r2u_1024cubed_n$N.err
python - <<PY
import json
for f in ["gpurun_out/r2u_weak_n$N.json", "gpurun_out/r2u_1024cubed_n$N.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["e2e"]["ms_per_step"], d["phases_ms"], d["parity_check"] and (d["parity_check"]["max_rel_err"], d["parity_check"]["pass"]))
    except Exception as e:
        print("no bench line", f, e)
PY
