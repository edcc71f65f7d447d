#!/usr/bin/env bash
# 4 GPUs, one process per GPU: pull halo with two DISTINCT neighbours per rank (IPC mappings of both), weak bench + parity gate (small cases)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-4}
BIB_GATE_SKIP_SUSPENSION=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus $N --steps 20 --warmup 5 --min-seconds 1.0 > gpurun_out/r2v_weak_n$N.json 2> gpurun_out/r2v_weak_n$N.err
echo "weak rc=$?"; tail -c 600 gpurun_out/r2v_weak_n$N.err
python - <<PY
import json
for f in ["gpurun_out/r2v_weak_n$N.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["e2e"]["ms_per_step"], d["phases_ms"], d["parity_check"] and (d["parity_check"]["max_rel_err"], d["parity_check"]["pass"]))
    except Exception as e:
        print("no bench line", f, e)
PY
