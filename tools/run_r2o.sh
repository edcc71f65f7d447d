#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2o_weak_n$N.json 2> gpurun_out/r2o_weak_n$N.err
tail -c 400 gpurun_out/r2o_weak_n$N.err
