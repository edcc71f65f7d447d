#!/usr/bin/env bash
# 8 GPUs: weak-scaling bench (1024^3 + 10 000 RBCs) and the strong-scaling point of the 256 x 1024 x 1024 lattice
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_weak_n$N.json 2> gpurun_out/r2f_weak_n$N.err
tail -c 600 gpurun_out/r2f_weak_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-parity > gpurun_out/r2f_strong_n$N.json 2> gpurun_out/r2f_strong_n$N.err
tail -c 600 gpurun_out/r2f_strong_n$N.err
