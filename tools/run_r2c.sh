#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py tests/test_gpu_parity.py -m gpu -q -k "multi or slab or two_devices" --timeout=900 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -c 1500 gpurun_out/r2c_bench_n2.err
