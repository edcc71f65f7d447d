#!/usr/bin/env bash
# round 2, first GPU session: new tests + the bench at N = 1
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lightweight or band_reset or graph" --timeout=600 2>&1 | tail -5
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 6000 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err
