#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for i in 1 2; do python -m pytest tests -m gpu -q --timeout=1500 2>&1 | grep -E "BibError|^E  |passed|failed|FAILED" | head -12; done
bash tools/profile_r2.sh > /dev/null 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -c 200 gpurun_out/r2m_bench_n1.err
