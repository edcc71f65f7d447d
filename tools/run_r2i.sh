#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; tail -c 400 gpurun_out/r2i_bench_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; tail -c 300 gpurun_out/r2i_bench_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_collide_stream -s 6 -c 3 --csv --log-file gpurun_out/r2i_lbm_traffic.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --min-seconds 0.01 > gpurun_out/r2i_lbm_traffic.log 2>&1
tail -4 gpurun_out/r2i_lbm_traffic.csv
