#!/usr/bin/env bash
# Round-1 (session 4) validation + profile capture, run under gpurun on ONE GPU:
#   1. the whole GPU test suite, smoke(), both bench arms (no profiler attached)
#   2. launch list of the bench command (gpu__time_duration.sum per launch)
#   3. ncu --set full of two k_collide_stream launches of the coupled step and of the IB kernels of one step
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r1e_pytest_gpu.log 2>&1; tail -4 gpurun_out/r1e_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1e_bench_ref.json 2> gpurun_out/r1e_bench_ref.err
python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; cat gpurun_out/r1e_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_r1e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide_stream -s 4 -c 2 -f -o gpurun_out/prof_lbm_r1e python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r1e.log 2>&1
ncu -i gpurun_out/prof_lbm_r1e.ncu-rep --page raw --csv > gpurun_out/prof_lbm_r1e_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread --clock-control none -k regex:"k_spread|k_band|k_interpolate|k_rbc_forces|k_move|k_raycast|k_phase" -c 16 --csv --log-file gpurun_out/ib_kernels_r1e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_ib_r1e.log 2>&1
ls -la gpurun_out | tail -8
