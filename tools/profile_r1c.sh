#!/usr/bin/env bash
# Round-1 (last session) profile capture, run under gpurun on ONE GPU:
#   1. launch list of the bench command (gpu__time_duration.sum per launch)
#   2. ncu --set full of two k_collide_stream launches of the coupled step (fused periodic-face tail)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_collide_stream -s 4 -c 2 -f -o gpurun_out/prof_lbm_r1d python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r1d.log 2>&1
ncu -i gpurun_out/prof_lbm_r1d.ncu-rep --page raw --csv > gpurun_out/prof_lbm_r1d_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
