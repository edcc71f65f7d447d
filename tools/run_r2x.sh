#!/usr/bin/env bash
# round 2: ncu launch list of the in-place layout (8 launches of k_collide_stream per step + k_ghost_wrap) with DRAM bytes per launch
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 140 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_collide_stream|k_ghost_wrap" -s 20 -c 60 --csv --log-file gpurun_out/r2x_inplace_launches.csv \
    python tools/exp_lbm.py --inplace 17 --rbcs 0 --seconds 0.05 > gpurun_out/r2x_inplace.log 2>&1
echo "rc=$?"; tail -c 300 gpurun_out/r2x_inplace.log; wc -l gpurun_out/r2x_inplace_launches.csv
