#!/usr/bin/env bash
# round 2 profiling pass (one B200): launch list of the bench command, full capture of the lattice kernel, metrics of the IB kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --min-seconds 0.01 > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_collide_stream -s 4 -c 2 -o gpurun_out/r2_prof_lbm \
    python tools/exp_lbm.py --seconds 0.05 > gpurun_out/r2_prof_lbm.log 2>&1
ncu -i gpurun_out/r2_prof_lbm.ncu-rep --page raw --csv > gpurun_out/r2_k_collide_stream_ncu_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:"k_interp_tile|k_spread|k_rbc_forces|k_band|k_move|k_raycast|k_phase|k_lightweight" -s 12 -c 16 --csv --log-file gpurun_out/r2_ib_kernels_ncu.csv \
    python tools/exp_lbm.py --seconds 0.05 > gpurun_out/r2_ib_kernels.log 2>&1
ls -la gpurun_out | tail -12
