#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 500 > gpurun_out/r2b_clocks.csv & echo $! > /tmp/smi.pid )
{
python tools/exp_lbm.py --tag exact
BIB_LIB=$PWD/build_exp/libbib_fmad.so python tools/exp_lbm.py --tag fmad
python tools/exp_lbm.py --tag exact_noquirk --envelope 0 --rbcs 0
BIB_LIB=$PWD/build_exp/libbib_fmad.so python tools/exp_lbm.py --tag fmad_noquirk --envelope 0 --rbcs 0
python tools/exp_lbm.py --tag exact_fs2 --face-swap 2 --rbcs 0
python tools/exp_lbm.py --tag exact_fs0 --face-swap 0 --rbcs 0
} > gpurun_out/r2b_exp.jsonl 2> gpurun_out/r2b_exp.err
kill $(cat /tmp/smi.pid)
cat gpurun_out/r2b_exp.jsonl; tail -3 gpurun_out/r2b_exp.err
