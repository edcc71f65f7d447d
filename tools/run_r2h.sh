#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
{
python tools/exp_lbm.py --tag base --seconds 1.0
BIB_LIB=$PWD/build_exp/libbib_interp3.so python tools/exp_lbm.py --tag interp3 --seconds 1.0
BIB_L2_FETCH=32 python tools/exp_lbm.py --tag l2fetch32 --seconds 1.0
BIB_L2_FETCH=128 python tools/exp_lbm.py --tag l2fetch128 --seconds 1.0
} > gpurun_out/r2h_exp.jsonl 2> gpurun_out/r2h_exp.err
cat gpurun_out/r2h_exp.jsonl; tail -3 gpurun_out/r2h_exp.err
