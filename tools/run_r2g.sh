#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
{
python tools/exp_lbm.py --tag base --rbcs 0 --seconds 1.0
for v in ld1 ld2 st1 st2 ld1st1 ld2st1 minb6 minb4; do
  BIB_LIB=$PWD/build_exp/libbib_$v.so python tools/exp_lbm.py --tag $v --rbcs 0 --seconds 1.0
done
python tools/exp_lbm.py --tag base_again --rbcs 0 --seconds 1.0
} > gpurun_out/r2g_exp.jsonl 2> gpurun_out/r2g_exp.err
cat gpurun_out/r2g_exp.jsonl; tail -3 gpurun_out/r2g_exp.err
