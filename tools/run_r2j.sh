#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
{
python tools/exp_lbm.py --tag base --rbcs 0 --seconds 1.0
for v in pf2 pf4 pf8 pf16 pf32 pf64; do
  BIB_LIB=$PWD/build_exp/libbib_$v.so python tools/exp_lbm.py --tag $v --rbcs 0 --seconds 1.0
done
python tools/exp_lbm.py --tag base_again --rbcs 0 --seconds 1.0
} > gpurun_out/r2j_exp.jsonl 2> gpurun_out/r2j_exp.err
cut -c1-200 gpurun_out/r2j_exp.jsonl; tail -3 gpurun_out/r2j_exp.err
