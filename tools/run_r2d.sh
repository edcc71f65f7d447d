#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q -x --timeout=900 2>&1 | tail -8
python tools/exp_lbm.py --tag fused_interp > gpurun_out/r2d_exp.jsonl 2> gpurun_out/r2d_exp.err
cat gpurun_out/r2d_exp.jsonl; tail -3 gpurun_out/r2d_exp.err
