#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
{
python tools/exp_lbm.py --tag pf_default --seconds 1.5
python tools/exp_lbm.py --tag pf_off --prefetch 0 --seconds 1.5
python tools/exp_lbm.py --tag pf_24k --prefetch 24576 --rbcs 0 --seconds 1.0
python tools/exp_lbm.py --tag pf_48k --prefetch 49152 --rbcs 0 --seconds 1.0
} > gpurun_out/r2k_exp.jsonl 2> gpurun_out/r2k_exp.err
cut -c1-420 gpurun_out/r2k_exp.jsonl; tail -3 gpurun_out/r2k_exp.err
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q --timeout=900 2>&1 | tail -3
