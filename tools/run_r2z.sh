#!/bin/bash
# round 2, last call: BIB_OPT_INTERP_COMPACT on by default - single-slab and multi-slab parity files
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 105 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -q -m gpu -x > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_tests.log
