#!/bin/bash
# round 2: pull halo (k_halo_pull) in the in-process slab group + in-place streaming, on one GPU; smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_inplace.py -q -m gpu -x > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2t_tests.log
BIB_HALO=copy timeout 600 python -m pytest tests/test_gpu_slabs.py -q -m gpu -x -k "lattice" > gpurun_out/r2t_tests_copy.log 2>&1; echo "copy-transport tests rc=$?"; tail -2 gpurun_out/r2t_tests_copy.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
