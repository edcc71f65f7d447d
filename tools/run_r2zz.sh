#!/bin/bash
# round 2, very last call: smoke + the byte-identical driver test on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 30 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 40 python -m pytest tests/test_gpu_overlay.py -q -m gpu -x -k "rbc_in_shear_driver_bit_identical" 2>&1 | tail -2
