#!/bin/bash
# round 2: BIB_OPT_INTERP_COMPACT - parity (own test + the coupled-step goldens with the option forced on through the environment) and timing on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "compact_interpolation" > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2y_tests.log
BIB_INTERP_COMPACT=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -q -m gpu -x -k "full_step or fused_interpolation or rbc_in_shear_vs_golden or slab_immersed or call_fusion" > gpurun_out/r2y_tests_env.log 2>&1; echo "env tests rc=$?"; tail -3 gpurun_out/r2y_tests_env.log
out=gpurun_out/r2y_exp.jsonl; : > $out
for c in 0 1 0 1; do
  timeout 120 python tools/exp_lbm.py --tag "compact$c" --interp-compact $c --seconds 0.7 >> $out 2>> gpurun_out/r2y_exp.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2y_exp.jsonl"):
    d = json.loads(l)
    print(d["tag"], "step", round(d["step_ms_median"], 4), "interp", d["phases"].get("interpolate"), "lbm", round(d["lbm_ms_median"], 4))
PY
