#!/bin/bash
# round 2: in-place streaming, shift sweep on ONE box (lattice kernel alone, then the coupled step), and the largest one-GPU lattice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2r_exp.jsonl; : > $out
for s in 0 9 17 33 65 129; do
  timeout 120 python tools/exp_lbm.py --tag "shift$s" --inplace $s --rbcs 0 --seconds 1.0 >> $out 2>> gpurun_out/r2r_exp.err
done
timeout 120 python tools/exp_lbm.py --tag "shift17_nopf" --inplace 17 --rbcs 0 --seconds 1.0 --prefetch 0 >> $out 2>> gpurun_out/r2r_exp.err
timeout 150 python tools/exp_lbm.py --tag "step_ab" --inplace 0 --seconds 1.5 >> $out 2>> gpurun_out/r2r_exp.err
timeout 150 python tools/exp_lbm.py --tag "step_inplace33" --inplace 33 --seconds 1.5 >> $out 2>> gpurun_out/r2r_exp.err
cat $out | cut -c1-400
# 768 x 1024 x 1024 = 805 M cells on ONE GPU (two buffers would need 264 GB)
timeout 600 python bench.py --nx 768 --rbcs 7500 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2r_bench_768x1024x1024.json 2> gpurun_out/r2r_bench_768x1024x1024.err; echo "768x1024x1024 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2r_bench_768x1024x1024.json").read().strip().splitlines()[-1])
    print("768x1024x1024:", d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["device_bytes"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2r_bench_768x1024x1024.err
