#!/bin/bash
# round 2, in-place streaming: parity tests, bench on the standard slab in place, 768^3 on one GPU (does not fit with two buffers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inplace.py -q -m gpu -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2q_tests.log
timeout 400 python bench.py --inplace 17 --no-cpu-baseline > gpurun_out/r2q_bench_inplace_n1.json 2> gpurun_out/r2q_bench_inplace_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2q_bench_inplace_n1.json").read().strip().splitlines()[-1])
    print("inplace n1:", d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d["roofline"]["sustained"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["device_bytes"], d["parity_check"])
except Exception as e:
    print("no bench line", e)
PY
timeout 500 python bench.py --nx 768 --ny 768 --nz 768 --rbcs 4219 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2q_bench_768.json 2> gpurun_out/r2q_bench_768.err; echo "768 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2q_bench_768.json").read().strip().splitlines()[-1])
    print("768^3:", d["ms_per_step"], d["value"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["device_bytes"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2q_bench_768.err
