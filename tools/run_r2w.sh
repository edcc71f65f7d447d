#!/bin/bash
# round 2, final tree: the whole GPU suite, smoke, the default bench line (what the driver runs at round end)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2w_tests.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
    print("n1:", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d.get("population_storage", d["config"].get("population_storage")), d["e2e"]["ms_per_step"], d["parity_check"]["pass"], d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
