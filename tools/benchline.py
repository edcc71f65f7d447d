"""Print the key timings of bench.py JSON lines found in the given log files."""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if l.startswith("{"):
            j = json.loads(l)
            if "ms_per_step" not in j: print(path, l[:200]); continue
            print(path, "n=%s step=%.3f kernel=%.3f ib=%s e2e=%.3f launches=%s" % (j.get("n_gpus"), j["ms_per_step"], (j.get("roofline") or {}).get("ms_per_launch", 0) or 0,
                  (j.get("ib") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("ms_per_step", 0) or 0, j.get("gpu_launches")))
