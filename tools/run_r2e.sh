#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 2>&1 | tail -12
python tools/driver_speed.py > gpurun_out/r2e_driver_speed.json 2>&1; tail -3 gpurun_out/r2e_driver_speed.json
