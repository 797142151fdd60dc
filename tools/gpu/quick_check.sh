#!/bin/bash
# parity suite + per-class microbench at 256^3 (a minute of box time)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_quick.json "push_streams=1,sort_streams=1" "push_streams=4,sort_streams=4" 2>&1 | grep -v "^ *per lap" | tail -3
python - <<PY
import json
for r in json.load(open('gpurun_out/micro_quick.json')):
    print(r['setting'], round(r['ms_per_lap'],3), {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items()})
PY
