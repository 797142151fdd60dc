#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro12.json \
  "sort_streams=1" "sort_streams=2" "sort_streams=4" 2>&1 | tee gpurun_out/micro12.log
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro12.json')):
    print(r['setting'], {k:v for k,v in r['ms_per_lap_by_class'].items()})
PY
