#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro13.json "" 2>&1 | tail -5
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro13.json')):
    print(r['setting'], {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items()})
PY
timeout 1200 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; tail -c 3000 gpurun_out/bench13.json; tail -5 gpurun_out/bench13.err
