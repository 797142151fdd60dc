#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro33.json "push_streams=1,sort_streams=1" "push_streams=4,sort_streams=4,sort_overlap=0" "sort_overlap=1" 2>&1 | grep -v "^ *per lap" | tail -4
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro33.json')):
    print(r['setting'], round(r['ms_per_lap'],3), {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items() if v>0.25})
    print('   ', [ (q['lap_mod5'], q['ms'], q['push_us']) for q in r['per_lap']])
PY
