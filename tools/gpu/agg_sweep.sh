#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_agg.json "push_streams=1,sort_streams=1,agg_min=2" "agg_min=3" "agg_min=1" "agg_min=2,push_streams=4,sort_streams=4" 2>&1 | grep -v "^ *per lap" | tail -9
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro_agg.json')):
    print(r['setting'], round(r['ms_per_lap'],3), [ (q['lap_mod5'], q['push_us']) for q in r['per_lap']])
PY
( timeout 1200 python -m pytest tests -m gpu -x -q -k "deposit or lap or golden or reflector" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 3 gpurun_out/pytest_gpu.log
