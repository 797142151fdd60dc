#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_streams.json "push_streams=4,sort_streams=4" "push_streams=6,sort_streams=6" "push_streams=8,sort_streams=8" "push_streams=8,sort_streams=4" "push_streams=2,sort_streams=2" "push_streams=4,sort_streams=4,push_group=4" 2>&1 | grep -v "^ *per lap" | tail -8
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro_streams.json')):
    print(r['setting'], round(r['ms_per_lap'],3), [ (q['lap_mod5'], q['ms']) for q in r['per_lap']])
PY
