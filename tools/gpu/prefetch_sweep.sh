#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro_pf.json "push_streams=1,sort_streams=1,push_block=256" "push_block=128" "push_block=256,push_streams=4,sort_streams=4" "push_block=128" 2>&1 | grep -v "^ *per lap" | tail -1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro_pf.json')):
    print(r['setting'], round(r['ms_per_lap'],3), [ (q['lap_mod5'], q['push_us'], q['ms']) for q in r['per_lap']])
PY
