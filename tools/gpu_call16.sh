#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 8 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro16.json "" "push_streams=1,sort_streams=1" 2>&1 | grep -v "^ *per lap" | tail -5
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro16.json')):
    print(r['setting'], {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items()})
    print('   ', r.get('us_per_launch'))
    print('   ', [ (q['lap_mod5'], q['ms'], q['push_us']) for q in r['per_lap']])
PY
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
tail -c 3000 gpurun_out/bench16.json; tail -5 gpurun_out/bench16.err
