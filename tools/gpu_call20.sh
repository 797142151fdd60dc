#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
( timeout 1200 python -m pytest tests -m gpu -x -q -k "sort" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python tools/microbench.py --cells 256 --laps 5 --out gpurun_out/micro20.json "push_streams=1,sort_streams=1" "push_minb=6,push_streams=4,sort_streams=4" 2>&1 | grep -v "^ *per lap" | tail -8
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro20.json')):
    print(r['setting'], round(r['ms_per_lap'],3), {k:round(v,3) for k,v in r['ms_per_lap_by_class'].items()})
    print('   ', r.get('us_per_launch'))
    print('   ', [ (q['lap_mod5'], q['ms'], q['push_us']) for q in r['per_lap']])
PY
export B2P_OPTS=push_streams=1,sort_streams=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sort|k_gather|k_max" -s 96 -c 6 -o /tmp/ncu/sort -f python tools/microbench.py --cells 128 --laps 1 "" > gpurun_out/ncu20.log 2>&1
ncu -i /tmp/ncu/sort.ncu-rep --page details > gpurun_out/r20_sort_details.txt 2>/dev/null
ncu -i /tmp/ncu/sort.ncu-rep --page source --csv > gpurun_out/r20_sort_source.csv 2>/dev/null
tail -3 gpurun_out/ncu20.log
